"""Importable alias of the `monkey-net_b200/` package directory (a hyphen is not a valid Python identifier).

All code lives in `monkey-net_b200/`; this stub only extends its own search path to that directory.
"""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'monkey-net_b200'))

from .version import __version__  # noqa: E402,F401
