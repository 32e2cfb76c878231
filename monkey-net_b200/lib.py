"""ctypes binding of libmonkey_b200.so (the C ABI declared in include/monkey_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.  Signature strings:
p = device/host pointer (void*), i = int, l = long long, f = float, d = double.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libmonkey_b200.so')

HEADER = os.path.join(os.path.dirname(HERE), 'include', 'monkey_b200.h')


def _parse_header(path=HEADER):
    """Derive the ctypes signatures from the C header so the binding cannot drift from the declared ABI."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    sigs = {}
    for m in re.finditer(r'\bint\s+(mk_\w+)\s*\(([^;]*?)\)\s*;', text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        sig = ''
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                if '*' in a:
                    sig += 'p'
                elif a.startswith('long long'):
                    sig += 'l'
                elif a.startswith('int'):
                    sig += 'i'
                elif a.startswith('float'):
                    sig += 'f'
                elif a.startswith('double'):
                    sig += 'd'
                else:
                    raise ValueError('unparsed C argument %r in %s' % (a, name))
        sigs[name] = sig
    return sigs


SIGNATURES = _parse_header()

_CT = {'p': ctypes.c_void_p, 'i': ctypes.c_int, 'l': ctypes.c_longlong, 'f': ctypes.c_float, 'd': ctypes.c_double}

_lib = None


class MonkeyLibError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MonkeyLibError(
            'libmonkey_b200.so not built (%s). Run `python -c "import __graft_entry__ as g; g.build()"` or '
            '`python monkey-net_b200/build.py`. There is no CPU / library fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.mk_last_error.restype = ctypes.c_char_p
    lib.mk_last_error.argtypes = []
    for name, sig in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = [_CT[ch] for ch in sig]
    _lib = lib
    return lib


class _Counter:
    """Counts C-ABI kernel-launching calls (bench.py reports it as gpu_launches)."""
    n = 0


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    _Counter.n += 1
    if rc != 0:
        raise MonkeyLibError('%s failed (%d): %s' % (name, rc, lib.mk_last_error().decode()))


def call_soft(name, soft, *args):
    """like call(), but the return codes in `soft` (an entry point declining a shape BEFORE touching device state,
    e.g. -2 'outside this kernel's envelope') are returned to the caller instead of raised."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc == 0:
        _Counter.n += 1
        return 0
    if rc in soft:
        return rc
    raise MonkeyLibError('%s failed (%d): %s' % (name, rc, lib.mk_last_error().decode()))


def query(name, *args):
    """host-side entry points that launch nothing (planner dry runs, size queries): not counted as kernel launches"""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise MonkeyLibError('%s failed (%d): %s' % (name, rc, lib.mk_last_error().decode()))


def launches():
    return _Counter.n
