"""Import the UNMODIFIED reference: from /root/reference (build container) or from its byte-compiled build
oracle/_ref/ (oracle/build_ref.py; travels to the GPU box, where /root/reference does not exist).

TEST INFRASTRUCTURE ONLY - used by `oracle/make_golden.py`, by the tests that pin `oracle/monkey_oracle.py` against
the live reference, by the drop-in test that runs the reference's own drivers against this repo's `modules/`
(tests/test_gpu_6_reference_drivers.py) and by `bench.py --impl reference` / `cpu_baseline`.  Never imported by the
product path.

Two compatibility shims, both restoring torch==0.4.1 behaviour the reference pins
(requirements.txt:24):
  1. `torch.gesv(B, A)` was removed -> `torch.linalg.solve(A, B)` (used at modules/util.py:223).
  2. `F.grid_sample` called without `align_corners` meant align_corners=True in 0.4.1
     (modules/generator.py:57, modules/movement_embedding.py:85).
The reference's modules are loaded under private names into an isolated namespace so that they do
not shadow this repo's own drop-in `modules` / `sync_batchnorm` packages.
"""
import importlib
import os
import sys

import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    """$MONKEY_REF, the read-only source tree, the byte-compiled build next to this file, the driver's install dir."""
    cands = [os.environ.get('MONKEY_REF'), '/root/reference', os.path.join(_HERE, '_ref'),
             os.path.join(os.path.dirname(_HERE), 'baseline', '_ref')]
    for c in cands:
        if c and (os.path.exists(os.path.join(c, 'modules', 'util.py')) or
                  os.path.exists(os.path.join(c, 'modules', 'util.pyc'))):
            return c
    return cands[1]


REF_ROOT = _find_root()


def available():
    return os.path.exists(os.path.join(REF_ROOT, 'modules', 'util.py')) or \
        os.path.exists(os.path.join(REF_ROOT, 'modules', 'util.pyc'))


def has_sources():
    """True when the reference's .py text is readable (build container); the GPU box only has the compiled build."""
    return os.path.exists(os.path.join(REF_ROOT, 'transfer.py'))


def kind():
    return 'source tree' if has_sources() else 'byte-compiled build (oracle/build_ref.py)'


_loaded = None


def _install_torch_shims():
    if not hasattr(torch, 'gesv'):
        torch.gesv = lambda b, a: (torch.linalg.solve(a, b), None)
    if not getattr(F.grid_sample, '_monkey_shim', False):
        orig = F.grid_sample

        def grid_sample(inp, grid, mode='bilinear', padding_mode='zeros', align_corners=None):
            return orig(inp, grid, mode=mode, padding_mode=padding_mode,
                        align_corners=True if align_corners is None else align_corners)

        grid_sample._monkey_shim = True
        F.grid_sample = grid_sample


def load():
    """Returns a dict of the reference's module objects: generator, discriminator, keypoint_detector, util,
    movement_embedding, dense_motion_module, losses."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('reference tree not found at %s' % REF_ROOT)
    _install_torch_shims()

    # Temporarily make `modules` / `sync_batchnorm` resolve to the reference tree, import, then restore
    # whatever the repo had registered under those names.
    names = ('modules', 'sync_batchnorm')
    saved = {k: v for k, v in sys.modules.items() if k.split('.')[0] in names}
    for k in saved:
        del sys.modules[k]
    # the reference's `modules/` has no __init__.py (namespace package), so this repo's regular `modules` package
    # would win regardless of path order: hide every other path entry that carries a `modules` / `sync_batchnorm`
    old_path = list(sys.path)
    sys.path[:] = [REF_ROOT] + [p for p in old_path
                                if not any(os.path.isdir(os.path.join(p or '.', n)) for n in names)]
    try:
        out = {}
        for short in ('util', 'keypoint_detector', 'movement_embedding', 'dense_motion_module', 'generator',
                      'discriminator', 'losses'):
            out[short] = importlib.import_module('modules.' + short)
        ref_mods = {k: v for k, v in sys.modules.items() if k.split('.')[0] in names}
    finally:
        sys.path[:] = old_path
        for k in list(sys.modules):
            if k.split('.')[0] in names:
                del sys.modules[k]
        sys.modules.update(saved)
    out['_sys_modules'] = ref_mods
    _loaded = out
    return out


def build_from_config(config):
    """Reference nets built exactly like run.py:50-63."""
    ref = load()
    mp = config['model_params']
    gen = ref['generator'].MotionTransferGenerator(**mp['generator_params'], **mp['common_params'])
    disc = ref['discriminator'].Discriminator(**mp['discriminator_params'], **mp['common_params'])
    kp = ref['keypoint_detector'].KPDetector(**mp['kp_detector_params'], **mp['common_params'])
    return gen, disc, kp


# ------------------------------------------------------------------------------------------------ reference drivers
class _RecordingLogger:
    """Stand-in for the reference's logger.Logger (logger.py:11-88 needs imageio / skimage / matplotlib, absent here;
    file IO and visualisation are out of scope): same constructor, context-manager protocol and
    log_iter / log_epoch / load_cpk surface as train.py:90,107,140-155 uses; it records what it is handed."""
    records = None

    def __init__(self, log_dir=None, visualizer_params=None, **kwargs):
        self.iters, self.epochs = [], []
        _RecordingLogger.records = self

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def log_iter(self, it, names, values, inp, out):
        self.iters.append((it, list(names), [float(v) for v in values], {k: tuple(v.shape) for k, v in inp.items()},
                           sorted(out.keys())))

    def log_epoch(self, epoch, models):
        self.epochs.append((epoch, sorted(models.keys())))

    @staticmethod
    def load_cpk(*args, **kwargs):
        raise AttributeError('no checkpoints in this harness')


def load_driver(name, modules='product'):
    """Import the reference's own `train` / `transfer` / `reconstruction` module, UNCHANGED, bound to either
      modules='product'   - this repo's drop-in `modules/` + `sync_batchnorm/` (what a user switching over gets), or
      modules='reference' - the reference's own modules (the CPU ground truth).
    IO-only imports of the drivers (logger, imageio, frames_dataset; SURVEY 8(c) "stubs for drivers") are stubbed for
    the duration of the import.  Returns the module object (a fresh one per call)."""
    import types
    if not available():
        raise RuntimeError('reference not found (looked at %s)' % REF_ROOT)
    _install_torch_shims()
    stubs = {}
    lg = types.ModuleType('logger')
    lg.Logger = _RecordingLogger
    lg.Visualizer = type('Visualizer', (), {'__init__': lambda self, **kw: None})
    stubs['logger'] = lg
    stubs['imageio'] = types.ModuleType('imageio')
    fd = types.ModuleType('frames_dataset')
    fd.PairedDataset = type('PairedDataset', (), {})
    stubs['frames_dataset'] = fd
    # product binding: the repo's packages stay registered (they ARE what the driver must see); reference binding:
    # swap them out for the duration of the import, exactly like load()
    names = ('modules', 'sync_batchnorm') if modules == 'reference' else ()
    saved = {k: v for k, v in sys.modules.items() if k.split('.')[0] in names or k in stubs or k == name}
    for k in saved:
        del sys.modules[k]
    old_path = list(sys.path)
    repo = os.path.dirname(_HERE)
    try:
        sys.modules.update(stubs)
        if modules == 'reference':
            ref = load()
            sys.modules.update(ref['_sys_modules'])
            sys.path[:] = [REF_ROOT] + [p for p in old_path
                                        if not any(os.path.isdir(os.path.join(p or '.', n)) for n in names)]
        else:
            import modules.util, sync_batchnorm  # noqa: F401,E401  (the repo's: regular packages, repo root on sys.path)
            sys.path[:] = [repo] + [p for p in old_path if p != repo] + [REF_ROOT]
        mod = importlib.import_module(name)
        bound = sys.modules['sync_batchnorm'].__file__ or ''
        assert (modules == 'reference') == (os.path.abspath(bound).startswith(os.path.abspath(REF_ROOT))), bound
    finally:
        sys.path[:] = old_path
        for k in list(sys.modules):
            if k.split('.')[0] in names or k in stubs or k == name:
                del sys.modules[k]
        sys.modules.update(saved)
    return mod


class cpu_data_parallel:
    """Context manager: make torch's DataParallel (the reference's DataParallelWithCallback base class) behave as on a
    host without accelerators, so the reference's own train.py can be run on the CPU of a GPU box as ground truth."""

    def __enter__(self):
        import sys as _sys
        dp = _sys.modules['torch.nn.parallel.data_parallel']   # the attribute of that name on the package is a function
        self._dp, self._orig = dp, dp._get_available_device_type
        dp._get_available_device_type = lambda: None
        return self

    def __exit__(self, *exc):
        self._dp._get_available_device_type = self._orig
        return False
