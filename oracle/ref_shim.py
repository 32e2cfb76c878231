"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY - used by `oracle/make_golden.py` and by the CPU tests that pin
`oracle/monkey_oracle.py` against the live reference.  /root/reference does not exist on the GPU
box, so nothing on the `-m gpu` / smoke / bench path may call this.

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

REF_ROOT = os.environ.get('MONKEY_REF', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'modules'))


_loaded = None


def load():
    """Returns a dict of the reference's module objects: generator, discriminator, keypoint_detector, util,
    movement_embedding, dense_motion_module, losses."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('reference tree not found at %s' % REF_ROOT)
    if not hasattr(torch, 'gesv'):
        torch.gesv = lambda b, a: (torch.linalg.solve(a, b), None)
    if not getattr(F.grid_sample, '_monkey_shim', False):
        orig = F.grid_sample

        def grid_sample(inp, grid, mode='bilinear', padding_mode='zeros', align_corners=None):
            return orig(inp, grid, mode=mode, padding_mode=padding_mode,
                        align_corners=True if align_corners is None else align_corners)

        grid_sample._monkey_shim = True
        F.grid_sample = grid_sample

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
