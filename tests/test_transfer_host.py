"""Host-side transfer glue (monkey_net_b200.transfer_step) against the reference's own `normalize_kp`
(transfer.py:31-62): the two function definitions are extracted from the UNMODIFIED reference source and executed
(build container only; the driver module itself cannot be imported here - it needs imageio / skimage / matplotlib)."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import ref_shim


def _reference_functions(names=('make_symetric_matrix', 'normalize_kp', 'transfer_one')):
    from scipy.spatial import ConvexHull
    from modules.util import matrix_inverse
    src = open(os.path.join(ref_shim.REF_ROOT, 'transfer.py')).read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(wanted) == len(names)
    if not hasattr(torch, 'gesv'):
        torch.gesv = lambda b, a: (torch.linalg.solve(a, b), None)
    ns = {'np': np, 'torch': torch, 'ConvexHull': ConvexHull, 'matrix_inverse': matrix_inverse}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), 'reference/transfer.py', 'exec'), ns)
    return ns


def _reference_normalize_kp():
    return _reference_functions()['normalize_kp']


@pytest.mark.skipif(not ref_shim.available(), reason='/root/reference only exists in the build container')
@pytest.mark.parametrize('name', ['shapes', 'moving-gif'])
def test_oracle_transfer_one_matches_reference_transfer_one(name):
    """The oracle's transfer_one (the checker of the product's transfer path in the GPU tests) against the reference's
    own `transfer_one` body (transfer.py:65-79) driving the reference's own modules, eval mode, 3 driving frames."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import helpers
    from oracle import monkey_oracle as mo
    cfg = helpers.load_config(name)
    torch.manual_seed(0)
    rg, rd, rk = ref_shim.build_from_config(cfg)
    og, od, ok = mo.build_from_config(cfg)
    helpers.perturb_flow_head(rg)
    og.load_state_dict(rg.state_dict()); ok.load_state_dict(rk.state_dict())
    for m in (rg, rk, og, ok):
        m.eval()
    src, drv = helpers.smooth_frames(1, 1, 64, 5), helpers.smooth_frames(1, 3, 64, 6)
    ref_transfer_one = _reference_functions()['transfer_one']
    with torch.no_grad():
        want = ref_transfer_one(rg, rk, src, drv, cfg['transfer_params'])
        got = mo.transfer_one(og, ok, src, drv, cfg['transfer_params']['normalization_params'])
    for k in ('video_prediction', 'video_deformed'):
        assert got[k].shape == want[k].shape
        assert float((got[k] - want[k]).abs().max()) < 1e-4, k
    for k in ('kp_driving', 'kp_source', 'kp_norm'):
        assert float((got[k]['mean'] - want[k]['mean']).abs().max()) < 1e-5, k


def _kps(b=2, d=3, k=5, seed=0):
    g = torch.Generator().manual_seed(seed)

    def spd(*shape):
        a = torch.randn(*shape, 2, 2, generator=g) * 0.1
        return a @ a.transpose(-1, -2) + 0.01 * torch.eye(2)
    video = {'mean': torch.rand(b, d, k, 2, generator=g) * 2 - 1, 'var': spd(b, d, k)}
    app = {'mean': torch.rand(b, 1, k, 2, generator=g) * 2 - 1, 'var': spd(b, 1, k)}
    return video, app


@pytest.mark.skipif(not ref_shim.available(), reason='/root/reference only exists in the build container')
@pytest.mark.parametrize('opts', [dict(), dict(move_location=True), dict(move_location=True, clip_mean=True),
                                  dict(move_location=True, movement_mult=True),
                                  dict(move_location=True, adapt_variance=True)])
def test_normalize_kp_matches_reference(opts):
    from monkey_net_b200 import transfer_step
    ref = _reference_normalize_kp()
    video, app = _kps()
    want = ref({k: v.clone() for k, v in video.items()}, {k: v.clone() for k, v in app.items()}, **opts)
    got = transfer_step.normalize_kp({k: v.clone() for k, v in video.items()}, {k: v.clone() for k, v in app.items()},
                                     **opts)
    assert set(got) == set(want)
    for k in want:
        assert got[k].shape == want[k].shape
        assert float((got[k] - want[k]).abs().max()) < 2e-5, (k, opts)


def test_normalize_kp_does_not_modify_inputs_and_broadcasts_source():
    from monkey_net_b200 import transfer_step
    video, app = _kps()
    keep = video['mean'].clone()
    out = transfer_step.normalize_kp(video, app, move_location=True)
    assert torch.equal(video['mean'], keep)
    assert torch.allclose(out['mean'][:, 0:1], app['mean'].expand_as(out['mean'][:, 0:1]), atol=1e-6)
