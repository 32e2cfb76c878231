"""The drop-in claim, executed: the reference's OWN `train.py` / `transfer.py` / `reconstruction.py` (imported
unchanged - byte-compiled build oracle/_ref on the GPU box, the source tree in the build container - through
oracle/ref_shim.load_driver; only the IO modules logger / imageio / frames_dataset are stubbed) run
against THIS repo's `modules/` + `sync_batchnorm/` on the B200, and the same driver code runs against the
reference's own modules on the host CPU as ground truth.  Same initial weights, same DataLoader order.

  train.train()            train.py:78-155   2 epochs x 2 iterations incl. torch.optim.Adam x3, MultiStepLR, DataLoader
  transfer.transfer_one()  transfer.py:65-79 per-frame keypoints, normalize_kp, per-frame generator
  reconstruction.generate  reconstruction.py:12-25
"""
import pytest
import torch

import helpers
from oracle import ref_shim

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_shim.available(), reason='reference (source tree or oracle/_ref build) absent')]


def _pairs(cfg):
    import test_gpu_2_modules as t2
    torch.manual_seed(0)
    rg, rd, rk = ref_shim.build_from_config(cfg)
    helpers.perturb_flow_head(rg)
    gen, disc, kp = t2.build_product(cfg)
    gen.load_state_dict(rg.state_dict()); disc.load_state_dict(rd.state_dict()); kp.load_state_dict(rk.state_dict())
    for m in (gen, disc, kp):
        m.cuda()
    return (rg, rd, rk), (gen, disc, kp)


@pytest.mark.parametrize('mode', ['fp32', 'tf32x3'])
def test_reference_train_py_runs_unchanged_on_the_drop_in(mode):
    from monkey_net_b200 import ops
    if mode not in ops.CONV_MODES:
        pytest.skip('conv mode %s not built' % mode)
    cfg = helpers.driver_config(num_epochs=2)
    (rg, rd, rk), (gen, disc, kp) = _pairs(cfg)
    w_before = gen.refinement_module[-1].weight.detach().clone()
    with ref_shim.cpu_data_parallel():
        want = helpers.run_reference_train(ref_shim.load_driver('train', modules='reference'), (rg, rd, rk), cfg, None)
    want_iters = list(want.iters)
    prev = ops.CONV_MODE
    ops.set_conv_mode(mode)
    try:
        got = helpers.run_reference_train(ref_shim.load_driver('train', modules='product'), (gen, disc, kp), cfg, [0])
    finally:
        ops.set_conv_mode(prev)
    assert len(got.iters) == len(want_iters) == 4 and [e[0] for e in got.epochs] == [0, 1]
    assert got.epochs[0][1] == ['discriminator', 'generator', 'kp_detector', 'optimizer_discriminator',
                                'optimizer_generator', 'optimizer_kp_detector']
    for it, (g, w) in enumerate(zip(got.iters, want_iters)):
        assert g[1] == w[1] and g[3] == w[3] and g[4] == w[4]          # loss names, input shapes, output keys
        rel = max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(g[2], w[2]))
        print('train.py iteration %d [%s]: losses %s vs CPU reference %s (rel %.2e)' % (it, mode, g[2], w[2], rel))
        # iteration 0: identical state.  Later: Adam's +-lr first steps on rounding-level gradients
        assert rel < (2e-3 if it == 0 else 1e-1), (it, g[2], w[2])
    assert float((gen.refinement_module[-1].weight.detach() - w_before).abs().max()) > 0, 'optimizer never stepped'
    assert all(torch.isfinite(p).all() for m in (gen, disc, kp) for p in m.parameters())


def test_reference_transfer_py_and_reconstruction_py_on_the_drop_in():
    cfg = helpers.driver_config()
    (rg, rd, rk), (gen, disc, kp) = _pairs(cfg)
    for m in (rg, rk, gen, kp):
        m.eval()
    src, drv = helpers.smooth_frames(2, 1, 32, 3), helpers.smooth_frames(2, 3, 32, 4)
    tparams = {'normalization_params': {'move_location': True, 'movement_mult': False, 'adapt_variance': False,
                                        'clip_mean': True}}
    ref_t = ref_shim.load_driver('transfer', modules='reference')
    our_t = ref_shim.load_driver('transfer', modules='product')
    ref_r = ref_shim.load_driver('reconstruction', modules='reference')
    our_r = ref_shim.load_driver('reconstruction', modules='product')
    with torch.no_grad():
        a = ref_t.transfer_one(rg, rk, src, drv, tparams)
        b = our_t.transfer_one(gen, kp, src.cuda(), drv.cuda()[:, :, :], tparams)   # D-slices stay non-contiguous inside
        e_pred = helpers.max_abs(a['video_prediction'], b['video_prediction'])
        e_kp = helpers.max_abs(a['kp_norm']['mean'], b['kp_norm']['mean'])
        print('transfer.py transfer_one: |frame| %.2e |kp| %.2e' % (e_pred, e_kp))
        assert e_pred < 1e-3 and e_kp < 1e-4 and b['video_prediction'].shape == (2, 3, 3, 32, 32)
        ks = rk(drv[:, :, :1])
        kv = {k: torch.cat([rk(drv[:, :, i:i + 1])[k] for i in range(3)], 1) for k in ks}
        c = ref_r.generate(rg, drv[:, :, :1], ks, kv)
        ksc, kvc = {k: v.cuda() for k, v in ks.items()}, {k: v.cuda() for k, v in kv.items()}
        d = our_r.generate(gen, drv.cuda()[:, :, :1], ksc, kvc)
        e_rec = helpers.max_abs(c['video_prediction'], d['video_prediction'])
        print('reconstruction.py generate: |frame| %.2e' % e_rec)
        assert e_rec < 1e-3
