"""The reference's OWN drivers (train.py / transfer.py / reconstruction.py, imported unchanged through
oracle/ref_shim.load_driver - from /root/reference in the build container, from the byte-compiled oracle/_ref on the
GPU box) with the reference's own modules, on CPU, against the oracle port: pins the port's training iteration
(train.py:110-136 incl. the three Adam steps) and its transfer / reconstruction loops to the real thing."""
import copy

import pytest
import torch

import helpers
from oracle import monkey_oracle as mo, ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason='reference (source tree or oracle/_ref build) absent')


def _ref_and_port(cfg):
    torch.manual_seed(0)
    rg, rd, rk = ref_shim.build_from_config(cfg)
    helpers.perturb_flow_head(rg)
    og, od, ok = mo.build_from_config(cfg)
    og.load_state_dict(rg.state_dict()); od.load_state_dict(rd.state_dict()); ok.load_state_dict(rk.state_dict())
    return (rg, rd, rk), (og, od, ok)


def test_reference_train_loop_matches_port_iterations():
    cfg = helpers.driver_config(num_epochs=2)
    tp = cfg['train_params']
    (rg, rd, rk), (og, od, ok) = _ref_and_port(cfg)
    train_mod = ref_shim.load_driver('train', modules='reference')
    with ref_shim.cpu_data_parallel():
        rec = helpers.run_reference_train(train_mod, (rg, rd, rk), cfg, device_ids=None)
    assert len(rec.iters) == 4 and [e[0] for e in rec.epochs] == [0, 1]
    assert rec.iters[0][1] == ['layer-0_rec', 'layer-1_rec', 'layer-2_rec', 'gen_gan', 'disc_gan'][:len(rec.iters[0][1])] \
        or len(rec.iters[0][1]) == len(rec.iters[0][2])
    # the same four batches through the port's restatement of the loop body
    torch.manual_seed(1234)
    loader = torch.utils.data.DataLoader(helpers.TinyPairs(), batch_size=tp['batch_size'], shuffle=True, num_workers=0,
                                         drop_last=True)
    for m in (og, od, ok):
        m.train()
    opts = mo.make_optimizers(og, od, ok, tp['lr'])
    it = 0
    for epoch in range(2):
        for x in loader:
            g_vals, d_vals = mo.train_iteration(ok, og, od, opts, tp, x)
            port = [float(v) for v in g_vals + d_vals]
            want = rec.iters[it][2]
            tol = 1e-4 if it == 0 else 5e-2   # after an Adam step, rounding-level gradients may flip the sign of lr
            assert len(port) == len(want)
            for a, b in zip(port, want):
                assert abs(a - b) <= tol * max(1.0, abs(b)), (it, port, want)
            it += 1
        if epoch + 1 in tp['epoch_milestones']:
            for o in opts:
                for gp in o.param_groups:
                    gp['lr'] *= 0.1
    assert it == 4


def test_reference_transfer_and_reconstruction_loops_match_port():
    cfg = helpers.driver_config()
    (rg, rd, rk), (og, od, ok) = _ref_and_port(cfg)
    for m in (rg, rk, og, ok):
        m.eval()
    src, drv = helpers.smooth_frames(2, 1, 32, 3), helpers.smooth_frames(2, 3, 32, 4)
    transfer_mod = ref_shim.load_driver('transfer', modules='reference')
    recon_mod = ref_shim.load_driver('reconstruction', modules='reference')
    tparams = {'normalization_params': {'move_location': True, 'movement_mult': False, 'adapt_variance': False,
                                        'clip_mean': False}}
    with torch.no_grad():
        a = transfer_mod.transfer_one(rg, rk, src, drv, tparams)
        b = mo.transfer_one(og, ok, src, drv, tparams['normalization_params'])
        assert helpers.max_abs(a['video_prediction'], b['video_prediction']) < 1e-4
        assert helpers.max_abs(a['kp_norm']['mean'], b['kp_norm']['mean']) < 1e-5
        kp_s = rk(drv[:, :, :1])
        kp_v = {k: torch.cat([rk(drv[:, :, i:i + 1])[k] for i in range(3)], 1) for k in kp_s}
        c = recon_mod.generate(rg, drv[:, :, :1], kp_s, kp_v)
        d = mo.reconstruct(og, ok, drv)
        assert helpers.max_abs(c['video_prediction'], d['video_prediction']) < 1e-4
