"""Parity at the resolution the numbers are quoted on (BASELINE.json configs[2..4], 256x256), in the arithmetic the bench
times ('auto'), against the CPU oracle loaded with the product's own state_dict:

  * moving-gif.yaml nets @256: the batched, fused, CUDA-graph-replayed transfer_one (what bench.py's transfer_256 times)
    on 2 sources x 2 driving frames - frames <= 1e-3, keypoints <= 2e-5 with identical pixel indices;
  * taichi.yaml nets @256: one train-mode G-step + D-step on 2 frame pairs - loss terms, generated frame, and every
    parameter gradient of the generator / keypoint detector / discriminator (cosine);
  * vox-full.yaml nets @256 (trilinear grid resize, 7-block generator, kp-embedding at scale 0.25): eval forward.

These are the shapes the 256x256 launch plans (halo kernels with resident / streamed weights, RB > 1, split-K of the
deep levels, sub-pixel upsampled convs, multi-CTA wgrad splits) actually run at; the CPU side takes ~10-40 s per test."""
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu


def _pair(name, batch, d=1):
    import test_gpu_2_modules as t2
    return t2._pair(helpers.load_config(name), 256, batch, d=d)


def test_moving_gif_transfer_one_256():
    from monkey_net_b200 import transfer_step
    from oracle import monkey_oracle as mo
    cfg = helpers.load_config('moving-gif')
    (gen, disc, kp), (og, od, ok), x = _pair('moving-gif', 2, d=2)
    with torch.no_grad():  # running statistics as a trained checkpoint would carry them, identical on both sides
        for m in (gen, kp):
            m.train()
        kj = kp(torch.cat([x['source'], x['video'][:, :, :1]], 2).cuda())
        gen(x['source'].cuda(), {k: v[:, 1:] for k, v in kj.items()}, {k: v[:, :1] for k, v in kj.items()})
    og.load_state_dict({k: v.cpu() for k, v in gen.state_dict().items()})
    ok.load_state_dict({k: v.cpu() for k, v in kp.state_dict().items()})
    for m in (gen, kp, og, ok):
        m.eval()
    tparams = cfg['transfer_params']
    runner = transfer_step.GraphedTransfer(gen, kp, tparams, use_graph=True)
    runner.run(x['source'].cuda(), x['video'].cuda())
    got = runner.run(x['source'].cuda(), x['video'].cuda())          # the graph replay is what gets compared
    with torch.no_grad():
        want = mo.transfer_one(og, ok, x['source'], x['video'], tparams['normalization_params'])
    e_frame = helpers.max_abs(got['video_prediction'], want['video_prediction'])
    e_def = helpers.max_abs(got['video_deformed'], want['video_deformed'])
    e_kp = helpers.max_abs(got['kp_driving']['mean'], want['kp_driving']['mean'])
    px = lambda m: torch.round(256 * (m.cpu() + 1) / 2)
    same = torch.equal(px(got['kp_driving']['mean']), px(want['kp_driving']['mean']))
    print('moving-gif@256 transfer_one (graph replay): |frame| %.2e |deformed| %.2e |kp| %.2e identical pixel indices %s'
          % (e_frame, e_def, e_kp, same))
    assert runner.graph is not None and got['video_prediction'].shape == (2, 3, 2, 256, 256)
    assert e_frame < 1e-3 and e_def < 1e-3 and e_kp < 2e-5 and same


def test_taichi_train_step_256():
    from oracle import monkey_oracle as mo
    import train_glue
    cfg = helpers.load_config('taichi')
    tp = cfg['train_params']
    (gen, disc, kp), (og, od, ok), x = _pair('taichi', 2)
    for m in (gen, disc, kp, og, od, ok):
        m.train()
    out = mo.generator_full(ok, og, od, tp, x)
    sum(v.mean() for v in out[:-2]).backward()
    xg = {k: v.cuda() for k, v in x.items()}
    pout = train_glue.generator_full(kp, gen, disc, tp, xg)
    sum(v.mean() for v in pout[:-2]).backward()
    e_loss = max(helpers.rel_err(a, b) for a, b in zip(pout[:-2], out[:-2]))
    e_frame = helpers.max_abs(pout[-2]['video_prediction'], out[-2]['video_prediction'])
    e_kp = helpers.max_abs(pout[-1]['mean'], out[-1]['mean'])
    coss = []
    for (n1, p1), (n2, p2) in zip(list(gen.named_parameters()) + list(kp.named_parameters()),
                                  list(og.named_parameters()) + list(ok.named_parameters())):
        if p2.grad is None or helpers.structurally_zero_grad(n1):
            continue
        a, b = p1.grad.detach().cpu().flatten(), p2.grad.flatten()
        coss.append((float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)), n1))
    coss.sort()
    med, p10 = coss[len(coss) // 2][0], coss[len(coss) // 10][0]
    e_mean = float((pout[-2]['video_prediction'].detach().cpu() - out[-2]['video_prediction'].detach()).abs().mean())
    print('taichi@256 G-step: loss terms rel %.2e |frame| max %.2e mean %.2e |kp| %.2e; gradient cosine median %.6f, '
          '10th pct %.6f, worst %s' % (e_loss, e_frame, e_mean, e_kp, med, p10, coss[:3]))
    # This configuration (default init, TRAIN-mode batch norm over 2 samples, 5-block hourglasses at 256x256) is
    # ill-conditioned in the reference itself: the EXACT fp32 FFMA kernels sit at frame max 3.5e-2 / mean 1.4e-4,
    # kp 1.2e-5, gradient cosine median 0.9989 / 10th percentile 0.995 against the oracle, and two exact runs agree
    # bit for bit (tools/diag_train256.py -> profiles/r2_diag_train256.txt; 3xTF32: 1.2e-1 / 3.3e-4, 3.8e-5, 0.9985 /
    # 0.926).  Rounding differences of 1e-7 are amplified ~1000x by the normalisation of near-constant channels and
    # flip isolated pixels of the warp.  The per-sample LOSS TERMS (what training consumes) agree to 5e-5; the bars
    # below are that envelope with margin - kernel-level parity at these shapes is pinned in tests/test_gpu_3_tc.py.
    assert e_loss < 1e-3 and e_mean < 2e-3 and e_kp < 2e-4
    assert med >= 0.995 and p10 >= 0.85 and coss[0][0] > 0.5
    # discriminator step on the same pair of generated frames
    for m in (gen, disc, kp, og, od, ok):
        m.zero_grad()
    dl = mo.discriminator_full(ok, og, od, tp, x, out[-1], out[-2])
    sum(v.mean() for v in dl).backward()
    pdl = train_glue.discriminator_full(kp, gen, disc, tp, xg, pout[-1], pout[-2])
    sum(v.mean() for v in pdl).backward()
    assert helpers.rel_err(pdl[0], dl[0]) < 2e-3
    for (n1, p1), (n2, p2) in zip(disc.named_parameters(), od.named_parameters()):
        if helpers.structurally_zero_grad(n1):
            continue
        a, b = p1.grad.detach().cpu().flatten(), p2.grad.flatten()
        c = float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))
        assert c > 0.99, (n1, c)


def test_vox_full_eval_256():
    (gen, disc, kp), (og, od, ok), x = _pair('vox-full', 1, d=2)
    for m in (gen, kp, og, ok):
        m.eval()
    with torch.no_grad():
        a = kp(x['video'].cuda())
        b = ok(x['video'])
        ks = ok(x['source'])
        want = og(x['source'], kp_driving=b, kp_source=ks)
        got = gen(x['source'].cuda(), kp_driving={k: v.cuda() for k, v in b.items()},
                  kp_source={k: v.cuda() for k, v in ks.items()})
    e_kp = helpers.max_abs(a['mean'], b['mean'])
    e_frame = helpers.max_abs(got['video_prediction'], want['video_prediction'])
    e_def = helpers.max_abs(got['video_deformed'], want['video_deformed'])
    px = lambda m: torch.round(256 * (m.cpu() + 1) / 2)
    print('vox-full@256 eval: |kp| %.2e |frame| %.2e |deformed| %.2e' % (e_kp, e_frame, e_def))
    assert e_kp < 2e-5 and torch.equal(px(a['mean']), px(b['mean']))
    assert e_frame < 1e-3 and e_def < 1e-3 and got['video_prediction'].shape == (1, 3, 2, 256, 256)
