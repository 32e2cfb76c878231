"""Host-side launch planning of the tensor-core kernels (mk_conv2d_tc_plan / mk_conv2d_wgrad_tc_plan: the same code the
real entry points run, as a dry run that needs no GPU) swept over EVERY convolution of the eight shipped configurations
at the resolutions / batch sizes of BASELINE.json - forward, input-gradient and weight-gradient shapes, with the padded
channel counts the NHWC layout produces.  Checks the invariants the kernels rely on: shared memory within 227 KB, ring
depths, TMEM columns, TMA box limits, split coverage without empty CTAs."""
import ctypes

import pytest
import torch

import helpers
from oracle import monkey_oracle as mo

CASES = [('shapes', 64, 32), ('taichi', 64, 32), ('taichi', 256, 8), ('moving-gif', 256, 16), ('vox-full', 256, 16),
         ('vox', 256, 4), ('nemo', 64, 32), ('bair', 64, 32), ('actions', 64, 32)]


def pad4(c):
    return (c + 3) & ~3


def conv_shapes(name, res):
    """(cin, cout, k, pad, H, W, frames_per_sample, upsampled) of every conv of KP detector (D=2), generator and
    discriminator, from forward hooks on the oracle."""
    cfg = helpers.load_config(name)
    og, od, ok = mo.build_from_config(cfg)
    for m in (og, od, ok):
        for p in m.parameters():
            torch.nn.init.normal_(p, std=0.01)
        m.eval()
    shapes = []

    def hook(m, inp, out):
        x = inp[0]
        co, cig, _, kh, kw = m.weight.shape
        groups = x.shape[1] // cig
        shapes.append((x.shape[1], co, kh, (kh - 1) // 2 if out.shape[-1] == x.shape[-1] else 0, x.shape[-2],
                       x.shape[-1], x.shape[0] * x.shape[2], groups))
    hs = [m.register_forward_hook(hook) for mod in (og, od, ok) for m in mod.modules() if isinstance(m, mo._Conv)]
    x = torch.rand(1, 3, 1, res, res)
    with torch.no_grad():
        kpj = ok(torch.cat([x, x], 2))
        kd, ks = {k: v[:, 1:] for k, v in kpj.items()}, {k: v[:, :1] for k, v in kpj.items()}
        g = og(x, kd, ks)
        od(g['video_prediction'], kd, ks)
    for h in hs:
        h.remove()
    return shapes


def plan(lib, fn, *args):
    out = (ctypes.c_int * 16)()
    rc = getattr(lib, fn)(*args, out)
    return rc, list(out)


@pytest.mark.parametrize('name,res,batch', CASES)
def test_every_layer_has_a_valid_tensor_core_plan(name, res, batch):
    from monkey_net_b200 import lib as mklib
    lib = mklib.load()
    n_layers = 0
    for cin, cout, k, pad, H, W, frames, groups in conv_shapes(name, res):
        N = frames * batch
        for cin_p in {pad4(cin), pad4(cin) + 4}:          # dense and concat-with-holes layouts
            cop = pad4(cout)
            ho, wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
            # forward (act 0 and a fused activation), input gradient (full correlation), weight gradient
            jobs = [('fwd', (N, H, W, cin_p, 0, k, k, pad, 0, cop, cop)), ('fwd_act', (N, H, W, cin_p, 0, k, k, pad, 1, cop, cop)),
                    ('dgrad', (N, ho, wo, cop, 0, k, k, k - 1 - pad, 0, cin_p, cin_p))]
            if k == 3 and groups == 1 and H % 2 == 0 and W % 2 == 0:
                # the same conv as the sub-pixel form of UpBlock3D (nearest x2 folded in): low-resolution input
                jobs.append(('ups', (N, H // 2, W // 2, cin_p, 1, 3, 3, 1, 0, cop, cop)))
            for tag, a in jobs:
                rc, o = plan(lib, 'mk_conv2d_tc_plan', *a)
                assert rc == 0, (name, tag, a, lib.mk_last_error())
                gx, gy, gz, smem, nst, ksplit, ips, niter, tmem, tw, th, tn, b_rows, stage, oh, ow = o
                assert smem <= 227 * 1024 and 2 <= nst <= 8 and nst * stage + 1280 == smem
                assert tmem in (32, 64, 128) and b_rows <= tmem and tw * th * tn == 128 and max(tw, th, tn) <= 256
                assert ksplit >= 1 and ips * ksplit >= niter and ips * (ksplit - 1) < niter      # no empty split
                assert gx < 2 ** 31 and gy <= 65535 and gz <= 65535
                if tag == 'fwd_act':
                    assert ksplit == 1                                                            # linear epilogue only
            # tf32x3 plans of the per-tap kernels (twice the stage bytes)
            for tag, a in jobs[:1] + jobs[2:]:
                a3 = list(a)
                a3[4] |= 2
                rc, o = plan(lib, 'mk_conv2d_tc_plan', *a3)
                assert rc == 0 and o[3] <= 227 * 1024 and 2 <= o[4] <= 8, (name, tag, 'x3', a3, lib.mk_last_error())
            rc, o = plan(lib, 'mk_conv2d_wgrad_tc_plan', N, H, W, cin_p, cop, k, k, pad | 256)
            assert rc == 0 and o[3] <= 227 * 1024 and o[6] * o[7] <= 512, (name, 'wgrad x3', lib.mk_last_error())
            # halo-window persistent kernels: either a valid plan or a clean decline (-2, few-tile / small-image layers)
            for x3 in (0, 1):
                for tag, (n_, h_, w_, ci_, k_, p_, co_, res_) in (
                        ('halo fwd', (N, H, W, cin_p, k, pad, cop, 0)), ('halo fwd+res', (N, H, W, cin_p, k, pad, cop, 1)),
                        ('halo dgrad', (N, ho, wo, cop, k, k - 1 - pad, cin_p, 0))):
                    rc, o = plan(lib, 'mk_conv2d_tc_halo_plan', n_, h_, w_, ci_, k_, k_, p_, co_, res_, x3)
                    assert rc in (0, -2), (name, tag, lib.mk_last_error())
                    if rc == 0:
                        gx, gy, rb, smem, a_st, b_sl, resident, tmem, ntiles, halo_rows, a_stage, b_slot, twv, ngr, npad, fl = o
                        assert k_ in (1, 3, 4) and smem <= 227 * 1024 and 2 <= a_st <= 4 and rb in (1, 2, 4)
                        assert 2 * rb * npad <= tmem <= 512 and halo_rows == 8 * rb + k_ and twv == 17 - k_
                        ct = (fl >> 3) & 1     # column taps stacked on N: one weight slot per tap ROW, RB = 1
                        assert not ct or (rb == 1 and k_ > 1 and co_ % 16 == 0 and npad == k_ * co_ <= 256)
                        nslots = (k_ if ct else k_ * k_) * ((ci_ + 31) // 32)
                        assert (resident and b_sl == nslots <= 40) or (not resident and 2 <= b_sl <= 8)
                        assert gx * gy <= 148 and ntiles >= gx and (fl & 1) == x3 and ((fl >> 1) & 3) in (1, 2, 3)
                        nstg = (fl >> 1) & 3
                        assert a_st * a_stage + b_sl * b_slot + nstg * (1 + res_) * twv * 8 * 128 + 3072 == smem
                if k == 3 and groups == 1 and H % 2 == 0 and W % 2 == 0:
                    # the upsampled conv on the halo kernels: four sub-pixel passes on the low-resolution grid (forward:
                    # 2x2 halo conv + 5-D store map; weight gradient: 2x2 taps, dY through a 5-D load map)
                    rc, o = plan(lib, 'mk_conv2d_tc_halo_ups_plan', N, H // 2, W // 2, cin_p, cop, x3)
                    assert rc in (0, -2), (name, 'halo ups', lib.mk_last_error())
                    if rc == 0:
                        gx, gy, rb, smem, a_st, b_sl, resident, tmem, ntiles, halo_rows, a_stage, b_slot, twv, ngr, npad, fl = o
                        assert smem <= 227 * 1024 and 2 <= a_st <= 4 and rb in (1, 2, 4) and (H // 2) % (8 * rb) == 0
                        assert halo_rows == 8 * rb + 2 and twv == 15 and 2 * rb * npad <= tmem <= 512
                        assert gx * gy <= 148 and ntiles >= gx and (fl & 1) == x3
                        nstg = (fl >> 1) & 3
                        assert a_st * a_stage + b_sl * b_slot + nstg * twv * 8 * 128 + 3072 == smem
                    rc, o = plan(lib, 'mk_conv2d_wgrad_halo_ups_plan', N, H // 2, W // 2, cin_p, cop, x3)
                    assert rc in (0, -2), (name, 'wgrad halo ups', lib.mk_last_error())
                    if rc == 0:
                        cot, cig, spl, smem, stages, tr, nci, nco, tmem, ntiles, tps, co_pad, stage, f3, twv, tilesh = o
                        assert smem <= 227 * 1024 and 2 <= stages <= 4 and tr in (4, 8) and (H // 2) % tr == 0 and twv == 15
                        assert nci * 2 * co_pad <= tmem <= 512 and cig * nci >= (cin_p + 31) // 32
                        assert tps * spl >= ntiles and tps * (spl - 1) < ntiles and spl <= 65535
                rc, o = plan(lib, 'mk_conv2d_wgrad_halo_plan', N, H, W, cin_p, cop, k, k, pad, x3)
                assert rc in (0, -2), (name, 'wgrad halo', lib.mk_last_error())
                if rc == 0:
                    cot, cig, spl, smem, stages, tr, nci, nco, tmem, ntiles, tps, co_pad, stage, f3, twv, tilesh = o
                    assert k in (3, 4) and smem <= 227 * 1024 and 2 <= stages <= 4 and tr in (4, 8)
                    assert nci * k * co_pad <= tmem <= 512 and cig * nci >= (cin_p + 31) // 32 and nco * 32 >= min(cop, 128)
                    assert tps * spl >= ntiles and tps * (spl - 1) < ntiles and spl <= 65535
                    assert stage == ((nci * (tr + 4) + nco * tr) * 2048) << x3 and stages * stage + 2048 == smem
            rc, o = plan(lib, 'mk_conv2d_wgrad_tc_plan', N, H, W, cin_p, cop, k, k, pad)
            assert rc == 0, (name, 'wgrad', lib.mk_last_error())
            gx, gy, gz, smem, a_slots, b_slots, taps, npad, tmem, tw, th, tn, nchunks, cps, na, nb = o
            assert smem <= 227 * 1024 and 1 <= a_slots <= 4 and 2 <= b_slots <= 12
            assert taps * npad <= 512 and tmem in (32, 64, 128, 256, 512) and taps * npad <= tmem
            assert tw * th * tn == 64 and cps * gz >= nchunks and cps * (gz - 1) < nchunks
            assert gy * taps >= k * k and gz <= 65535
            n_layers += 1
    assert n_layers > 40


def test_wgrad_pixel_splits_fill_whole_waves():
    """tc_common.cuh:pick_splits - the weight-gradient kernels split their pixel range so that the CTAs form whole waves of
    the 148 SMs: ceil(148 / groups) splits used to overshoot one wave (16 groups x 10 = 160 CTAs = two waves of 64 tiles
    where one wave of 72 does; measured 795 -> 447 us on 512->128 @64x64 x 16)."""
    from monkey_net_b200 import lib as mklib
    lib = mklib.load()
    for N, H, W, ci, co, k, pad in [(16, 64, 64, 512, 128, 3, 1), (8, 61, 61, 128, 256, 4, 0), (16, 32, 32, 256, 512, 3, 1),
                                    (8, 256, 256, 48, 48, 3, 1), (16, 256, 256, 128, 32, 3, 1)]:
        rc, o = plan(lib, 'mk_conv2d_wgrad_halo_plan', N, H, W, ci, co, k, k, pad, 1)
        assert rc == 0
        cot, cig, spl, _, _, _, _, _, _, ntiles, tps = o[:11]
        ctas = cot * cig * spl
        waves = -(-ctas // 148)
        # no split count with fewer waves x tiles-per-CTA exists among the neighbours (+ the per-CTA epilogue of ~6 tiles)
        cost = waves * (tps + 6)
        for alt in range(1, min(ntiles, 4 * 148 // (cot * cig) + 1) + 1):
            a_tps = -(-ntiles // alt)
            a_eff = -(-ntiles // a_tps)
            a_cost = -(-(cot * cig * a_eff) // 148) * (a_tps + 6)
            assert cost <= a_cost * 1.001, (N, H, W, ci, co, spl, alt)
        assert tps * spl >= ntiles and tps * (spl - 1) < ntiles


def test_column_taps_on_n_is_chosen_where_the_mma_phase_dominates():
    """conv_halo.cu:halo_wants_ct through the planner's flag (bit 3 of out[15]): narrow layers with many K steps stack their
    column taps on N, layers whose epilogue dominates (few input channels, or many output groups) keep the plain schedule,
    and a CT plan that does not fit shared memory falls back instead of failing."""
    from monkey_net_b200 import lib as mklib
    lib = mklib.load()

    def ct(N, H, W, ci, k, pad, co, res, x3):
        rc, o = plan(lib, 'mk_conv2d_tc_halo_plan', N, H, W, ci, k, k, pad, co, res, x3)
        assert rc == 0, lib.mk_last_error()
        return (o[15] >> 3) & 1, o
    assert ct(8, 256, 256, 128, 3, 1, 32, 0, 0)[0] == 1      # 128->32: 16 K steps per tap
    assert ct(8, 253, 253, 64, 4, 3, 16, 0, 0)[0] == 1       # 64->16 4x4 (dgrad of the discriminator's first block)
    assert ct(8, 256, 256, 48, 3, 1, 48, 0, 1)[0] == 1       # 48->48 in reference precision (2 MMAs per K step)
    assert ct(16, 256, 256, 4, 3, 1, 64, 0, 0)[0] == 0       # image input: one K step, two output groups
    assert ct(8, 256, 256, 32, 3, 1, 128, 0, 0)[0] == 0      # S * Cout > 256: not eligible
    assert ct(8, 256, 256, 36, 3, 1, 12, 0, 0)[0] == 0       # Cout_p % 16 != 0: TMEM column groups would be unaligned
    flag, o = ct(8, 256, 256, 16, 4, 0, 64, 0, 1)            # N = 256 in reference precision: no shared-memory plan -> plain
    assert flag == 0 and o[2] in (1, 2, 4)
