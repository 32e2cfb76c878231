"""Tensor-core (tcgen05 / TMA, TF32) convolution: kernel-level parity against the exact fp32 kernel on the same
device buffers, and module-level parity of the whole generator in 'tf32' conv mode against the CPU oracle.
TF32 keeps 10 mantissa bits, so kernel tolerances are relative 2e-3 (of the output's max-abs)."""
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu

CASES = [
    # cin, cout, k, pad, H, W, N, resid, act
    (32, 64, 3, 1, 16, 16, 4, False, 0),
    (64, 128, 3, 1, 32, 32, 2, False, 0),
    (16, 32, 3, 1, 64, 64, 2, False, 0),      # Cin < one 32-channel chunk
    (20, 48, 3, 1, 13, 9, 3, False, 0),       # ragged channels, partial tiles
    (128, 256, 3, 1, 8, 8, 4, False, 0),      # two N tiles, image smaller than the 16-wide tile
    (256, 144, 3, 1, 4, 4, 8, False, 0),      # N tile of 128 + 16, several images per tile
    (512, 512, 3, 1, 2, 2, 32, False, 0),
    (64, 64, 4, 0, 29, 29, 2, False, 0),      # discriminator: 4x4 valid
    (48, 16, 1, 0, 16, 16, 2, False, 2),      # 1x1 + sigmoid
    (48, 48, 3, 1, 16, 16, 2, True, 1),       # residual + relu epilogue
    (64, 32, 3, 3 - 1 - 1, 16, 16, 2, False, 0),
    (32, 64, 4, 3, 10, 10, 2, False, 0),      # dgrad of the 4x4 valid conv: full correlation, pad = 3
]


@pytest.mark.parametrize('cin,cout,k,pad,H,W,N,resid,act', CASES)
def test_conv_tc_matches_fp32_kernel(cin, cout, k, pad, H, W, N, resid, act):
    from monkey_net_b200 import lib
    torch.manual_seed(cin + cout + k)
    dev = torch.device('cuda')
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(N, H, W, cin, device=dev)
    w = torch.randn(cout, cin, 1, k, k, device=dev) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device=dev)
    Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    r = torch.randn(N, Ho, Wo, cout, device=dev) if resid else None
    wp, wt = torch.empty(k * k * cin * cout, device=dev), torch.empty(k * k * cin * cout, device=dev)
    bp = torch.empty(cout, device=dev)
    lib.call('mk_pack_weight', w.data_ptr(), cout, cin, k, k, 1, None, cin, cout, 0, wp.data_ptr(), b.data_ptr(),
             bp.data_ptr(), st)
    lib.call('mk_pack_weight', w.data_ptr(), cout, cin, k, k, 1, None, cin, cout, 2, wt.data_ptr(), None, None, st)
    y0 = torch.empty(N, Ho, Wo, cout, device=dev)
    y1 = torch.full((N, Ho, Wo, cout), float('nan'), device=dev)
    rp = r.data_ptr() if resid else None
    lib.call('mk_conv2d', x.data_ptr(), N, H, W, cin, cin, 0, wp.data_ptr(), k, k, pad, None, bp.data_ptr(), rp,
             cout if resid else 0, act, 0.0, y0.data_ptr(), cout, cout, 0, st)
    lib.call('mk_conv2d_tc', x.data_ptr(), N, H, W, cin, cin, wt.data_ptr(), k, k, pad, None, bp.data_ptr(), rp,
             cout if resid else 0, act, 0.0, y1.data_ptr(), cout, cout, st)
    torch.cuda.synchronize()
    assert not torch.isnan(y1).any(), 'tensor-core kernel left outputs unwritten'
    err = float((y0 - y1).abs().max()) / (float(y0.abs().max()) + 1e-12)
    assert err < 2e-3, err


def test_conv_tc_rejects_unsupported_shapes_without_touching_output():
    from monkey_net_b200 import lib
    dev = torch.device('cuda')
    x = torch.randn(1, 8, 8, 4, device=dev)
    y = torch.zeros(1, 8, 8, 16, device=dev)
    w = torch.zeros(9 * 4 * 16, device=dev)
    with pytest.raises(RuntimeError, match='unsupported'):
        lib.call('mk_conv2d_tc', x.data_ptr(), 1, 8, 8, 4, 4, w.data_ptr(), 3, 3, 1, None, None, None, 0, 0, 0.0,
                 y.data_ptr(), 16, 16, torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize('name,res', [('taichi', 64), ('shapes', 64)])
def test_generator_tf32_mode_against_oracle(name, res):
    """Whole generator + keypoint detector with the tensor-core convs; bar = north-star 1e-3 on the frame."""
    from monkey_net_b200 import ops
    from oracle import monkey_oracle as mo
    import test_gpu_2_modules as t2
    cfg = helpers.load_config(name)
    (gen, disc, kp), (og, od, ok), x = t2._pair(cfg, res, 2, d=1)
    for m in (gen, kp, og, ok):
        m.eval()
    ops.set_conv_mode('tf32')
    try:
        with torch.no_grad():
            a = kp(x['video'].cuda())
            b = ok(x['video'])
            ks = {k: v for k, v in b.items()}
            oa = og(x['source'], kp_driving=b, kp_source=ks)
            bc = {k: v.cuda() for k, v in b.items()}
            ga = gen(x['source'].cuda(), kp_driving=bc, kp_source=bc)
    finally:
        ops.set_conv_mode('fp32')
    e_kp = helpers.max_abs(a['mean'], b['mean'])
    e_pred = helpers.max_abs(ga['video_prediction'], oa['video_prediction'])
    e_def = helpers.max_abs(ga['video_deformed'], oa['video_deformed'])
    print('tf32 mode %s: |kp mean| %.2e  |prediction| %.2e  |deformed| %.2e' % (name, e_kp, e_pred, e_def))
    assert e_kp < 2e-3 and e_pred < 5e-3
