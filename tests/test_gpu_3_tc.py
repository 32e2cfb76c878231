"""Tensor-core (tcgen05 / TMA) convolution kernels - 1xTF32, reference precision (tf32x3: TF32 main term + BF16 cross terms) - and the exact FFMA kernels: kernel-level parity
against torch's F.conv2d / autograd in double precision on the same device buffers, and module-level parity of the
whole generator in the tensor-core modes against the CPU oracle."""
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
    (4, 32, 3, 1, 32, 32, 2, False, 0),       # image input: 4 physical channels ride on the TMA zero fill
    (24, 24, 3, 1, 16, 16, 2, True, 0),       # refinement ResBlock of shapes.yaml: 23 -> 24 padded channels
    (36, 8, 3, 1, 16, 16, 2, False, 0),       # hourglass head: Cout_p = 8
    (256, 4, 1, 0, 5, 5, 4, False, 0),        # discriminator score conv: Cout_p = 4
    (132, 128, 3, 1, 8, 8, 2, False, 0),      # decoder concat: Cin_p = 132
    (128, 128, 3, 1, 2, 2, 32, True, 0),      # one tile, 36 K iterations: split-K with residual
    (24, 24, 3, 1, 64, 64, 32, True, 0),      # 1024 tiles, 3 resident CTAs per SM
    (1024, 1024, 3, 1, 4, 4, 8, False, 0),    # taichi depth: 8 cout tiles x split-K
]


# Every kernel is compared with torch's own convolution evaluated in DOUBLE precision on the same device buffers
# (F.conv2d + autograd, the op the reference calls) - not with another kernel of this repository.  Tolerances are
# relative to the output's max-abs: TF32 keeps 10 mantissa bits (2e-3), tf32x3 and the FFMA kernel are fp32-accurate.
KERNEL_TOL = {'tf32': 2e-3, 'tf32x3': 2e-5, 'ffma': 1e-5}


def _torch_conv(x, w, b, pad, r, act, ups=False):
    """NHWC fp32 inputs -> NHWC double reference of act(conv2d(x) + b + r)"""
    xd = x.double().permute(0, 3, 1, 2)
    if ups:
        xd = torch.nn.functional.interpolate(xd, scale_factor=2, mode='nearest')
    y = torch.nn.functional.conv2d(xd, w[:, :, 0].double(), b.double() if b is not None else None, padding=pad)
    y = y.permute(0, 2, 3, 1)
    if r is not None:
        y = y + r.double()
    if act == 1:
        y = torch.relu(y)
    elif act == 2:
        y = torch.sigmoid(y)
    return y


def _run_conv_kernel(kernel, x, w, b, k, pad, r, act, ups=0):
    from monkey_net_b200 import lib
    dev = x.device
    st = torch.cuda.current_stream().cuda_stream
    N, H, W, cin = x.shape
    cout = w.shape[0]
    Hl, Wl = H << ups, W << ups
    Ho, Wo = Hl + 2 * pad - k + 1, Wl + 2 * pad - k + 1
    x3 = kernel == 'tf32x3'
    taps = (16 if ups else k * k) if kernel != 'ffma' else k * k
    wp = torch.empty(taps * cout * (cin + (((cin + 7) & ~7) if x3 else 0)), device=dev)
    bp = torch.empty(cout, device=dev)
    mode = 0 if kernel == 'ffma' else ((4 if ups else 2) | (8 if x3 else 0))
    lib.call('mk_pack_weight', w.data_ptr(), cout, cin, k, k, 1, None, cin, cout, mode, wp.data_ptr(),
             b.data_ptr() if b is not None else None, bp.data_ptr() if b is not None else None, st)
    y = torch.full((N, Ho, Wo, cout), float('nan'), device=dev)
    rp = r.data_ptr() if r is not None else None
    bpp = bp.data_ptr() if b is not None else None
    if kernel == 'ffma':
        lib.call('mk_conv2d', x.data_ptr(), N, H, W, cin, cin, ups, wp.data_ptr(), k, k, pad, None, bpp, rp,
                 cout if r is not None else 0, act, 0.0, y.data_ptr(), cout, cout, 0, st)
    else:
        lib.call('mk_conv2d_tc_x3' if x3 else 'mk_conv2d_tc', x.data_ptr(), N, H, W, cin, cin, ups, wp.data_ptr(), k, k,
                 pad, None, bpp, rp, cout if r is not None else 0, act, 0.0, y.data_ptr(), cout, cout, st)
    torch.cuda.synchronize()
    return y


@pytest.mark.parametrize('kernel', ['tf32', 'tf32x3', 'ffma'])
@pytest.mark.parametrize('cin,cout,k,pad,H,W,N,resid,act', CASES)
def test_conv_kernels_match_torch_conv2d(cin, cout, k, pad, H, W, N, resid, act, kernel):
    torch.manual_seed(cin + cout + k)
    dev = torch.device('cuda')
    x = torch.randn(N, H, W, cin, device=dev)
    w = torch.randn(cout, cin, 1, k, k, device=dev) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device=dev)
    Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    r = torch.randn(N, Ho, Wo, cout, device=dev) if resid else None
    y = _run_conv_kernel(kernel, x, w, b, k, pad, r, act)
    assert not torch.isnan(y).any(), 'kernel left outputs unwritten'
    ref = _torch_conv(x, w, b, pad, r, act)
    err = float((y.double() - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
    assert err < KERNEL_TOL[kernel], err


@pytest.mark.parametrize('kernel', ['tf32', 'tf32x3', 'ffma'])
@pytest.mark.parametrize('cin,cout,H,W,N', [(64, 32, 16, 16, 2), (128, 64, 4, 4, 4), (40, 16, 9, 5, 3), (256, 128, 2, 2, 8)])
def test_conv_upsampled_matches_torch(cin, cout, H, W, N, kernel):
    """conv3x3(nearest_x2(x)) (util.py:84-85): four 2x2 sub-pixel convs on tensor cores / on-the-fly upsample in the
    FFMA kernel == F.interpolate + F.conv2d."""
    torch.manual_seed(cin + H)
    dev = torch.device('cuda')
    x = torch.randn(N, H, W, cin, device=dev)
    w = torch.randn(cout, cin, 1, 3, 3, device=dev) / (cin * 9) ** 0.5
    b = torch.randn(cout, device=dev)
    y = _run_conv_kernel(kernel, x, w, b, 3, 1, None, 0, ups=1)
    assert not torch.isnan(y).any()
    ref = _torch_conv(x, w, b, 1, None, 0, ups=True)
    err = float((y.double() - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
    # the sub-pixel pack pre-sums up to four taps before the TF32 rounding: same order of error
    assert err < KERNEL_TOL[kernel] * (2 if kernel == 'tf32x3' else 1), err


@pytest.mark.parametrize('x3', [False, True])
@pytest.mark.parametrize('cin,cout,H,W,N,act', [(128, 32, 64, 64, 8, 0),     # Cout 32: column taps on N (2 x 32)
                                                (140, 32, 64, 64, 8, 1),     # 5 channel chunks, leaky epilogue
                                                (64, 64, 32, 48, 16, 0),     # non-square, 3 + 1 column tiles
                                                (36, 12, 64, 64, 8, 2),      # Cout_p = 12, sigmoid
                                                (256, 160, 32, 32, 16, 0)])  # two cout tiles, streamed weights
def test_conv_upsampled_halo_matches_torch(cin, cout, H, W, N, act, x3):
    """mk_conv2d_tc_halo_ups: conv3x3(nearest_x2(x)) as four sub-pixel 2x2 halo-window passes (5-D TMA store of each
    output parity) == F.interpolate + F.conv2d in double."""
    from monkey_net_b200 import lib
    torch.manual_seed(cin + cout)
    dev = torch.device('cuda')
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(N, H, W, cin, device=dev)
    w = torch.randn(cout, cin, 1, 3, 3, device=dev) / (cin * 9) ** 0.5
    b = torch.randn(cout, device=dev)
    wp = torch.empty(16 * cout * (cin + (((cin + 7) & ~7) if x3 else 0)), device=dev)
    bp = torch.empty(cout, device=dev)
    lib.call('mk_pack_weight', w.data_ptr(), cout, cin, 3, 3, 1, None, cin, cout, 4 | (8 if x3 else 0), wp.data_ptr(),
             b.data_ptr(), bp.data_ptr(), st)
    y = torch.full((N, 2 * H, 2 * W, cout), float('nan'), device=dev)
    lib.call('mk_conv2d_tc_halo_ups_x3' if x3 else 'mk_conv2d_tc_halo_ups', x.data_ptr(), N, H, W, cin, cin, wp.data_ptr(),
             None, bp.data_ptr(), act, 0.2, y.data_ptr(), cout, cout, st)
    torch.cuda.synchronize()
    assert not torch.isnan(y).any(), 'sub-pixel halo passes left outputs unwritten'
    ref = _torch_conv(x, w, b, 1, None, act, ups=True) if act != 1 else None
    if act == 1:   # leaky slope 0.2
        r0 = _torch_conv(x, w, b, 1, None, 0, ups=True)
        ref = torch.where(r0 > 0, r0, 0.2 * r0)
    err = float((y.double() - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
    assert err < KERNEL_TOL['tf32x3' if x3 else 'tf32'] * (2 if x3 else 1), err


def test_conv_upsampled_halo_declines_ragged_heights():
    """(n, h) are one merged dimension of the sub-pixel store map: heights that are not multiples of 8 are refused (-2)
    before anything is launched"""
    from monkey_net_b200 import lib
    dev = torch.device('cuda')
    x = torch.randn(8, 60, 64, 64, device=dev)
    y = torch.zeros(8, 120, 128, 32, device=dev)
    wp = torch.zeros(16 * 32 * 64, device=dev)
    rc = lib.call_soft('mk_conv2d_tc_halo_ups', (-2,), x.data_ptr(), 8, 60, 64, 64, 64, wp.data_ptr(), None, None, 0, 0.0,
                       y.data_ptr(), 32, 32, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rc == -2 and float(y.abs().max()) == 0.0


@pytest.mark.parametrize('x3', [False, True])
@pytest.mark.parametrize('cin,cout,H,W,N', [(128, 32, 64, 64, 4), (140, 32, 32, 32, 4), (64, 64, 32, 48, 4),
                                             (36, 12, 64, 64, 2), (256, 160, 32, 32, 4), (20, 48, 32, 32, 2)])
def test_wgrad_upsampled_halo_matches_torch_autograd(cin, cout, H, W, N, x3):
    """mk_conv2d_wgrad_halo_ups + mk_unpack_wgrad_ups: d/dw of conv3x3(nearest_x2(x)) taken on the low-resolution grid
    (four sub-pixel passes, dY through a 5-D TMA map, adjoint of the sub-pixel pack) == torch autograd in double."""
    from monkey_net_b200 import lib
    torch.manual_seed(cin * 3 + cout)
    dev = torch.device('cuda')
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(N, H, W, cin, device=dev)
    dy = torch.randn(N, 2 * H, 2 * W, cout, device=dev)
    dwp = torch.full((16 * cin * cout,), float('nan'), device=dev)
    rc = lib.call_soft('mk_conv2d_wgrad_halo_ups_x3' if x3 else 'mk_conv2d_wgrad_halo_ups', (-2,), x.data_ptr(), N, H, W,
                       cin, cin, dy.data_ptr(), cout, cout, dwp.data_ptr(), st)
    assert rc == 0, 'declined'
    dw = torch.empty(cout, cin, 1, 3, 3, device=dev)
    lib.call('mk_unpack_wgrad_ups', dwp.data_ptr(), cout, cin, None, cin, cout, dw.data_ptr(), 0, st)
    torch.cuda.synchronize()
    wd = torch.zeros(cout, cin, 3, 3, dtype=torch.double, device=dev, requires_grad=True)
    xu = torch.nn.functional.interpolate(x.double().permute(0, 3, 1, 2), scale_factor=2, mode='nearest')
    out = torch.nn.functional.conv2d(xu, wd, padding=1)
    gw, = torch.autograd.grad(out, wd, dy.double().permute(0, 3, 1, 2))
    err = float((dw[:, :, 0].double() - gw).abs().max()) / (float(gw.abs().max()) + 1e-12)
    assert err < KERNEL_TOL['tf32x3' if x3 else 'tf32'] * (2 if x3 else 1), err
    # accumulate flag: += into the destination
    dw2 = dw.clone()
    lib.call('mk_unpack_wgrad_ups', dwp.data_ptr(), cout, cin, None, cin, cout, dw2.data_ptr(), 1, st)
    torch.cuda.synchronize()
    assert torch.allclose(dw2, 2 * dw, rtol=1e-6, atol=0)


WGRAD_CASES = [(32, 64, 3, 1, 16, 16, 4), (64, 128, 3, 1, 32, 32, 2),
               (16, 32, 3, 1, 64, 64, 2), (48, 144, 3, 1, 13, 9, 3),
               (256, 256, 3, 1, 4, 4, 8), (512, 128, 3, 1, 2, 2, 32),
               (64, 128, 4, 0, 29, 29, 2), (48, 16, 1, 0, 16, 16, 2),
               (160, 32, 3, 1, 8, 8, 4), (4, 32, 3, 1, 32, 32, 2),
               (24, 24, 3, 1, 16, 16, 2), (132, 128, 3, 1, 8, 8, 2),
               (36, 8, 3, 1, 16, 16, 2),
               (24, 24, 3, 1, 64, 64, 8),     # many chunks per CTA: both rings wrap
               (4, 32, 4, 0, 64, 64, 2),      # discriminator block 0: 16 taps x 16 cols
               (32, 64, 4, 0, 31, 31, 2),     # 16 taps x 32 cols = all 512 TMEM columns
               (128, 128, 3, 1, 2, 2, 3),     # TMA box larger than the batch (TN > N)
               (300, 40, 1, 0, 7, 7, 5)]      # 1x1, three ci tiles (128+128+44)


@pytest.mark.parametrize('kernel', ['tf32', 'tf32x3', 'ffma'])
@pytest.mark.parametrize('cin,cout,k,pad,H,W,N', WGRAD_CASES)
def test_wgrad_kernels_match_torch_autograd(cin, cout, k, pad, H, W, N, kernel):
    from monkey_net_b200 import lib
    torch.manual_seed(cin + cout)
    dev = torch.device('cuda')
    st = torch.cuda.current_stream().cuda_stream
    Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    x = torch.randn(N, H, W, cin, device=dev)
    dy = torch.randn(N, Ho, Wo, cout, device=dev)
    d = torch.full((k * k * cin * cout,), float('nan'), device=dev)
    if kernel == 'ffma':
        lib.call('mk_conv2d_wgrad', x.data_ptr(), N, H, W, cin, cin, 0, dy.data_ptr(), cout, cout, k, k, pad, d.data_ptr(), st)
    else:
        lib.call('mk_conv2d_wgrad_tc_x3' if kernel == 'tf32x3' else 'mk_conv2d_wgrad_tc', x.data_ptr(), N, H, W, cin, cin,
                 dy.data_ptr(), cout, cout, k, k, pad, d.data_ptr(), st)
    torch.cuda.synchronize()
    assert not torch.isnan(d).any()
    wz = torch.zeros(cout, cin, k, k, device=dev, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), wz, None, padding=pad)
    (gw,) = torch.autograd.grad(y, wz, dy.double().permute(0, 3, 1, 2))
    ref = gw.permute(2, 3, 1, 0).reshape(-1)          # (Co,Ci,R,S) -> [tap][Ci][Co]
    err = float((d.double() - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
    assert err < KERNEL_TOL[kernel] * (2 if kernel == 'tf32x3' else 1), err


def _grad_cosines(cfg, res, batch, mode):
    from monkey_net_b200 import ops, train_step
    from oracle import monkey_oracle as mo
    import test_gpu_2_modules as t2
    (gen, disc, kp), (og, od, ok), x = t2._pair(cfg, res, batch)
    tp = cfg['train_params']
    for m in (gen, disc, kp, og, od, ok):
        m.train()
    out = mo.generator_full(ok, og, od, tp, x)
    sum(v.mean() for v in out[:-2]).backward()
    prev = ops.CONV_MODE
    ops.set_conv_mode(mode)
    try:
        pout = train_step.GeneratorFullModel(kp, gen, disc, tp)({k: v.cuda() for k, v in x.items()})
        sum(v.mean() for v in pout[:-2]).backward()
    finally:
        ops.set_conv_mode(prev)
    coss = []
    for (n1, p1), (n2, p2) in zip(list(gen.named_parameters()) + list(kp.named_parameters()),
                                  list(og.named_parameters()) + list(ok.named_parameters())):
        if p2.grad is None or helpers.structurally_zero_grad(n1):
            continue
        a, b = p1.grad.detach().cpu().flatten(), p2.grad.flatten()
        coss.append((float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)), n1))
    coss.sort()
    return coss


def test_train_step_tf32_mode_gradients():
    """EXPLICIT 'tf32' mode (every conv 1xTF32; not the training default): G-step gradients vs the fp32 oracle."""
    coss = _grad_cosines(helpers.load_config('shapes'), 64, 2, 'tf32')
    med = coss[len(coss) // 2][0]
    print('tf32 train step: gradient cosine vs fp32 oracle: median %.5f, 5 worst %s' % (med, coss[:5]))
    # TF32 (10-bit mantissa) through ~25 conv+BN layers and the warp's d(grid).  The fp32 reference algorithm itself
    # turns 5e-4 relative conv-output noise into median-cosine 0.97 / worst 0.90 gradients (tools/noise_sensitivity.py),
    # so that is the envelope a 1xTF32 implementation can be held to - the reason training defaults to tf32x3.
    assert med > 0.9 and coss[0][0] > 0.7, coss[:5]


@pytest.mark.parametrize('name,res,batch', [('tiny', 32, 3), ('shapes', 64, 2)])
def test_train_step_default_mode_gradient_cosine(name, res, batch):
    """The product's DEFAULT training arithmetic ('auto' -> tf32x3 tensor-core convolutions): composed G-step
    parameter gradients against the fp32 CPU oracle, cosine per parameter tensor.  Bar: median >= 0.9999 (SURVEY
    8(c)); the tail is bounded by the reference's own conditioning (tools/noise_sensitivity_tiny.py: 1e-6 relative
    noise on the oracle's conv outputs moves single gradients by up to 2e-1)."""
    cfg = helpers.tiny_config() if name == 'tiny' else helpers.load_config(name)
    coss = _grad_cosines(cfg, res, batch, 'auto')
    med, p10 = coss[len(coss) // 2][0], coss[len(coss) // 10][0]
    print('%s auto (tf32x3) train step: gradient cosine vs fp32 oracle: median %.6f, 10th percentile %.6f, 5 worst %s'
          % (name, med, p10, coss[:5]))
    assert med >= 0.9999 and p10 >= 0.999 and coss[0][0] > 0.9, coss[:5]


def test_conv_tc_rejects_unsupported_shapes_without_touching_output():
    from monkey_net_b200 import lib
    dev = torch.device('cuda')
    x = torch.randn(1, 8, 8, 6, device=dev)
    y = torch.zeros(1, 8, 8, 16, device=dev)
    w = torch.zeros(9 * 6 * 16, device=dev)
    with pytest.raises(RuntimeError, match='unsupported'):  # pixel stride 6 floats is not 16-byte aligned
        lib.call('mk_conv2d_tc', x.data_ptr(), 1, 8, 8, 6, 6, 0, w.data_ptr(), 3, 3, 1, None, None, None, 0, 0, 0.0,
                 y.data_ptr(), 16, 16, torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize('mode', ['auto', 'tf32'])
@pytest.mark.parametrize('name,res', [('taichi', 64), ('shapes', 64)])
def test_generator_tensor_core_modes_against_oracle(name, res, mode):
    """Whole keypoint detector + generator, eval / no_grad, against the fp32 CPU oracle.
      'auto' (product default, what bench.py times): geometry networks tf32x3, appearance path 1xTF32 -
             keypoints <= 2e-5 AND identical pixel indices (the logger.py:99-100 rule, north-star "bit-exact"),
             frame <= 1e-3 (north-star), deformed frame <= 1e-3;
      'tf32' (everything 1xTF32): frame <= 1e-3, keypoints <= 1e-4 (pixel indices may flip next to a .5 boundary)."""
    from monkey_net_b200 import ops
    import test_gpu_2_modules as t2
    cfg = helpers.load_config(name)
    (gen, disc, kp), (og, od, ok), x = t2._pair(cfg, res, 2, d=1)
    for m in (gen, kp, og, ok):
        m.eval()
    prev = ops.CONV_MODE
    ops.set_conv_mode(mode)
    try:
        with torch.no_grad():
            a = kp(x['video'].cuda())
            b = ok(x['video'])
            ks = {k: v for k, v in b.items()}
            oa = og(x['source'], kp_driving=b, kp_source=ks)
            bc = {k: v.cuda() for k, v in b.items()}
            ga = gen(x['source'].cuda(), kp_driving=bc, kp_source=bc)
    finally:
        ops.set_conv_mode(prev)
    e_kp = helpers.max_abs(a['mean'], b['mean'])
    e_pred = helpers.max_abs(ga['video_prediction'], oa['video_prediction'])
    e_def = helpers.max_abs(ga['video_deformed'], oa['video_deformed'])
    same_px = torch.equal(torch.round(res * (a['mean'].cpu() + 1) / 2), torch.round(res * (b['mean'] + 1) / 2))
    print('%s mode %s: |kp mean| %.2e  |prediction| %.2e  |deformed| %.2e  identical pixel indices: %s'
          % (mode, name, e_kp, e_pred, e_def, same_px))
    assert e_pred < 1e-3
    if mode == 'auto':
        assert e_kp < 2e-5 and same_px and e_def < 1e-3
    else:
        assert e_kp < 1e-4


# ------------------------------------------------------------------------------------------------ halo-window kernel
HALO_CASES = [
    # cin, cout, k, pad, H, W, N, resid, act
    (48, 48, 3, 1, 64, 64, 16, True, 1),      # refinement ResBlock conv of taichi: 2 chunks, resident weights, residual
    (24, 24, 3, 1, 64, 64, 32, True, 0),      # shapes.yaml ResBlock
    (4, 32, 3, 1, 64, 64, 8, False, 0),       # image input: 4 physical channels ride on the TMA zero fill
    (64, 128, 4, 0, 61, 61, 8, False, 0),     # discriminator 4x4 valid, TWv = 13, 4 output groups
    (160, 32, 3, 1, 64, 64, 8, False, 1),     # 5 channel chunks: 45 weight slots -> streaming ring
    (32, 144, 3, 1, 60, 52, 8, False, 2),     # two cout tiles (128 + 16), partial tiles both ways, sigmoid
    (44, 44, 1, 0, 64, 64, 12, False, 0),     # 1x1 (TWv = 16)
    (16, 64, 4, 3, 61, 61, 8, False, 0),      # dgrad of a 4x4 valid conv: full correlation, pad 3
    (64, 128, 3, 1, 64, 64, 16, False, 0),    # weights do not fit: streaming, several row-blocks per super-tile
    (48, 48, 3, 1, 61, 50, 10, True, 1),      # H, W not multiples of the tile
    (128, 32, 3, 1, 64, 64, 8, False, 0),     # 36 resident weight slots
    (36, 12, 3, 1, 64, 64, 16, False, 0),     # hourglass head: Cout_p = 12 (one 16-column accumulator)
    (64, 16, 4, 3, 61, 61, 8, False, 0),      # column taps on N: 4 x 16 = 64 accumulator columns, full correlation
    (96, 64, 3, 1, 64, 64, 8, False, 2),      # column taps on N: 3 x 64 = 192 columns, two output groups, sigmoid
    (64, 64, 4, 0, 67, 67, 8, True, 1),       # column taps on N at the limit N = 256, residual
]


@pytest.mark.parametrize('kernel', ['halo', 'halo_x3'])
@pytest.mark.parametrize('cin,cout,k,pad,H,W,N,resid,act', HALO_CASES)
def test_conv_halo_matches_torch_conv2d(cin, cout, k, pad, H, W, N, resid, act, kernel):
    """csrc/conv_halo.cu (persistent, halo windows, resident weights, TMA-store epilogue) vs F.conv2d in double."""
    from monkey_net_b200 import lib
    torch.manual_seed(cin + cout + k)
    dev = torch.device('cuda')
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(N, H, W, cin, device=dev)
    w = torch.randn(cout, cin, 1, k, k, device=dev) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device=dev)
    Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    r = torch.randn(N, Ho, Wo, cout, device=dev) if resid else None
    x3 = kernel == 'halo_x3'
    wp = torch.empty(k * k * cout * (cin + (((cin + 7) & ~7) if x3 else 0)), device=dev)
    bp = torch.empty(cout, device=dev)
    lib.call('mk_pack_weight', w.data_ptr(), cout, cin, k, k, 1, None, cin, cout, 2 | (8 if x3 else 0), wp.data_ptr(),
             b.data_ptr(), bp.data_ptr(), st)
    y = torch.full((N, Ho, Wo, cout), float('nan'), device=dev)
    lib.call('mk_conv2d_tc_halo_x3' if x3 else 'mk_conv2d_tc_halo', x.data_ptr(), N, H, W, cin, cin, wp.data_ptr(), k, k,
             pad, None, bp.data_ptr(), r.data_ptr() if resid else None, cout if resid else 0, act, 0.0, y.data_ptr(), cout,
             cout, st)
    torch.cuda.synchronize()
    assert not torch.isnan(y).any(), 'halo kernel left outputs unwritten'
    ref = _torch_conv(x, w, b, pad, r, act)
    err = float((y.double() - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
    assert err < KERNEL_TOL['tf32x3' if x3 else 'tf32'], err


def test_conv_halo_declines_small_layers_untouched():
    """few-tile / small-image layers are left to mk_conv2d_tc: return code -2 before anything is launched"""
    from monkey_net_b200 import lib
    dev = torch.device('cuda')
    x = torch.randn(2, 8, 8, 64, device=dev)
    y = torch.zeros(2, 8, 8, 64, device=dev)
    w = torch.zeros(9 * 64 * 64, device=dev)
    rc = lib.call_soft('mk_conv2d_tc_halo', (-2,), x.data_ptr(), 2, 8, 8, 64, 64, w.data_ptr(), 3, 3, 1, None, None, None,
                       0, 0, 0.0, y.data_ptr(), 64, 64, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rc == -2 and float(y.abs().max()) == 0.0


WGRAD_HALO_CASES = [
    # cin, cout, k, pad, H, W, N
    (48, 48, 3, 1, 64, 64, 4),       # two ci chunks (32 + 16), co = 48
    (24, 24, 3, 1, 64, 64, 8),
    (4, 32, 3, 1, 32, 32, 2),        # image input
    (128, 32, 3, 1, 32, 32, 2),      # 4 ci chunks on the TMEM columns
    (32, 128, 3, 1, 32, 32, 2),
    (140, 32, 3, 1, 30, 26, 3),      # 5 chunks -> 2 chunk groups, partial tiles
    (16, 64, 4, 0, 35, 35, 2),       # 4x4 valid: all four tap rows useful
    (64, 128, 4, 0, 61, 61, 2),
    (36, 12, 3, 1, 64, 64, 4),       # co = 12 -> N = 16
    (64, 256, 3, 1, 32, 32, 2),      # two co tiles
]


@pytest.mark.parametrize('kernel', ['whalo', 'whalo_x3'])
@pytest.mark.parametrize('cin,cout,k,pad,H,W,N', WGRAD_HALO_CASES)
def test_wgrad_halo_matches_torch_autograd(cin, cout, k, pad, H, W, N, kernel):
    """csrc/wgrad_halo.cu (halo windows, tap rows on the M dimension) vs autograd of F.conv2d in double."""
    from monkey_net_b200 import lib
    torch.manual_seed(cin + cout)
    dev = torch.device('cuda')
    st = torch.cuda.current_stream().cuda_stream
    Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    x = torch.randn(N, H, W, cin, device=dev)
    dy = torch.randn(N, Ho, Wo, cout, device=dev)
    d = torch.full((k * k * cin * cout,), float('nan'), device=dev)
    x3 = kernel == 'whalo_x3'
    lib.call('mk_conv2d_wgrad_halo_x3' if x3 else 'mk_conv2d_wgrad_halo', x.data_ptr(), N, H, W, cin, cin, dy.data_ptr(),
             cout, cout, k, k, pad, d.data_ptr(), st)
    torch.cuda.synchronize()
    assert not torch.isnan(d).any()
    wz = torch.zeros(cout, cin, k, k, device=dev, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), wz, None, padding=pad)
    (gw,) = torch.autograd.grad(y, wz, dy.double().permute(0, 3, 1, 2))
    ref = gw.permute(2, 3, 1, 0).reshape(-1)          # (Co,Ci,R,S) -> [tap][Ci][Co]
    err = float((d.double() - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
    assert err < (4e-5 if x3 else 2e-3), err
