"""Op-level parity: every C-ABI kernel family (through the autograd wrappers in monkey_net_b200.ops) against the CPU
oracle / the plain torch op it replaces, forward AND gradients, on seeded inputs.  Tolerances are fp32-rounding
level: 2e-5 forward, 2e-4 relative on gradients (different summation orders)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import max_abs, rel_err

pytestmark = pytest.mark.gpu

FWD_TOL = 2e-5
GRAD_TOL = 2e-4


def ops():
    from monkey_net_b200 import ops as o
    return o


def oracle():
    from oracle import monkey_oracle as mo
    return mo


def dev(t, grad=False):
    t = t.detach().clone().cuda()
    t.requires_grad_(grad)
    return t


def run5(fn, x5):
    """x5 CPU (B,C,D,H,W) -> cuda leaf, apply fn(Act)->Act, return (out5 cuda, leaf)."""
    o = ops()
    leaf = dev(x5, True)
    out = fn(o.to_nhwc(leaf))
    return o.from_nhwc(o.compact(out), x5.shape[0]), leaf


def check_grads(out_gpu, out_ref, pairs, seed=0):
    """Backprop the same random cotangent through both graphs and compare gradients of (gpu_leaf, ref_leaf) pairs."""
    g = torch.Generator().manual_seed(seed)
    r = torch.randn(out_ref.shape, generator=g)
    (out_gpu * r.cuda()).sum().backward()
    (out_ref * r).sum().backward()
    for a, b in pairs:
        assert a.grad is not None and b.grad is not None
        assert rel_err(a.grad, b.grad) < GRAD_TOL, rel_err(a.grad, b.grad)


# ------------------------------------------------------------------------------------------------------ layout
@pytest.mark.parametrize('step', [1, 2, 4])
def test_layout_roundtrip(step):
    o = ops()
    torch.manual_seed(0)
    x = torch.rand(2, 3, 2, 16, 24)
    big = torch.rand(2, 3, 5, 16, 24)
    big[:, :, 1:3] = x
    src = big.cuda()[:, :, 1:3]  # non-contiguous D-slice, like transfer.py:68
    a = o.to_nhwc(src, step)
    assert a.t.shape == (4, 16 // step, 24 // step, 4)
    ref = x[:, :, :, ::step, ::step]
    back = o.from_nhwc(a, 2)
    assert back.shape == ref.shape
    assert max_abs(back, ref) == 0.0
    assert float(a.t[..., 3].abs().max()) == 0.0


def test_layout_grad():
    o = ops()
    x = dev(torch.rand(2, 3, 1, 8, 8), True)
    y = o.from_nhwc(o.to_nhwc(x, 2), 2)
    r = torch.rand(y.shape).cuda()
    (y * r).sum().backward()
    ref = torch.zeros(2, 3, 1, 8, 8)
    ref[:, :, :, ::2, ::2] = r.cpu()
    assert max_abs(x.grad, ref) == 0.0


# ------------------------------------------------------------------------------------------------------ conv
CONV_CASES = [
    # cin, cout, k, pad, ups, groups, H, W, B
    (3, 16, 3, 1, False, 1, 16, 16, 2),
    (16, 32, 3, 1, False, 1, 9, 7, 3),
    (35, 10, 3, 1, False, 1, 8, 8, 2),
    (64, 128, 3, 1, True, 1, 4, 4, 2),
    (20, 20, 1, 0, False, 5, 8, 8, 2),
    (13, 32, 4, 0, False, 1, 13, 13, 2),
    (128, 64, 3, 1, False, 1, 2, 2, 4),
    (23, 3, 1, 0, False, 1, 8, 8, 2),
    (40, 72, 3, 1, True, 1, 6, 5, 1),
]


@pytest.mark.parametrize('cin,cout,k,pad,ups,groups,H,W,B', CONV_CASES)
def test_conv_fwd_bwd(cin, cout, k, pad, ups, groups, H, W, B):
    o = ops()
    torch.manual_seed(cin * 7 + cout)
    x = torch.randn(B, cin, 1, H, W)
    w = torch.randn(cout, cin // groups, 1, k, k) / (cin * k * k / groups) ** 0.5
    b = torch.randn(cout)
    wg, bg = dev(w, True), dev(b, True)
    out, xg = run5(lambda a: o.conv(a, wg, bg, pad=pad, groups=groups, ups=ups), x)
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    xi = xr[:, :, 0]
    if ups:
        xi = F.interpolate(xi, scale_factor=2)
    ref = F.conv2d(xi, wr[:, :, 0], br, padding=pad, groups=groups)[:, :, None]
    assert out.shape == ref.shape
    assert max_abs(out, ref) < 5e-5
    check_grads(out, ref, [(xg, xr), (wg, wr), (bg, br)])


def test_conv_concat_holes_and_residual():
    """Input = concat of padded segments (23 | 3 | 10 logical channels) + residual add in the epilogue."""
    o = ops()
    torch.manual_seed(3)
    parts = [torch.randn(2, c, 1, 8, 8) for c in (23, 3, 10)]
    w = torch.randn(36, 36, 1, 3, 3) * 0.05
    b = torch.randn(36)
    res = torch.randn(2, 36, 1, 8, 8)
    leaves = [dev(p, True) for p in parts]
    wg, bg, rg = dev(w, True), dev(b, True), dev(res, True)
    acts = [o.to_nhwc(p) for p in leaves]
    cat = o.concat(acts)
    assert cat.t.shape[3] == 24 + 4 + 12
    y = o.conv(cat, wg, bg, pad=1, resid=o.to_nhwc(rg))
    out = o.from_nhwc(y, 2)
    refs = [p.clone().requires_grad_() for p in parts]
    wr, br, rr = w.clone().requires_grad_(), b.clone().requires_grad_(), res.clone().requires_grad_()
    ref = F.conv2d(torch.cat(refs, 1)[:, :, 0], wr[:, :, 0], br, padding=1)[:, :, None] + rr
    assert max_abs(out, ref) < 5e-5
    check_grads(out, ref, list(zip(leaves, refs)) + [(wg, wr), (bg, br), (rg, rr)])
    # compact() gathers the logical channels of the holed tensor
    comp = o.from_nhwc(o.compact(cat), 2)
    assert max_abs(comp, torch.cat(parts, 1)) == 0.0


def test_conv_sigmoid_epilogue():
    o = ops()
    torch.manual_seed(4)
    x = torch.randn(2, 23, 1, 8, 8)
    w, b = torch.randn(3, 23, 1, 1, 1) * 0.3, torch.randn(3)
    wg, bg = dev(w, True), dev(b, True)
    out, xg = run5(lambda a: o.conv(a, wg, bg, pad=0, act='sigmoid'), x)
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    ref = torch.sigmoid(F.conv2d(xr[:, :, 0], wr[:, :, 0], br))[:, :, None]
    assert max_abs(out, ref) < FWD_TOL
    check_grads(out, ref, [(xg, xr), (wg, wr), (bg, br)])


# ------------------------------------------------------------------------------------------------------ norm
class _Holder(torch.nn.Module):
    def __init__(self, c, bn=True):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.rand(c) + 0.5)
        self.bias = torch.nn.Parameter(torch.randn(c))
        if bn:
            self.register_buffer('running_mean', torch.randn(c))
            self.register_buffer('running_var', torch.rand(c) + 0.5)
            self.register_buffer('num_batches_tracked', torch.tensor(3, dtype=torch.long))


@pytest.mark.parametrize('c,H,W,pool,slope,training', [(16, 8, 8, 1, 0.0, True), (13, 9, 7, 1, 0.0, True),
                                                       (32, 6, 6, 0, 0.0, True), (45, 8, 8, 0, 0.0, False),
                                                       (256, 4, 4, 0, 0.0, True), (512, 2, 2, 1, 0.0, True)])
def test_batchnorm_act_pool(c, H, W, pool, slope, training):
    o = ops()
    torch.manual_seed(c)
    x = torch.randn(3, c, 2, H, W) * 2 + 0.5
    hold = _Holder(c)
    ref_mod = _Holder(c)
    ref_mod.load_state_dict(hold.state_dict())
    hold = hold.cuda()
    hold.train(training)
    out, xg = run5(lambda a: o.norm_act(a, hold, mode='bn', slope=slope, pool=pool), x)
    xr = x.clone().requires_grad_()
    x4 = xr.permute(0, 2, 1, 3, 4).reshape(6, c, H, W)
    y = F.relu(F.batch_norm(x4, ref_mod.running_mean, ref_mod.running_var, ref_mod.weight, ref_mod.bias, training,
                            0.1, 1e-5))
    if pool:
        y = F.avg_pool2d(y, 2)
    ref = y.reshape(3, 2, c, y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)
    assert max_abs(out, ref) < 5e-5
    if training:
        assert max_abs(hold.running_mean, ref_mod.running_mean) < 1e-5
        assert max_abs(hold.running_var, ref_mod.running_var) < 1e-5
        assert int(hold.num_batches_tracked) == 4
        check_grads(out, ref, [(xg, xr), (hold.weight, ref_mod.weight), (hold.bias, ref_mod.bias)])


def test_norm_concat_extras():
    o = ops()
    torch.manual_seed(9)
    x, e1, e2 = torch.randn(2, 16, 1, 8, 8), torch.randn(2, 13, 1, 8, 8), torch.randn(2, 4, 1, 8, 8)
    hold = _Holder(16).cuda().train()
    xg, e1g, e2g = dev(x, True), dev(e1, True), dev(e2, True)
    y = o.norm_act(o.to_nhwc(xg), hold, mode='bn', slope=0.0, extras=[o.to_nhwc(e1g), o.to_nhwc(e2g)])
    assert y.segs == ((16, 16), (13, 16), (4, 4))
    out = o.from_nhwc(o.compact(y), 2)
    xr, e1r, e2r = (t.clone().requires_grad_() for t in (x, e1, e2))
    w, b = hold.weight.detach().cpu().requires_grad_(), hold.bias.detach().cpu().requires_grad_()
    bn = F.relu(F.batch_norm(xr[:, :, 0], None, None, w, b, True, 0.1, 1e-5))
    ref = torch.cat([bn, e1r[:, :, 0], e2r[:, :, 0]], 1)[:, :, None]
    assert max_abs(out, ref) < 5e-5
    check_grads(out, ref, [(xg, xr), (e1g, e1r), (e2g, e2r)])


@pytest.mark.parametrize('norm', [True, False])
def test_instancenorm_leaky_pool(norm):
    o = ops()
    torch.manual_seed(11)
    x = torch.randn(3, 24, 1, 13, 13)
    hold = _Holder(24, bn=False).cuda() if norm else None
    out, xg = run5(lambda a: o.norm_act(a, hold, mode='in' if norm else 'none', slope=0.2, pool=1), x)
    xr = x.clone().requires_grad_()
    y = xr[:, :, 0]
    if norm:
        w, b = hold.weight.detach().cpu().requires_grad_(), hold.bias.detach().cpu().requires_grad_()
        y = F.instance_norm(y, None, None, w, b, True, 0.1, 1e-5)
    ref = F.avg_pool2d(F.leaky_relu(y, 0.2), 2)[:, :, None]
    assert out.shape == ref.shape == (3, 24, 1, 6, 6)
    assert max_abs(out, ref) < 5e-5
    g = torch.randn(ref.shape)
    (out * g.cuda()).sum().backward()
    (ref * g).sum().backward()
    assert rel_err(xg.grad, xr.grad) < GRAD_TOL
    if norm:
        assert rel_err(hold.weight.grad, w.grad) < GRAD_TOL and rel_err(hold.bias.grad, b.grad) < GRAD_TOL


# ------------------------------------------------------------------------------------------------------ sampling
@pytest.mark.parametrize('c,h,h0,mode,d', [(3, 16, 16, 'nearest', 1), (16, 8, 16, 'nearest', 2),
                                            (32, 4, 16, 'trilinear', 1), (128, 2, 16, 'nearest', 1),
                                            (8, 32, 16, 'trilinear', 1), (1024, 2, 8, 'nearest', 1),
                                            (20, 13, 16, 'nearest', 1),     # 5 channel vectors, partial 16x16 patches
                                            (68, 9, 16, 'trilinear', 2),    # 17 channel vectors: ragged last item pass
                                            (64, 24, 16, 'nearest', 1)])    # 3x3 patches of 8x8, 4 items per thread
def test_grid_sample(c, h, h0, mode, d):
    o = ops()
    torch.manual_seed(c + h)
    B = 2
    inp = torch.randn(B, c, 1, h, h)
    base = oracle().coord_grid(h0, h0, inp)[None].expand(B * d, h0, h0, 2)
    deform = (base + 0.3 * torch.randn(B * d, h0, h0, 2)).contiguous()
    ig, dg = dev(inp, True), dev(deform, True)
    out = o.from_nhwc(o.grid_sample(o.to_nhwc(ig), dg, d, mode), B)
    ir, dr = inp.clone().requires_grad_(), deform.clone().requires_grad_()
    g = dr.permute(0, 3, 1, 2)
    if h != h0:
        g = F.interpolate(g, size=(h, h), mode='nearest') if mode == 'nearest' else \
            F.interpolate(g, size=(h, h), mode='bilinear', align_corners=False)
    src = ir[:, None, :, 0].expand(B, d, c, h, h).reshape(B * d, c, h, h)
    y = F.grid_sample(src, g.permute(0, 2, 3, 1), mode='bilinear', padding_mode='zeros', align_corners=True)
    ref = y.reshape(B, d, c, h, h).permute(0, 2, 1, 3, 4)
    assert max_abs(out, ref) < FWD_TOL
    check_grads(out, ref, [(ig, ir), (dg, dr)])


@pytest.mark.parametrize('mode', ['nearest', 'trilinear'])
@pytest.mark.parametrize('h0,h', [(16, 8), (16, 2), (8, 16)])
def test_resize(mode, h0, h):
    o = ops()
    torch.manual_seed(h)
    x = torch.randn(2, 10, 1, h0, h0)
    out, xg = run5(lambda a: o.resize(a, h, h, mode), x)
    xr = x.clone().requires_grad_()
    y = F.interpolate(xr[:, :, 0], size=(h, h), mode='nearest') if mode == 'nearest' else \
        F.interpolate(xr[:, :, 0], size=(h, h), mode='bilinear', align_corners=False)
    ref = y[:, :, None]
    assert max_abs(out, ref) < FWD_TOL
    check_grads(out, ref, [(xg, xr)])


# ------------------------------------------------------------------------------------------------------ keypoints
@pytest.mark.parametrize('K,H,kp_variance,clip', [(10, 16, 'matrix', None), (4, 32, 'matrix', 0.5),
                                                   (10, 16, 'matrix', 0.001), (5, 16, 'single', None),
                                                   # several pixel chunks per frame (chunked many-block kernels):
                                                   (10, 64, 'matrix', 0.001), (7, 50, 'matrix', None),
                                                   (3, 64, 'single', None), (13, 48, 'matrix', 0.5)])
def test_kp_head(K, H, kp_variance, clip):
    o, mo = ops(), oracle()
    torch.manual_seed(K + H)
    B, D = 2, 2
    logits = torch.randn(B, K, D, H, H) * 0.3
    lg = dev(logits, True)
    kp = o.kp_head(o.to_nhwc(lg), B, D, K, 0.1, kp_variance, clip)
    lr = logits.clone().requires_grad_()
    p = F.softmax(lr.reshape(B, K, D, -1) / 0.1, dim=3).reshape(B, K, D, H, H)
    ref = mo.heat_to_kp(p, kp_variance, clip)
    assert kp['mean'].shape == ref['mean'].shape and kp['var'].shape == ref['var'].shape
    assert max_abs(kp['mean'], ref['mean']) < 1e-5
    assert rel_err(kp['var'], ref['var']) < 5e-5
    # keypoint pixel indices (Visualizer: spatial_size*(mean+1)/2) and heatmap argmax are bit-exact
    assert torch.equal(torch.round(H * (kp['mean'].cpu() + 1) / 2), torch.round(H * (ref['mean'] + 1) / 2))
    g1, g2 = torch.randn(ref['mean'].shape), torch.randn(ref['var'].shape)
    ((kp['mean'] * g1.cuda()).sum() + (kp['var'] * g2.cuda()).sum()).backward()
    ((ref['mean'] * g1).sum() + (ref['var'] * g2).sum()).backward()
    assert rel_err(lg.grad, lr.grad) < 5e-4


def _rand_kp(B, d, K, seed, spread=0.5):
    g = torch.Generator().manual_seed(seed)
    mean = (torch.rand(B, d, K, 2, generator=g) * 2 - 1) * spread
    a = torch.randn(B, d, K, 2, 2, generator=g) * 0.15
    var = a @ a.transpose(-1, -2) + 0.05 * torch.eye(2)
    return mean, var


EMBED_CASES = [
    # use_heatmap, use_difference, use_deformed, add_bg, heatmap_type, norm_const, kp_variance
    (True, False, True, True, 'difference', 100, 'matrix'),
    (True, True, True, True, 'difference', 100, 'matrix'),
    (True, False, False, False, 'gaussian', 10, 'matrix'),
    (True, False, False, False, 'difference', 'sum', 'matrix'),
    (False, True, False, True, 'gaussian', 'sum', 'matrix'),
    (True, False, False, False, 'gaussian', 10, 0.01),
    (True, False, False, False, 'difference', 10, 'single'),
]


@pytest.mark.parametrize('hm,diff,deformed,bg,htype,norm,kpv', EMBED_CASES)
def test_movement_embedding(hm, diff, deformed, bg, htype, norm, kpv):
    o, mo = ops(), oracle()
    B, d, K, h = 2, 2, 3, 16
    torch.manual_seed(5)
    src = torch.rand(B, 3, 1, h, h)
    md, vd = _rand_kp(B, d, K, 1)
    ms, vs = _rand_kp(B, 1, K, 2)
    if kpv == 'single':
        vd, vs = vd[..., :1, :1].contiguous().abs() + 0.05, vs[..., :1, :1].contiguous().abs() + 0.05
    leaves = [dev(t, True) for t in (md, vd, ms, vs)]
    refs = [t.clone().requires_grad_() for t in (md, vd, ms, vs)]
    kd = {'mean': leaves[0], 'var': leaves[1]}
    ks = {'mean': leaves[2], 'var': leaves[3]}
    a = o.movement_embed(o.to_nhwc(src.cuda()) if deformed else None, kd, ks, h, h, num_channels=3, kp_variance=kpv,
                         use_heatmap=hm, use_difference=diff, use_deformed=deformed, add_bg=bg, heatmap_type=htype,
                         norm_const=norm)
    out = o.from_nhwc(a, B)
    m = mo.MovementEmbedding(num_kp=K, kp_variance=kpv, num_channels=3, use_deformed_source_image=deformed,
                             use_difference=diff, use_heatmap=hm, add_bg_feature_map=bg, heatmap_type=htype,
                             norm_const=norm)
    ref = m(src, {'mean': refs[0], 'var': refs[1]}, {'mean': refs[2], 'var': refs[3]})
    assert out.shape == ref.shape
    assert max_abs(out, ref) < 3e-5 * max(1.0, float(ref.abs().max()))
    g = torch.randn(ref.shape)
    (out * g.cuda()).sum().backward()
    (ref * g).sum().backward()
    for a_, b_ in zip(leaves, refs):
        if b_.grad is None:
            assert a_.grad is None or float(a_.grad.abs().max()) == 0.0
            continue
        assert rel_err(a_.grad, b_.grad) < 5e-4, rel_err(a_.grad, b_.grad)


def test_flow_head():
    o, mo = ops(), oracle()
    B, d, K, h = 2, 2, 4, 8
    torch.manual_seed(8)
    pred = torch.randn(B * d, K + 3, 1, h, h)
    md, _ = _rand_kp(B, d, K, 3)
    ms, _ = _rand_kp(B, 1, K, 4)
    pg, mdg, msg = dev(pred, True), dev(md, True), dev(ms, True)
    deform = o.flow_head(o.to_nhwc(pg), {'mean': mdg}, {'mean': msg}, True, True)
    pr, mdr, msr = pred.clone().requires_grad_(), md.clone().requires_grad_(), ms.clone().requires_grad_()
    p4 = pr[:, :, 0]
    mask = F.softmax(p4[:, :K + 1], dim=1)
    shift = (msr - mdr).reshape(B * d, K, 2)
    shift = torch.cat([torch.zeros(B * d, 1, 2), shift], 1)
    flow = torch.einsum('nkhw,nkc->nhwc', mask, shift) + p4[:, -2:].permute(0, 2, 3, 1)
    ref = flow + mo.coord_grid(h, h, pred)[None]
    assert max_abs(deform, ref) < FWD_TOL
    check_grads(deform, ref, [(pg, pr), (mdg, mdr), (msg, msr)])


# ------------------------------------------------------------------------------------------------------ losses
def test_losses():
    o = ops()
    torch.manual_seed(2)
    a, b = torch.rand(3, 5, 1, 7, 9), torch.rand(3, 5, 1, 7, 9)
    ag, bg = dev(a, True), dev(b, True)
    # one operand NHWC-backed (like a discriminator map), the other reference NCDHW
    av = o.from_nhwc(o.to_nhwc(ag), 3)
    l1 = o.loss_mean('l1', av, bg, 10.0)
    gg = o.loss_mean('gen_gan', av, None, 1.0)
    dg = o.loss_mean('disc_gan', av, bg, 2.0)
    ar, br = a.clone().requires_grad_(), b.clone().requires_grad_()
    mb = lambda v: v.reshape(3, -1).mean(-1)
    r1, r2, r3 = 10 * mb((ar - br).abs()), mb((1 - ar) ** 2), 2 * mb((1 - ar) ** 2 + br ** 2)
    for x, y in ((l1, r1), (gg, r2), (dg, r3)):
        assert x.shape == y.shape and max_abs(x, y) < 1e-5
    w = torch.tensor([1.0, -2.0, 0.5])
    ((l1 + gg + dg) * w.cuda()).sum().backward()
    ((r1 + r2 + r3) * w).sum().backward()
    assert rel_err(ag.grad, ar.grad) < 1e-5 and rel_err(bg.grad, br.grad) < 1e-5
