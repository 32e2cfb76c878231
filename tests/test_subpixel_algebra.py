"""The algebra behind the sub-pixel kernels of the upsampled conv (modules/util.py:84-85 in the reference:
F.interpolate(scale_factor=(1,2,2)) followed by a 3x3 conv, pad 1), restated in plain torch on the CPU:

  forward   (mk_pack_weight mode 4 + mk_conv2d_tc_halo_ups):  y[2h+py][2w+px] = sum_{r2,s2} Wsub[py,px][r2,s2] x[h+r2-(1-py)][w+s2-(1-px)]
  weight gradient (mk_conv2d_wgrad_halo_ups + mk_unpack_wgrad_ups):  dWsub[py,px][r2,s2] = sum dY[2h+py][2w+px] x[...],
            folded back onto the 3x3 taps by the adjoint of the tap sums.

Tap row r of the 3x3 kernel belongs to sub-row r2 = (r >= 1) for py = 0 ({0} | {1,2}) and r2 = (r >= 2) for py = 1
({0,1} | {2}); same for columns.  The GPU tests check the kernels against torch; this file pins the derivation itself."""
import torch
import torch.nn.functional as F


def _sub(p, r):
    return int(r >= 1) if p == 0 else int(r >= 2)


def _subkernels(w):
    """w (Co,Ci,3,3) -> Wsub[py][px] (Co,Ci,2,2): sums of the 3x3 taps that hit the same low-resolution pixel"""
    out = [[torch.zeros(w.shape[0], w.shape[1], 2, 2, dtype=w.dtype) for _ in range(2)] for _ in range(2)]
    for py in range(2):
        for px in range(2):
            for r in range(3):
                for s in range(3):
                    out[py][px][:, :, _sub(py, r), _sub(px, s)] += w[:, :, r, s]
    return out


def _shifted(x, dr, ds):
    """x[n,c,h+dr,w+ds] with zeros outside"""
    N, C, H, W = x.shape
    xp = F.pad(x, (1, 1, 1, 1))
    return xp[:, :, 1 + dr:1 + dr + H, 1 + ds:1 + ds + W]


def test_forward_is_four_subpixel_2x2_convs():
    torch.manual_seed(0)
    x = torch.randn(2, 5, 6, 7, dtype=torch.double)
    w = torch.randn(4, 5, 3, 3, dtype=torch.double)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode='nearest'), w, padding=1)
    sub = _subkernels(w)
    y = torch.zeros_like(ref)
    for py in range(2):
        for px in range(2):
            acc = 0
            for r2 in range(2):
                for s2 in range(2):
                    acc = acc + torch.einsum('oc,nchw->nohw', sub[py][px][:, :, r2, s2], _shifted(x, r2 - (1 - py), s2 - (1 - px)))
            y[:, :, py::2, px::2] = acc
    assert float((y - ref).abs().max()) < 1e-12


def test_weight_gradient_on_the_low_resolution_grid():
    torch.manual_seed(1)
    x = torch.randn(2, 5, 6, 7, dtype=torch.double)
    w = torch.zeros(4, 5, 3, 3, dtype=torch.double, requires_grad=True)
    dy = torch.randn(2, 4, 12, 14, dtype=torch.double)
    gw, = torch.autograd.grad(F.conv2d(F.interpolate(x, scale_factor=2, mode='nearest'), w, padding=1), w, dy)
    # gradient of the 16 sub-kernels: 16 tap-pixel products per low-resolution pixel (36 at full resolution)
    dsub = [[torch.zeros(4, 5, 2, 2, dtype=torch.double) for _ in range(2)] for _ in range(2)]
    for py in range(2):
        for px in range(2):
            d = dy[:, :, py::2, px::2]
            for r2 in range(2):
                for s2 in range(2):
                    dsub[py][px][:, :, r2, s2] = torch.einsum('nohw,nchw->oc', d, _shifted(x, r2 - (1 - py), s2 - (1 - px)))
    # adjoint of the tap sums (mk_unpack_wgrad_ups): every 3x3 tap collects its four sub-kernel gradients
    dw = torch.zeros(4, 5, 3, 3, dtype=torch.double)
    for r in range(3):
        for s in range(3):
            for py in range(2):
                for px in range(2):
                    dw[:, :, r, s] += dsub[py][px][:, :, _sub(py, r), _sub(px, s)]
    assert float((dw - gw).abs().max()) < 1e-10 * float(gw.abs().max())
