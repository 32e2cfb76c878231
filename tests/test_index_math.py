"""The resize index rules the kernels implement in registers (csrc/common.cuh: nearest_src, linear_src, grid_coord) are
restated here in float32 numpy, operation for operation, and compared with ATen's F.interpolate / the reference's
make_coordinate_grid over every size pair a configuration can produce - the GPU tests only sample a few sizes."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

f32 = np.float32


def nearest_src(dst, n_in, n_out):
    """common.cuh:nearest_src - floor(dst * (in / out)) in fp32, clamped to in-1"""
    scale = f32(n_in) / f32(n_out)
    s = np.floor(dst.astype(f32) * scale).astype(np.int64)
    return np.minimum(s, n_in - 1)


def linear_src(dst, n_in, n_out):
    """common.cuh:linear_src - align_corners=False source index and weight of the upper neighbour"""
    scale = f32(n_in) / f32(n_out)
    # scale * (dst + 0.5) - 0.5 is ONE fused multiply-add on the device (nvcc contracts it, as it does in ATen's
    # kernel): a single rounding, emulated here through float64
    s = (np.float64(scale) * np.float64(dst.astype(f32) + f32(0.5)) - 0.5).astype(f32)
    s = np.maximum(s, f32(0))
    i0 = np.minimum(s.astype(np.int64), n_in - 1)
    i1 = i0 + (i0 < n_in - 1)
    return i0, i1, (s - i0.astype(f32)).astype(f32)


SIZES = [(a, b) for a in (2, 4, 8, 13, 16, 32, 61, 64, 128, 256) for b in (2, 4, 7, 8, 16, 29, 32, 64, 128, 256)]


@pytest.mark.parametrize('n_in,n_out', SIZES)
def test_nearest_index_rule_equals_aten(n_in, n_out):
    src = torch.arange(n_in, dtype=torch.float32).view(1, 1, 1, n_in)
    ref = F.interpolate(src, size=(1, n_out), mode='nearest').view(-1).long().numpy()
    assert np.array_equal(nearest_src(np.arange(n_out), n_in, n_out), ref)


@pytest.mark.parametrize('n_in,n_out', SIZES)
def test_linear_index_rule_equals_aten(n_in, n_out):
    i0, i1, l1 = linear_src(np.arange(n_out), n_in, n_out)
    x = torch.rand(1, 1, 1, n_in, generator=torch.Generator().manual_seed(n_in * 1000 + n_out))
    ref = F.interpolate(x, size=(1, n_out), mode='bilinear', align_corners=False).view(-1).numpy()
    xs = x.view(-1).numpy()
    got = (f32(1) - l1) * xs[i0] + l1 * xs[i1]
    assert np.abs(got - ref).max() < 1e-6  # same indices and weights; the blend itself rounds differently


@pytest.mark.parametrize('n', [2, 3, 16, 32, 64, 128, 256])
def test_coordinate_grid_rule_equals_reference_formula(n):
    """common.cuh:grid_coord = 2*(j/(n-1)) - 1 in fp32 == modules.util.make_coordinate_grid (util.py:26-42)"""
    from modules.util import make_coordinate_grid
    j = np.arange(n).astype(f32)
    mine = f32(2) * (j / f32(n - 1)) - f32(1)
    ref = make_coordinate_grid((n, n), torch.FloatTensor)[0, :, 0].numpy()
    assert np.array_equal(mine, ref)
