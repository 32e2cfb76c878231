"""Building blocks of the frame-generation path - drop-in for the reference's `modules/util.py`.

Same class names, constructor arguments and `state_dict()` keys (`conv.weight (Co,Ci,1,3,3)`, `conv.bias`,
`norm.{weight,bias,running_mean,running_var,num_batches_tracked}`); `nn.Conv3d` / `SynchronizedBatchNorm3d`
objects are kept purely as PARAMETER HOLDERS (identical default init and RNG consumption as the reference), the
math runs through hand-written sm_100a kernels (monkey_net_b200.ops) on NHWC activations:

  DownBlock3D  conv3x3 -> BN -> ReLU -> avgpool(1,2,2)        reference util.py:91-108
  UpBlock3D    nearest x2 -> conv3x3 -> BN -> ReLU            util.py:71-88   (upsample folded into the conv gather)
  SameBlock3D  grouped conv -> BN -> ReLU                     util.py:111-126 (block-diagonal packed weight)
  ResBlock3D   BN-ReLU-conv, BN-ReLU-conv, += x               util.py:45-68   (residual add in the conv epilogue)
  Encoder / Decoder / Hourglass                               util.py:129-203 (torch.cat never materialised
                                                              separately: the BN-apply kernel writes the concat)

Every block has `run(act) -> act` on the internal layout and a `forward(x)` that accepts the reference's
(B,C,D,H,W) tensors.  `temporal=True` (3x3x3 kernels) is not used by any caller or config and is not implemented.
"""
import torch
from torch import nn

from sync_batchnorm import SynchronizedBatchNorm3d as BatchNorm3d
from monkey_net_b200 import ops


def make_coordinate_grid(spatial_size, type):
    """(h, w, 2) mesh over [-1,1]^2, last dim (x, y) (reference util.py:26-42).  Host-side helper only: the kernels
    compute the same coordinates in registers."""
    h, w = spatial_size
    xs = 2 * (torch.arange(w).type(type) / (w - 1)) - 1
    ys = 2 * (torch.arange(h).type(type) / (h - 1)) - 1
    return torch.stack([xs.view(1, w).expand(h, w), ys.view(h, 1).expand(h, w)], dim=2)


def _abcd(m):
    return m[..., 0, 0], m[..., 0, 1], m[..., 1, 0], m[..., 1, 1]


def matrix_inverse(batch_of_matrix, eps=0):
    """Closed-form inverse of (...,2,2) matrices (reference util.py:206-224; its eps=0 branch called the removed
    torch.gesv).  With eps != 0 the determinant is floored at eps as in the reference."""
    a, b, c, d = _abcd(batch_of_matrix)
    det = a * d - b * c
    if eps != 0:
        det = det.clamp(min=eps)
    inv = torch.stack([torch.stack([d, -b], -1), torch.stack([-c, a], -1)], -2)
    return inv / det[..., None, None]


def matrix_det(batch_of_matrix):
    a, b, c, d = _abcd(batch_of_matrix)
    return (a * d - b * c).unsqueeze(-1)


def matrix_trace(batch_of_matrix):
    a, _, _, d = _abcd(batch_of_matrix)
    return (a + d).unsqueeze(-1)


def smallest_singular(batch_of_matrix):
    a, b, c, d = _abcd(batch_of_matrix)
    s1 = a ** 2 + b ** 2 + c ** 2 + d ** 2
    s2 = torch.sqrt((a ** 2 + b ** 2 - c ** 2 - d ** 2) ** 2 + 4 * (a * c + b * d) ** 2)
    return torch.sqrt((s1 - s2) / 2).unsqueeze(-1)


def _spatial(k):
    """(1,k,k)-style kernel/padding spec -> the 2-D value; rejects temporal kernels."""
    if isinstance(k, (tuple, list)):
        if k[0] not in (0, 1) or k[1] != k[2]:
            raise NotImplementedError('temporal (3x3x3) convolutions are not on the B200 hot path')
        return k[1]
    return k


class _Block(nn.Module):
    """forward() adapter: reference-layout tensor in, reference-layout tensor out."""

    def forward(self, x):
        b = x.shape[0]
        return ops.from_nhwc(ops.compact(self.run(ops.to_nhwc(x))), b)


class ResBlock3D(_Block):
    def __init__(self, in_features, kernel_size, padding):
        super(ResBlock3D, self).__init__()
        self.conv1 = nn.Conv3d(in_channels=in_features, out_channels=in_features, kernel_size=kernel_size,
                               padding=padding)
        self.conv2 = nn.Conv3d(in_channels=in_features, out_channels=in_features, kernel_size=kernel_size,
                               padding=padding)
        self.norm1 = BatchNorm3d(in_features, affine=True)
        self.norm2 = BatchNorm3d(in_features, affine=True)
        self.pad = _spatial(padding)
        _spatial(kernel_size)

    def run(self, a):
        a = ops.compact(a)
        t = ops.norm_act(a, self.norm1, mode='bn', slope=0.0)
        t = ops.conv(t, self.conv1.weight, self.conv1.bias, pad=self.pad, feeds_train_norm=bool(self.norm2.training))
        t = ops.norm_act(t, self.norm2, mode='bn', slope=0.0)
        return ops.conv(t, self.conv2.weight, self.conv2.bias, pad=self.pad, resid=a)


class UpBlock3D(_Block):
    def __init__(self, in_features, out_features, kernel_size=3, padding=1):
        super(UpBlock3D, self).__init__()
        self.conv = nn.Conv3d(in_channels=in_features, out_channels=out_features, kernel_size=kernel_size,
                              padding=padding)
        self.norm = BatchNorm3d(out_features, affine=True)
        self.pad = _spatial(padding)
        _spatial(kernel_size)

    def run(self, a, extras=()):
        return ops.conv_bn_relu(a, self.conv, self.norm, self.pad, ups=True, extras=extras)


class DownBlock3D(_Block):
    def __init__(self, in_features, out_features, kernel_size=3, padding=1):
        super(DownBlock3D, self).__init__()
        self.conv = nn.Conv3d(in_channels=in_features, out_channels=out_features, kernel_size=kernel_size,
                              padding=padding)
        self.norm = BatchNorm3d(out_features, affine=True)
        self.pool = nn.AvgPool3d(kernel_size=(1, 2, 2))
        self.pad = _spatial(padding)
        _spatial(kernel_size)

    def run(self, a):
        return ops.conv_bn_relu(a, self.conv, self.norm, self.pad, pool=1)


class SameBlock3D(_Block):
    def __init__(self, in_features, out_features, groups=None, kernel_size=3, padding=1):
        super(SameBlock3D, self).__init__()
        self.conv = nn.Conv3d(in_channels=in_features, out_channels=out_features, kernel_size=kernel_size,
                              padding=padding, groups=groups if groups is not None else 1)
        self.norm = BatchNorm3d(out_features, affine=True)
        self.pad = _spatial(padding)
        self.groups = groups if groups is not None else 1
        _spatial(kernel_size)

    def run(self, a):
        return ops.conv_bn_relu(a, self.conv, self.norm, self.pad, groups=self.groups)


class Encoder(nn.Module):
    """Returns the list of skips, input first (util.py:148-152)."""

    def __init__(self, block_expansion, in_features, num_blocks=3, max_features=256, temporal=False):
        super(Encoder, self).__init__()
        if temporal:
            raise NotImplementedError('temporal=True is not on the B200 hot path')
        down_blocks = []
        for i in range(num_blocks):
            down_blocks.append(DownBlock3D(in_features if i == 0 else min(max_features, block_expansion * (2 ** i)),
                                           min(max_features, block_expansion * (2 ** (i + 1))),
                                           kernel_size=(1, 3, 3), padding=(0, 1, 1)))
        self.down_blocks = nn.ModuleList(down_blocks)

    def run(self, a):
        outs = [a]
        for blk in self.down_blocks:
            outs.append(blk.run(outs[-1]))
        return outs

    def forward(self, x):
        b = x.shape[0]
        outs = self.run(ops.to_nhwc(x))
        return [x] + [ops.from_nhwc(o, b) for o in outs[1:]]


class Decoder(nn.Module):
    def __init__(self, block_expansion, in_features, out_features, num_blocks=3, max_features=256, temporal=False,
                 additional_features_for_block=0, use_last_conv=True):
        super(Decoder, self).__init__()
        if temporal:
            raise NotImplementedError('temporal=True is not on the B200 hot path')
        up_blocks = []
        for i in range(num_blocks)[::-1]:
            up_blocks.append(UpBlock3D((1 if i == num_blocks - 1 else 2) * min(max_features, block_expansion * (
                2 ** (i + 1))) + additional_features_for_block,
                                       min(max_features, block_expansion * (2 ** i)),
                                       kernel_size=(1, 3, 3), padding=(0, 1, 1)))
        self.up_blocks = nn.ModuleList(up_blocks)
        if use_last_conv:
            self.conv = nn.Conv3d(in_channels=block_expansion + in_features + additional_features_for_block,
                                  out_channels=out_features, kernel_size=(1, 3, 3), padding=(0, 1, 1))
        else:
            self.conv = None

    def run(self, skips):
        """skips: list (shallow -> deep); each entry an Act or a list of Acts forming that level's concat."""
        skips = [s if isinstance(s, (list, tuple)) else [s] for s in skips]
        deepest = skips.pop()
        out = deepest[0] if len(deepest) == 1 else ops.concat(deepest)
        for blk in self.up_blocks:
            out = blk.run(out, extras=skips.pop())
        if self.conv is not None:
            return ops.conv(out, self.conv.weight, self.conv.bias, pad=1)
        return out

    def forward(self, x):
        b = x[0].shape[0]
        out = self.run([ops.to_nhwc(t) for t in x])
        return ops.from_nhwc(ops.compact(out), b)


class Hourglass(nn.Module):
    def __init__(self, block_expansion, in_features, out_features, num_blocks=3, max_features=256, temporal=False):
        super(Hourglass, self).__init__()
        self.encoder = Encoder(block_expansion, in_features, num_blocks, max_features, temporal=temporal)
        self.decoder = Decoder(block_expansion, in_features, out_features, num_blocks, max_features,
                               temporal=temporal)

    def run(self, a):
        return self.decoder.run(self.encoder.run(a))

    def forward(self, x):
        b = x.shape[0]
        return ops.from_nhwc(ops.compact(self.run(ops.to_nhwc(x))), b)
