"""CPU oracle for the Monkey-Net frame-generation hot path.

TEST INFRASTRUCTURE ONLY.  This file is a plain-PyTorch fp32 CPU *restatement* of the reference's
algorithm (4-D NCHW tensors with the frame axis D folded into batch, `F.conv2d`, closed-form 2x2
algebra).  It is the checker that `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs compare the CUDA product against.  Nothing under
`modules/`, `sync_batchnorm/` or `monkey-net_b200/` may import it.

Parity pin: the reference ships no golden vectors or tests, so this oracle
is pinned by running the UNMODIFIED reference modules from /root/reference in the build container
(`oracle/ref_shim.py`) and committing their outputs as fixtures (`oracle/make_golden.py` ->
`tests/golden/*.npz`); `tests/test_oracle_golden.py` checks this file against those fixtures and
(when /root/reference is present) against the live reference.

Module tree / parameter names follow the reference's `state_dict()` keys exactly, so a reference
checkpoint, this oracle and the CUDA product all exchange `state_dict`s.

Reference citations are `file:line` into /root/reference.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn


# ----------------------------------------------------------------------------- small algebra
def coord_grid(h, w, like):
    """(h,w,2) grid, last dim (x,y), x_j = 2*(j/(w-1))-1.  modules/util.py:26-42."""
    xs = torch.arange(w, dtype=like.dtype, device=like.device)
    ys = torch.arange(h, dtype=like.dtype, device=like.device)
    xs = 2 * (xs / (w - 1)) - 1
    ys = 2 * (ys / (h - 1)) - 1
    return torch.stack([xs[None, :].expand(h, w), ys[:, None].expand(h, w)], dim=-1)


def inv2x2(m):
    """General (non-symmetric) closed-form 2x2 inverse; replaces torch.gesv at modules/util.py:220-224."""
    a, b, c, d = m[..., 0, 0], m[..., 0, 1], m[..., 1, 0], m[..., 1, 1]
    det = a * d - b * c
    return torch.stack([torch.stack([d, -b], -1), torch.stack([-c, a], -1)], -2) / det[..., None, None]


def sigma_min(m):
    """Smallest singular value of 2x2 matrices.  modules/util.py:244-255."""
    a, b, c, d = m[..., 0, 0], m[..., 0, 1], m[..., 1, 0], m[..., 1, 1]
    s1 = a ** 2 + b ** 2 + c ** 2 + d ** 2
    s2 = torch.sqrt((a ** 2 + b ** 2 - c ** 2 - d ** 2) ** 2 + 4 * (a * c + b * d) ** 2)
    return torch.sqrt((s1 - s2) / 2)


def fold(x5):
    """(B,C,D,H,W) -> (B*D,C,H,W)."""
    b, c, d, h, w = x5.shape
    return x5.permute(0, 2, 1, 3, 4).reshape(b * d, c, h, w)


def unfold(x4, b):
    """(B*D,C,H,W) -> (B,C,D,H,W)."""
    n, c, h, w = x4.shape
    return x4.reshape(b, n // b, c, h, w).permute(0, 2, 1, 3, 4)


def nearest_down(x4, s):
    """F.interpolate(scale_factor=(1,s,s)) nearest: keypoint_detector.py:99, dense_motion_module.py:44."""
    if s == 1:
        return x4
    return F.interpolate(x4, scale_factor=(s, s))


# ----------------------------------------------------------------------------- parameter holders
class _Conv(nn.Module):
    """Holds a Conv3d-shaped weight (Co,Ci/g,1,kh,kw) + bias; applies it as a per-frame 2-D conv."""

    def __init__(self, cin, cout, k, pad, groups=1):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin // groups, 1, k, k))
        self.bias = nn.Parameter(torch.empty(cout))
        self.pad, self.groups = pad, groups

    def forward(self, x4):
        return F.conv2d(x4, self.weight[:, :, 0], self.bias, padding=self.pad, groups=self.groups)


class _BN(nn.Module):
    """Batch norm over (N,D,H,W): biased var to normalise, unbiased into running_var, momentum 0.1
    (sync_batchnorm/batchnorm.py:50-53 -> F.batch_norm)."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer('running_mean', torch.zeros(c))
        self.register_buffer('running_var', torch.ones(c))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))

    def forward(self, x4):
        if self.training:
            self.num_batches_tracked += 1
        return F.batch_norm(x4, self.running_mean, self.running_var, self.weight, self.bias,
                            self.training, 0.1, 1e-5)


class _IN(nn.Module):
    """InstanceNorm3d(affine=True), no running stats.  modules/discriminator.py:20."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))

    def forward(self, x4):
        return F.instance_norm(x4, None, None, self.weight, self.bias, True, 0.1, 1e-5)


class Down(nn.Module):  # modules/util.py:91-108
    def __init__(self, cin, cout):
        super().__init__()
        self.conv, self.norm = _Conv(cin, cout, 3, 1), _BN(cout)

    def forward(self, x):
        return F.avg_pool2d(F.relu(self.norm(self.conv(x))), 2)


class Up(nn.Module):  # modules/util.py:71-88
    def __init__(self, cin, cout):
        super().__init__()
        self.conv, self.norm = _Conv(cin, cout, 3, 1), _BN(cout)

    def forward(self, x):
        x = F.interpolate(x, scale_factor=2)
        return F.relu(self.norm(self.conv(x)))


class Same(nn.Module):  # modules/util.py:111-126 (grouped 1x1 as used at dense_motion_module.py:26-27)
    def __init__(self, c, groups):
        super().__init__()
        self.conv, self.norm = _Conv(c, c, 1, 0, groups), _BN(c)

    def forward(self, x):
        return F.relu(self.norm(self.conv(x)))


class Res(nn.Module):  # modules/util.py:45-68
    def __init__(self, c):
        super().__init__()
        self.conv1, self.conv2 = _Conv(c, c, 3, 1), _Conv(c, c, 3, 1)
        self.norm1, self.norm2 = _BN(c), _BN(c)

    def forward(self, x):
        y = self.conv1(F.relu(self.norm1(x)))
        y = self.conv2(F.relu(self.norm2(y)))
        return y + x


def _width(be, mf, i):
    return min(mf, be * (2 ** i))


class Enc(nn.Module):  # modules/util.py:129-152
    def __init__(self, be, cin, nb, mf):
        super().__init__()
        self.down_blocks = nn.ModuleList(
            [Down(cin if i == 0 else _width(be, mf, i), _width(be, mf, i + 1)) for i in range(nb)])

    def forward(self, x):
        outs = [x]
        for blk in self.down_blocks:
            outs.append(blk(outs[-1]))
        return outs


class Dec(nn.Module):  # modules/util.py:155-189
    def __init__(self, be, cin, cout, nb, mf, extra=0, last_conv=True):
        super().__init__()
        ups = []
        for i in reversed(range(nb)):
            mult = 1 if i == nb - 1 else 2
            ups.append(Up(mult * _width(be, mf, i + 1) + extra, _width(be, mf, i)))
        self.up_blocks = nn.ModuleList(ups)
        self.conv = _Conv(be + cin + extra, cout, 3, 1) if last_conv else None

    def forward(self, skips):
        skips = list(skips)
        out = skips.pop()
        for blk in self.up_blocks:
            out = torch.cat([blk(out), skips.pop()], dim=1)
        return self.conv(out) if self.conv is not None else out


class HG(nn.Module):  # modules/util.py:192-203
    def __init__(self, be, cin, cout, nb, mf):
        super().__init__()
        self.encoder, self.decoder = Enc(be, cin, nb, mf), Dec(be, cin, cout, nb, mf)

    def forward(self, x):
        return self.decoder(self.encoder(x))


# ----------------------------------------------------------------------------- keypoints
def heat_to_kp(p, kp_variance='matrix', clip_variance=None):
    """Soft-argmax moments of a normalised heatmap p (B,K,D,H,W).  keypoint_detector.py:43-78.
    Returns mean (B,D,K,2), var (B,D,K,2,2) (or (B,D,K,1,1) for 'single')."""
    b, k, d, h, w = p.shape
    p = p + 1e-7  # not renormalised (keypoint_detector.py:49)
    g = coord_grid(h, w, p)  # (h,w,2)
    mean = torch.einsum('bkdhw,hwc->bkdc', p, g)
    out = {'mean': mean.permute(0, 2, 1, 3)}
    if kp_variance == 'matrix':
        dlt = g[None, None, None] - mean[:, :, :, None, None, :]  # (b,k,d,h,w,2)
        var = torch.einsum('bkdhw,bkdhwi,bkdhwj->bkdij', p, dlt, dlt).permute(0, 2, 1, 3, 4)
        if clip_variance:
            sg = sigma_min(var)[..., None, None]
            var = torch.clamp(sg, min=clip_variance) * var / sg
        out['var'] = var
    elif kp_variance == 'single':
        dlt = g[None, None, None] - mean[:, :, :, None, None, :]
        var = (p[..., None] * dlt ** 2).sum(dim=(3, 4)).mean(-1)
        out['var'] = var.permute(0, 2, 1)[..., None, None]
    return out


def kp_to_heat(kp, h, w, kp_variance='matrix'):
    """Render (B,D,K,h,w) gaussians.  keypoint_detector.py:7-40."""
    mean = kp['mean']
    g = coord_grid(h, w, mean)
    dlt = g[None, None, None] - mean[:, :, :, None, None, :]  # (B,D,K,h,w,2)
    if kp_variance == 'matrix':
        iv = inv2x2(kp['var'])  # (B,D,K,2,2)
        q = torch.einsum('bdkhwi,bdkij,bdkhwj->bdkhw', dlt, iv, dlt)
        return torch.exp(-0.5 * q)
    if kp_variance == 'single':
        return torch.exp(-0.5 * (dlt ** 2).sum(-1) / kp['var'])
    return torch.exp(-0.5 * (dlt ** 2).sum(-1) / kp_variance)


class KPDetector(nn.Module):  # keypoint_detector.py:81-109
    def __init__(self, block_expansion, num_kp, num_channels, max_features, num_blocks, temperature,
                 kp_variance, scale_factor=1, clip_variance=None):
        super().__init__()
        self.predictor = HG(block_expansion, num_channels, num_kp, num_blocks, max_features)
        self.temperature, self.kp_variance = temperature, kp_variance
        self.scale_factor, self.clip_variance = scale_factor, clip_variance

    def heatmap(self, x5):
        b = x5.shape[0]
        x = nearest_down(fold(x5), self.scale_factor)
        logits = unfold(self.predictor(x), b)  # (B,K,D,h,w)
        s = logits.shape
        return F.softmax(logits.reshape(s[0], s[1], s[2], -1) / self.temperature, dim=3).reshape(s)

    def forward(self, x5):
        return heat_to_kp(self.heatmap(x5), self.kp_variance, self.clip_variance)


# ----------------------------------------------------------------------------- movement embedding
class MovementEmbedding(nn.Module):  # movement_embedding.py:8-92 (parameter-free)
    def __init__(self, num_kp, kp_variance, num_channels, use_deformed_source_image=False, use_difference=False,
                 use_heatmap=True, add_bg_feature_map=False, heatmap_type='gaussian', norm_const='sum',
                 scale_factor=1):
        super().__init__()
        assert heatmap_type in ('gaussian', 'difference')
        assert use_heatmap or use_deformed_source_image or use_difference
        self.nslots = num_kp + int(add_bg_feature_map)
        self.nfeat = int(use_heatmap) + 2 * int(use_difference) + num_channels * int(use_deformed_source_image)
        self.out_channels = self.nslots * self.nfeat
        self.kp_variance, self.heatmap_type, self.norm_const = kp_variance, heatmap_type, norm_const
        self.use_heatmap, self.use_difference = use_heatmap, use_difference
        self.use_deformed = use_deformed_source_image
        self.bg, self.scale_factor = add_bg_feature_map, scale_factor

    def _norm(self, hm):
        if self.norm_const == 'sum':
            return hm / hm.sum(dim=(3, 4), keepdim=True)
        return hm / self.norm_const

    def forward(self, source_image, kp_driving, kp_source):
        b = source_image.shape[0]
        src = nearest_down(fold(source_image), self.scale_factor)  # (B,C,h,w); source has D=1
        h, w = src.shape[-2:]
        d = kp_driving['mean'].shape[1]
        feats = []
        if self.use_heatmap:
            hm = self._norm(kp_to_heat(kp_driving, h, w, self.kp_variance))
            if self.heatmap_type == 'difference':
                hm = hm - self._norm(kp_to_heat(kp_source, h, w, self.kp_variance))
            if self.bg:
                hm = torch.cat([torch.zeros_like(hm[:, :, :1]), hm], dim=2)
            feats.append(hm[:, :, :, None])  # (B,d,S,1,h,w)
        if self.use_difference or self.use_deformed:
            shift = kp_source['mean'] - kp_driving['mean']  # (B,d,K,2)
            if self.bg:
                shift = torch.cat([torch.zeros_like(shift[:, :, :1]), shift], dim=2)
        if self.use_difference:
            feats.append(shift[..., None, None].expand(b, d, self.nslots, 2, h, w))
        if self.use_deformed:
            g = coord_grid(h, w, src)
            grid = g[None, None, None] + shift[:, :, :, None, None, :]  # (B,d,S,h,w,2)
            img = src[:, None, None].expand(b, d, self.nslots, src.shape[1], h, w)
            warped = F.grid_sample(img.reshape(-1, src.shape[1], h, w), grid.reshape(-1, h, w, 2),
                                   mode='bilinear', padding_mode='zeros', align_corners=True)
            feats.append(warped.reshape(b, d, self.nslots, -1, h, w))
        emb = torch.cat(feats, dim=3).reshape(b, d, -1, h, w)  # slot-major, feature-minor
        return emb.permute(0, 2, 1, 3, 4)  # (B, S*F, d, h, w)


# ----------------------------------------------------------------------------- dense motion
class DenseMotion(nn.Module):  # dense_motion_module.py:8-76
    def __init__(self, block_expansion, num_blocks, max_features, mask_embedding_params, num_kp, num_channels,
                 kp_variance, use_correction, use_mask, bg_init=2, num_group_blocks=0, scale_factor=1):
        super().__init__()
        self.mask_embedding = MovementEmbedding(num_kp=num_kp, kp_variance=kp_variance, num_channels=num_channels,
                                                add_bg_feature_map=True, **mask_embedding_params)
        self.difference_embedding = MovementEmbedding(num_kp=num_kp, kp_variance=kp_variance,
                                                      num_channels=num_channels, add_bg_feature_map=True,
                                                      use_difference=True, use_heatmap=False,
                                                      use_deformed_source_image=False)
        c = self.mask_embedding.out_channels
        self.group_blocks = nn.ModuleList([Same(c, num_kp + 1) for _ in range(num_group_blocks)])
        self.hourglass = HG(block_expansion, c, (num_kp + 1) * use_mask + 2 * use_correction, num_blocks,
                            max_features)
        self.num_kp, self.use_correction, self.use_mask = num_kp, use_correction, use_mask
        self.scale_factor = scale_factor

    def forward(self, source_image, kp_driving, kp_source):
        b = source_image.shape[0]
        if self.scale_factor != 1:
            source_image = unfold(nearest_down(fold(source_image), self.scale_factor), b)
        emb = self.mask_embedding(source_image, kp_driving, kp_source)  # (B,C,d,h,w)
        x = fold(emb)
        for blk in self.group_blocks:
            x = F.leaky_relu(blk(x), 0.2)
        pred = unfold(self.hourglass(x), b)  # (B,P,d,h,w)
        _, _, d, h, w = pred.shape
        flow = 0
        if self.use_mask:
            mask = F.softmax(pred[:, :self.num_kp + 1], dim=1)  # (B,K+1,d,h,w)
            shift = self.difference_embedding(source_image, kp_driving, kp_source)
            shift = shift.reshape(b, self.num_kp + 1, 2, d, h, w)
            flow = (shift * mask[:, :, None]).sum(dim=1)  # (B,2,d,h,w)
        if self.use_correction:
            flow = flow + pred[:, -2:]
        flow = flow.permute(0, 2, 3, 4, 1) + coord_grid(h, w, pred)[None, None]
        return torch.cat([flow, torch.zeros_like(flow[..., :1])], dim=-1)  # (B,d,h,w,3)


class IdentityDeformation(nn.Module):  # dense_motion_module.py:79-87
    def forward(self, source_image, kp_driving, kp_source):
        b, _, _, h, w = source_image.shape
        d = kp_driving['mean'].shape[1]
        g = coord_grid(h, w, source_image)[None, None].expand(b, d, h, w, 2)
        return torch.cat([g, torch.zeros_like(g[..., :1])], dim=-1)


# ----------------------------------------------------------------------------- generator
class Generator(nn.Module):  # generator.py:10-82
    def __init__(self, num_channels, num_kp, kp_variance, block_expansion, max_features, num_blocks,
                 num_refinement_blocks, dense_motion_params=None, kp_embedding_params=None,
                 interpolation_mode='nearest'):
        super().__init__()
        self.appearance_encoder = Enc(block_expansion, num_channels, num_blocks, max_features)
        if kp_embedding_params is not None:
            self.kp_embedding_module = MovementEmbedding(num_kp=num_kp, kp_variance=kp_variance,
                                                         num_channels=num_channels, **kp_embedding_params)
            extra = self.kp_embedding_module.out_channels
        else:
            self.kp_embedding_module, extra = None, 0
        if dense_motion_params is not None:
            self.dense_motion_module = DenseMotion(num_kp=num_kp, kp_variance=kp_variance,
                                                   num_channels=num_channels, **dense_motion_params)
        else:
            self.dense_motion_module = IdentityDeformation()
        self.video_decoder = Dec(block_expansion, num_channels, num_channels, num_blocks, max_features,
                                 extra=extra, last_conv=False)
        c = block_expansion + num_channels + extra
        self.refinement_module = nn.Sequential()
        for i in range(num_refinement_blocks):
            self.refinement_module.add_module('r%d' % i, Res(c))
        self.refinement_module.add_module('conv-last', _Conv(c, num_channels, 1, 0))
        self.interpolation_mode = interpolation_mode

    def _resize(self, x4, h, w):
        """F.interpolate(size=(d,h,w), mode) with d preserved: nearest, or trilinear==bilinear
        (align_corners=False).  generator.py:55,72."""
        if self.interpolation_mode == 'nearest':
            return F.interpolate(x4, size=(h, w), mode='nearest')
        return F.interpolate(x4, size=(h, w), mode='bilinear', align_corners=False)

    def warp(self, skip4, deform, d):
        """skip4 (B,C,h,w); deform (B,d,h0,w0,3) -> (B*d,C,h,w).  generator.py:51-58."""
        b, c, h, w = skip4.shape
        g = deform[..., :2].reshape(b * d, deform.shape[2], deform.shape[3], 2).permute(0, 3, 1, 2)
        g = self._resize(g, h, w).permute(0, 2, 3, 1)
        src = skip4[:, None].expand(b, d, c, h, w).reshape(b * d, c, h, w)
        return F.grid_sample(src, g, mode='bilinear', padding_mode='zeros', align_corners=True)

    def forward(self, source_image, kp_driving, kp_source):
        b = source_image.shape[0]
        d = kp_driving['mean'].shape[1]
        skips = self.appearance_encoder(fold(source_image))
        deform = self.dense_motion_module(source_image=source_image, kp_driving=kp_driving, kp_source=kp_source)
        warped = [self.warp(s, deform, d) for s in skips]
        if self.kp_embedding_module is not None:
            emb = fold(self.kp_embedding_module(source_image=source_image, kp_driving=kp_driving,
                                                kp_source=kp_source))
            warped = [torch.cat([s, self._resize(emb, s.shape[2], s.shape[3])], dim=1) for s in warped]
        video_deformed = self.warp(fold(source_image), deform, d)
        out = self.refinement_module(self.video_decoder(warped))
        return {'video_prediction': unfold(torch.sigmoid(out), b), 'video_deformed': unfold(video_deformed, b)}


# ----------------------------------------------------------------------------- discriminator
class _DBlock(nn.Module):  # modules/discriminator.py:7-31
    def __init__(self, cin, cout, norm):
        super().__init__()
        self.conv = _Conv(cin, cout, 4, 0)
        self.norm = _IN(cout) if norm else None

    def forward(self, x):
        y = self.conv(x)
        if self.norm is not None:
            y = self.norm(y)
        return F.avg_pool2d(F.leaky_relu(y, 0.2), 2)


class Discriminator(nn.Module):  # modules/discriminator.py:34-79
    def __init__(self, num_channels=3, num_kp=10, kp_variance=0.01, scale_factor=1, block_expansion=64,
                 num_blocks=4, max_features=512, kp_embedding_params=None):
        super().__init__()
        if kp_embedding_params is not None:
            self.kp_embedding = MovementEmbedding(num_kp=num_kp, kp_variance=kp_variance,
                                                  num_channels=num_channels, **kp_embedding_params)
            extra = self.kp_embedding.out_channels
        else:
            self.kp_embedding, extra = None, 0
        self.down_blocks = nn.ModuleList([
            _DBlock(num_channels + extra if i == 0 else _width(block_expansion, max_features, i),
                    _width(block_expansion, max_features, i + 1), norm=(i != 0)) for i in range(num_blocks)])
        self.conv = _Conv(_width(block_expansion, max_features, num_blocks), 1, 1, 0)
        self.scale_factor = scale_factor

    def forward(self, x, kp_driving, kp_source):
        b = x.shape[0]
        maps = [x]
        if self.scale_factor != 1:
            x = unfold(nearest_down(fold(x), self.scale_factor), b)
        out = fold(x)
        if self.kp_embedding is not None:
            out = torch.cat([out, fold(self.kp_embedding(x, kp_driving, kp_source))], dim=1)
        for blk in self.down_blocks:
            out = blk(out)
            maps.append(unfold(out, b))
        maps.append(unfold(self.conv(out), b))
        return maps


# ----------------------------------------------------------------------------- losses / full models
def mean_batch(v):  # modules/losses.py:4-5
    return v.reshape(v.shape[0], -1).mean(-1)


def generator_loss(maps_gen, maps_real, video_deformed, loss_weights):  # modules/losses.py:46-60
    vals = []
    if loss_weights['reconstruction_deformed'] != 0:
        vals.append(loss_weights['reconstruction_deformed'] * mean_batch((maps_real[0] - video_deformed).abs()))
    if loss_weights['reconstruction'] != 0:
        for wgt, a, g in zip(loss_weights['reconstruction'], maps_real[:-1], maps_gen[:-1]):
            if wgt != 0:
                vals.append(wgt * mean_batch((g - a).abs()))
    vals.append(loss_weights['generator_gan'] * mean_batch((1 - maps_gen[-1]) ** 2))
    return vals


def discriminator_loss(maps_gen, maps_real, loss_weights):  # modules/losses.py:63-67
    return [loss_weights['discriminator_gan'] * mean_batch((1 - maps_real[-1]) ** 2 + maps_gen[-1] ** 2)]


def split_kp(kp_joined, detach=False):  # train.py:14-21
    f = (lambda t: t.detach()) if detach else (lambda t: t)
    return {'kp_driving': {k: f(v[:, 1:]) for k, v in kp_joined.items()},
            'kp_source': {k: f(v[:, :1]) for k, v in kp_joined.items()}}


def generator_full(kp_detector, generator, discriminator, train_params, x):  # train.py:36-53
    kp_joined = kp_detector(torch.cat([x['source'], x['video']], dim=2))
    generated = generator(x['source'], **split_kp(kp_joined, train_params['detach_kp_generator']))
    kp_dict = split_kp(kp_joined, False)
    maps_gen = discriminator(generated['video_prediction'], **kp_dict)
    maps_real = discriminator(x['video'], **kp_dict)
    generated.update(kp_dict)
    losses = generator_loss(maps_gen, maps_real, generated['video_deformed'], train_params['loss_weights'])
    return tuple(losses) + (generated, kp_joined)


def discriminator_full(kp_detector, generator, discriminator, train_params, x, kp_joined, generated):  # train.py:68-75
    kp_dict = split_kp(kp_joined, train_params['detach_kp_discriminator'])
    maps_gen = discriminator(generated['video_prediction'].detach(), **kp_dict)
    maps_real = discriminator(x['video'], **kp_dict)
    return discriminator_loss(maps_gen, maps_real, train_params['loss_weights'])


def make_optimizers(generator, discriminator, kp_detector, lr):  # train.py:81-83
    mk = lambda m: torch.optim.Adam(m.parameters(), lr=lr, betas=(0.5, 0.999))
    return mk(generator), mk(discriminator), mk(kp_detector)


def train_iteration(kp_detector, generator, discriminator, optimizers, train_params, x):
    """One body of the training loop, train.py:110-136 (logging excluded).  Returns the loss scalars."""
    opt_g, opt_d, opt_kp = optimizers
    out = generator_full(kp_detector, generator, discriminator, train_params, x)
    g_vals = [v.mean() for v in out[:-2]]
    generated, kp_joined = out[-2], out[-1]
    sum(g_vals).backward(retain_graph=not train_params['detach_kp_discriminator'])
    opt_g.step(); opt_g.zero_grad(); opt_d.zero_grad()
    if train_params['detach_kp_discriminator']:
        opt_kp.step(); opt_kp.zero_grad()
    d_vals = [v.mean() for v in discriminator_full(kp_detector, generator, discriminator, train_params, x,
                                                    kp_joined, generated)]
    sum(d_vals).backward()
    opt_d.step(); opt_d.zero_grad()
    if not train_params['detach_kp_discriminator']:
        opt_kp.step(); opt_kp.zero_grad()
    return [float(v) for v in g_vals], [float(v) for v in d_vals]


# ----------------------------------------------------------------------------- inference drivers
def cat_kp(kps, dim):
    return {k: torch.cat([kp[k] for kp in kps], dim=dim) for k in kps[0]}


def normalize_kp(kp_video, kp_source, move_location=False, movement_mult=False, adapt_variance=False,
                 clip_mean=False):
    """transfer.py:31-62, the options the shipped configs exercise (move_location only)."""
    assert not movement_mult and not adapt_variance and not clip_mean, 'oracle covers move_location only'
    kp_video = dict(kp_video)
    if move_location:
        kp_video['mean'] = kp_video['mean'] - kp_video['mean'][:, 0:1] + kp_source['mean']
    return kp_video


def transfer_one(generator, kp_detector, source_image, driving_video, normalization_params):
    """transfer.py:65-79: one KP pass per driving frame, one generator pass per frame."""
    d = driving_video.shape[2]
    kp_driving = cat_kp([kp_detector(driving_video[:, :, i:i + 1]) for i in range(d)], dim=1)
    kp_source = kp_detector(source_image)
    kp_norm = normalize_kp(kp_driving, kp_source, **normalization_params)
    outs = [generator(source_image=source_image, kp_driving={k: v[:, i:i + 1] for k, v in kp_norm.items()},
                      kp_source=kp_source) for i in range(d)]
    out = {k: torch.cat([o[k] for o in outs], dim=2) for k in outs[0]}
    out.update(kp_driving=kp_driving, kp_source=kp_source, kp_norm=kp_norm)
    return out


def reconstruct(generator, kp_detector, video):
    """reconstruction.py:12-25,57-62: frame 0 is the source, every frame is a driving frame."""
    d = video.shape[2]
    kp_source = kp_detector(video[:, :, :1])
    kp_video = cat_kp([kp_detector(video[:, :, i:i + 1]) for i in range(d)], dim=1)
    outs = [generator(video[:, :, :1], kp_driving={k: v[:, i:i + 1] for k, v in kp_video.items()},
                      kp_source=kp_source) for i in range(d)]
    out = {k: torch.cat([o[k] for o in outs], dim=2) for k in outs[0]}
    out.update(kp_driving=kp_video, kp_source=kp_source)
    return out


def build_from_config(config):
    """Same kwargs-splat as run.py:50-63."""
    mp = config['model_params']
    gen = Generator(**mp['generator_params'], **mp['common_params'])
    disc = Discriminator(**mp['discriminator_params'], **mp['common_params'])
    kp = KPDetector(**mp['kp_detector_params'], **mp['common_params'])
    return gen, disc, kp


def conv_flops(module, *args, **kwargs):
    """Algorithmic conv FLOPs (2*MACs) of one forward of `module`, by hooking every conv."""
    total = [0]
    hooks = []

    def hook(m, inp, out):
        co, cig, _, kh, kw = m.weight.shape
        total[0] += 2 * out.numel() * cig * kh * kw

    for m in module.modules():
        if isinstance(m, _Conv):
            hooks.append(m.register_forward_hook(hook))
    with torch.no_grad():
        module(*args, **kwargs)
    for hh in hooks:
        hh.remove()
    return total[0]
