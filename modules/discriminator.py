"""Discriminator - drop-in for the reference's `modules/discriminator.py` (Pix2Pix style; returns ALL feature maps
for the feature-matching loss).  4x4 valid convs on the implicit-GEMM kernel, InstanceNorm + LeakyReLU(0.2) +
avgpool(1,2,2) fused in the normalisation kernels; the returned maps are zero-copy (B,C,1,H,W) views of the NHWC
buffers, which `modules.losses` consumes stride-aware."""
from torch import nn

from modules.movement_embedding import MovementEmbeddingModule
from modules.keypoint_detector import _step
from monkey_net_b200 import ops


class DownBlock3D(nn.Module):
    """conv4x4 (valid) -> [InstanceNorm] -> LeakyReLU(0.2) -> avgpool (reference discriminator.py:7-31)."""

    def __init__(self, in_features, out_features, norm=False, kernel_size=4):
        super(DownBlock3D, self).__init__()
        self.conv = nn.Conv3d(in_channels=in_features, out_channels=out_features,
                              kernel_size=(1, kernel_size, kernel_size))
        if norm:
            self.norm = nn.InstanceNorm3d(out_features, affine=True)
        else:
            self.norm = None

    def run(self, a):
        y = ops.conv(a, self.conv.weight, self.conv.bias, pad=0, feeds_train_norm=self.norm is not None)
        return ops.norm_act(y, self.norm, mode='in' if self.norm is not None else 'none', slope=0.2, pool=1)

    def forward(self, x):
        return ops.from_nhwc(self.run(ops.to_nhwc(x)), x.shape[0])


class Discriminator(nn.Module):
    def __init__(self, num_channels=3, num_kp=10, kp_variance=0.01, scale_factor=1,
                 block_expansion=64, num_blocks=4, max_features=512, kp_embedding_params=None):
        super(Discriminator, self).__init__()
        if kp_embedding_params is not None:
            self.kp_embedding = MovementEmbeddingModule(num_kp=num_kp, kp_variance=kp_variance,
                                                        num_channels=num_channels, **kp_embedding_params)
            embedding_channels = self.kp_embedding.out_channels
        else:
            self.kp_embedding = None
            embedding_channels = 0
        down_blocks = []
        for i in range(num_blocks):
            down_blocks.append(DownBlock3D(
                num_channels + embedding_channels if i == 0 else min(max_features, block_expansion * (2 ** i)),
                min(max_features, block_expansion * (2 ** (i + 1))), norm=(i != 0), kernel_size=4))
        self.down_blocks = nn.ModuleList(down_blocks)
        self.conv = nn.Conv3d(self.down_blocks[-1].conv.out_channels, out_channels=1, kernel_size=1)
        self.scale_factor = scale_factor

    def forward(self, x, kp_driving, kp_source):
        b = x.shape[0]
        out_maps = [x]
        a = ops.to_nhwc(x, _step(self.scale_factor))
        if self.kp_embedding is not None:
            src = a if self.kp_embedding.use_deformed_source_image else None
            if _step(self.kp_embedding.scale_factor) != 1:
                raise NotImplementedError('discriminator kp_embedding scale_factor != 1 is unused by the configs')
            emb = self.kp_embedding.run(src, a.shape[1], a.shape[2], kp_driving, kp_source)
            a = ops.concat([a, emb])
        for down_block in self.down_blocks:
            a = down_block.run(a)
            out_maps.append(ops.from_nhwc(a, b))
        score = ops.conv(a, self.conv.weight, self.conv.bias, pad=0)
        out_maps.append(ops.from_nhwc(score, b))
        return out_maps
