"""Motion transfer generator - drop-in for the reference's `modules/generator.py`.

appearance Encoder -> dense motion -> every skip warped by ONE grid_sample kernel per level that reads the coarse
deformation field directly (the reference materialises a resized (B,d,h,w,3) grid per level, generator.py:51-58) ->
kp-embedding resized per level -> Decoder (concats written in place by the BN-apply kernel) -> ResBlocks -> 1x1 conv
with the sigmoid fused in its epilogue.  `video_deformed` is the level-0 warp (the reference computes the same
tensor twice, generator.py:66,77).
"""
import torch
from torch import nn

from modules.util import Encoder, Decoder, ResBlock3D
from modules.dense_motion_module import DenseMotionModule, IdentityDeformation
from modules.movement_embedding import MovementEmbeddingModule
from monkey_net_b200 import ops


class MotionTransferGenerator(nn.Module):
    """Given keypoints and an appearance frame reconstruct the target frame; returns the warped source
    (`video_deformed`) and the refined prediction (`video_prediction`), both (B,C,d,H,W)."""

    def __init__(self, num_channels, num_kp, kp_variance, block_expansion, max_features, num_blocks,
                 num_refinement_blocks, dense_motion_params=None, kp_embedding_params=None,
                 interpolation_mode='nearest'):
        super(MotionTransferGenerator, self).__init__()
        self.appearance_encoder = Encoder(block_expansion, in_features=num_channels, max_features=max_features,
                                          num_blocks=num_blocks)
        if kp_embedding_params is not None:
            self.kp_embedding_module = MovementEmbeddingModule(num_kp=num_kp, kp_variance=kp_variance,
                                                               num_channels=num_channels, **kp_embedding_params)
            embedding_features = self.kp_embedding_module.out_channels
        else:
            self.kp_embedding_module = None
            embedding_features = 0
        if dense_motion_params is not None:
            self.dense_motion_module = DenseMotionModule(num_kp=num_kp, kp_variance=kp_variance,
                                                         num_channels=num_channels, **dense_motion_params)
        else:
            self.dense_motion_module = IdentityDeformation()
        self.video_decoder = Decoder(block_expansion=block_expansion, in_features=num_channels,
                                     out_features=num_channels, max_features=max_features, num_blocks=num_blocks,
                                     additional_features_for_block=embedding_features, use_last_conv=False)
        self.refinement_module = torch.nn.Sequential()
        in_features = block_expansion + num_channels + embedding_features
        for i in range(num_refinement_blocks):
            self.refinement_module.add_module('r' + str(i),
                                              ResBlock3D(in_features, kernel_size=(1, 3, 3), padding=(0, 1, 1)))
        self.refinement_module.add_module('conv-last', nn.Conv3d(in_features, num_channels, kernel_size=1, padding=0))
        if interpolation_mode not in ('nearest', 'trilinear'):
            raise NotImplementedError('interpolation_mode %r (configs use nearest / trilinear)' % interpolation_mode)
        self.interpolation_mode = interpolation_mode

    def deform_input(self, inp, deformations_absolute):
        """Reference-layout helper (generator.py:51-58): inp (B,C,1,h,w), deformation (B,d,h0,w0,3)."""
        b, d, h0, w0, _ = deformations_absolute.shape
        deform = deformations_absolute[..., :2].reshape(b * d, h0, w0, 2).contiguous()
        return ops.from_nhwc(ops.grid_sample(ops.to_nhwc(inp), deform, d, self.interpolation_mode), b)

    def forward(self, source_image, kp_driving, kp_source):
        b = source_image.shape[0]
        d = kp_driving['mean'].shape[1]
        skips = self.appearance_encoder.run(ops.to_nhwc(source_image))
        deform = self.dense_motion_module.run(source_image, kp_driving, kp_source)
        warped = [ops.grid_sample(s, deform, d, self.interpolation_mode) for s in skips]
        if self.kp_embedding_module is not None:
            emb = self.kp_embedding_module.run_from_image(source_image, kp_driving, kp_source)
            levels = [[s, ops.resize(emb, s.shape[1], s.shape[2], self.interpolation_mode)] for s in warped]
        else:
            levels = [[s] for s in warped]
        out = ops.compact(self.video_decoder.run(levels))
        for name, layer in self.refinement_module.named_children():
            if name == 'conv-last':
                out = ops.conv(out, layer.weight, layer.bias, pad=0, act='sigmoid')
            else:
                out = layer.run(out)
        return {"video_prediction": ops.from_nhwc(out, b), "video_deformed": ops.from_nhwc(warped[0], b)}
