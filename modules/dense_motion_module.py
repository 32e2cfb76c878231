"""Dense motion module - drop-in for the reference's `modules/dense_motion_module.py`.

mask embedding -> grouped 1x1 SameBlocks -> Hourglass -> [channel softmax over K+1 masks x keypoint shifts +
correction + identity grid] (reference dense_motion_module.py:42-76).  The head is one fused kernel producing the
(x,y) deformation; the all-zero z column of the reference's (B,d,h,w,3) output exists only at this module's public
`forward` edge (inside the generator the 2-channel field feeds the grid_sample kernel directly).
"""
import contextlib

import torch
from torch import nn

from modules.util import Hourglass, SameBlock3D, make_coordinate_grid
from modules.movement_embedding import MovementEmbeddingModule
from modules.keypoint_detector import _step
from monkey_net_b200 import ops


class DenseMotionModule(nn.Module):
    """Predicts a dense optical flow from the displacement of keypoints and the appearance of the first frame."""

    def __init__(self, block_expansion, num_blocks, max_features, mask_embedding_params, num_kp,
                 num_channels, kp_variance, use_correction, use_mask, bg_init=2, num_group_blocks=0, scale_factor=1):
        super(DenseMotionModule, self).__init__()
        self.mask_embedding = MovementEmbeddingModule(num_kp=num_kp, kp_variance=kp_variance, num_channels=num_channels,
                                                      add_bg_feature_map=True, **mask_embedding_params)
        self.difference_embedding = MovementEmbeddingModule(num_kp=num_kp, kp_variance=kp_variance,
                                                            num_channels=num_channels,
                                                            add_bg_feature_map=True, use_difference=True,
                                                            use_heatmap=False, use_deformed_source_image=False)
        group_blocks = []
        for i in range(num_group_blocks):
            group_blocks.append(SameBlock3D(self.mask_embedding.out_channels, self.mask_embedding.out_channels,
                                            groups=num_kp + 1, kernel_size=(1, 1, 1), padding=(0, 0, 0)))
        self.group_blocks = nn.ModuleList(group_blocks)

        self.hourglass = Hourglass(block_expansion=block_expansion, in_features=self.mask_embedding.out_channels,
                                   out_features=(num_kp + 1) * use_mask + 2 * use_correction,
                                   max_features=max_features, num_blocks=num_blocks)
        # zero-initialised head: flow == identity, mask == softmax([bg_init, 0, ...]) (reference lines 33-35)
        self.hourglass.decoder.conv.weight.data.zero_()
        bias_init = ([bg_init] + [0] * num_kp) * use_mask + [0, 0] * use_correction
        self.hourglass.decoder.conv.bias.data.copy_(torch.tensor(bias_init, dtype=torch.float))

        self.num_kp = num_kp
        self.use_correction = use_correction
        self.use_mask = use_mask
        self.scale_factor = scale_factor

    def run(self, source_image, kp_driving, kp_source):
        """-> deformation tensor [B*d, h, w, 2] in normalised (x, y)."""
        step = _step(self.scale_factor)
        h, w = source_image.shape[3] // step, source_image.shape[4] // step
        src = ops.to_nhwc(source_image, step) if self.mask_embedding.use_deformed_source_image else None
        x = self.mask_embedding.run(src, h, w, kp_driving, kp_source)
        # The deformation field is GEOMETRY, like the keypoints: a 1e-4 error of a sampling coordinate is multiplied by
        # the image gradient (unbounded at the zero-padding border of grid_sample), measured 2e-3 ... 2e-2 on
        # `video_deformed` with 1xTF32 convolutions.  Under the 'auto' policy this network therefore always runs
        # fp32-accurately (3xTF32); the appearance encoder / decoder keep 1xTF32 for no_grad inference.
        with (ops.reference_precision() if ops.KP_PRECISE else contextlib.nullcontext()):
            for block in self.group_blocks:
                x = block.run(x)  # F.leaky_relu(relu(.), 0.2) of reference line 49 is the identity on a ReLU output
            pred = self.hourglass.run(x)
        return ops.flow_head(pred, kp_driving, kp_source, self.use_mask, self.use_correction)

    def forward(self, source_image, kp_driving, kp_source):
        b = source_image.shape[0]
        deform = self.run(source_image, kp_driving, kp_source)
        n, h, w, _ = deform.shape
        deform = deform.view(b, n // b, h, w, 2)
        return torch.cat([deform, torch.zeros_like(deform[..., :1])], dim=-1)


class IdentityDeformation(nn.Module):
    def run(self, source_image, kp_driving, kp_source):
        b, _, _, h, w = source_image.shape
        d = kp_driving['mean'].shape[1]
        grid = make_coordinate_grid((h, w), type=source_image.type())
        return grid.view(1, h, w, 2).repeat(b * d, 1, 1, 1)

    def forward(self, appearance_frame, kp_video, kp_appearance):
        b = appearance_frame.shape[0]
        grid = self.run(appearance_frame, kp_video, kp_appearance)
        n, h, w, _ = grid.shape
        grid = grid.view(b, n // b, h, w, 2)
        return torch.cat([grid, torch.zeros_like(grid[..., :1])], dim=-1)
