"""Movement embedding - drop-in for the reference's `modules/movement_embedding.py` (parameter-free).

One fused kernel renders, per keypoint slot, [normalised (driving - source) gaussian heatmap | keypoint shift |
source image translated by the shift] straight into the slot-major / feature-minor channel layout the grouped 1x1
convs rely on (reference movement_embedding.py:51-92); kp2gaussian's 2x2 inverse is closed form, the 4-D
F.grid_sample of the reference (line 85) is a constant-shift bilinear fetch inside the same kernel.
"""
from torch import nn

from monkey_net_b200 import ops
from modules.keypoint_detector import _step


class MovementEmbeddingModule(nn.Module):
    """Produce a keypoint representation that will be further used by other modules."""

    def __init__(self, num_kp, kp_variance, num_channels, use_deformed_source_image=False, use_difference=False,
                 use_heatmap=True, add_bg_feature_map=False, heatmap_type='gaussian', norm_const='sum', scale_factor=1):
        super(MovementEmbeddingModule, self).__init__()
        assert heatmap_type in ['gaussian', 'difference']
        assert ((int(use_heatmap) + int(use_deformed_source_image) + int(use_difference)) >= 1)
        self.out_channels = (1 * use_heatmap + 2 * use_difference + num_channels * use_deformed_source_image) * (
            num_kp + add_bg_feature_map)
        self.kp_variance = kp_variance
        self.heatmap_type = heatmap_type
        self.use_difference = use_difference
        self.use_deformed_source_image = use_deformed_source_image
        self.use_heatmap = use_heatmap
        self.add_bg_feature_map = add_bg_feature_map
        self.norm_const = norm_const
        self.scale_factor = scale_factor
        self.num_channels = num_channels

    def run(self, src, h, w, kp_driving, kp_source):
        """src: Act of the (already down-scaled) source image or None when it is not sampled; (h,w) output size."""
        return ops.movement_embed(src, kp_driving, kp_source, h, w, num_channels=self.num_channels,
                                  kp_variance=self.kp_variance, use_heatmap=self.use_heatmap,
                                  use_difference=self.use_difference, use_deformed=self.use_deformed_source_image,
                                  add_bg=self.add_bg_feature_map, heatmap_type=self.heatmap_type,
                                  norm_const=self.norm_const)

    def run_from_image(self, source_image, kp_driving, kp_source):
        step = _step(self.scale_factor)
        h, w = source_image.shape[3] // step, source_image.shape[4] // step
        src = ops.to_nhwc(source_image, step) if self.use_deformed_source_image else None
        return self.run(src, h, w, kp_driving, kp_source)

    def forward(self, source_image, kp_driving, kp_source):
        a = self.run_from_image(source_image, kp_driving, kp_source)
        return ops.from_nhwc(a, source_image.shape[0])  # (B, out_channels, d, h, w)
