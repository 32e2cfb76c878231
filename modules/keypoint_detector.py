"""Keypoint detector - drop-in for the reference's `modules/keypoint_detector.py`.

KPDetector.forward(x) -> {'mean': (B,D,K,2), 'var': (B,D,K,2,2)}: [nearest down-scale] -> Hourglass -> spatial
softmax with temperature -> soft-argmax moments.  The hourglass runs on the conv kernels; softmax + gaussian2kp
(reference keypoint_detector.py:43-78,101-107) is ONE block-reduction kernel per (frame, keypoint) with a
hand-written backward, and the down-scale is folded into the NCDHW->NHWC layout kernel.
"""
from torch import nn

from modules.util import Hourglass
from monkey_net_b200 import ops


def _step(scale_factor):
    if scale_factor == 1:
        return 1
    step = int(round(1.0 / scale_factor))
    if step < 1 or abs(step * scale_factor - 1.0) > 1e-6:
        raise NotImplementedError('scale_factor must be 1/integer (configs use 1, 0.5, 0.25), got %r' % scale_factor)
    return step


def kp2gaussian(kp, spatial_size, kp_variance='matrix'):
    """Gaussian heatmaps (B,D,K,h,w) of a keypoint dict (reference keypoint_detector.py:7-40)."""
    h, w = spatial_size
    mean = kp['mean']
    lead = mean.shape[:-2]
    kp4 = {k: v.reshape((-1, 1) + v.shape[len(lead):]) for k, v in kp.items()}  # (B*,1,K,...) : d = 1 frames
    a = ops.movement_embed(None, kp4, kp4, h, w, num_channels=0, kp_variance=kp_variance, use_heatmap=True,
                           use_difference=False, use_deformed=False, add_bg=False, heatmap_type='gaussian',
                           norm_const=1.0)
    k = mean.shape[-2]
    out = ops.from_nhwc(a, kp4['mean'].shape[0])  # (B*, K, 1, h, w)
    return out[:, :, 0].reshape(lead + (k, h, w))


def gaussian2kp(heatmap, kp_variance='matrix', clip_variance=None):
    """Moments of an already-normalised heatmap (B,K,D,H,W) (reference keypoint_detector.py:43-78).  The fused
    kernel takes logits, so this helper feeds log(heatmap) at temperature 1 (softmax(log p) == p for normalised p)."""
    b, k, d = heatmap.shape[:3]
    a = ops.to_nhwc(heatmap.clamp_min(1e-30).log())
    return ops.kp_head(a, b, d, k, 1.0, kp_variance, clip_variance)


class KPDetector(nn.Module):
    """Detecting keypoints. Returns keypoint position and variance (reference keypoint_detector.py:81-109)."""

    def __init__(self, block_expansion, num_kp, num_channels, max_features, num_blocks, temperature,
                 kp_variance, scale_factor=1, clip_variance=None):
        super(KPDetector, self).__init__()
        self.predictor = Hourglass(block_expansion, in_features=num_channels, out_features=num_kp,
                                   max_features=max_features, num_blocks=num_blocks)
        self.temperature = temperature
        self.kp_variance = kp_variance
        self.scale_factor = scale_factor
        self.clip_variance = clip_variance
        self.num_kp = num_kp

    def forward(self, x):
        b, _, d = x.shape[:3]
        a = ops.to_nhwc(x, _step(self.scale_factor))
        if ops.KP_PRECISE:
            with ops.reference_precision():   # bit-exact keypoint pixel indices need fp32-accurate convolutions
                logits = self.predictor.run(a)
        else:
            logits = self.predictor.run(a)
        return ops.kp_head(logits, b, d, self.num_kp, self.temperature, self.kp_variance, self.clip_variance)
