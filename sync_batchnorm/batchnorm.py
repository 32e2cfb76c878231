"""SynchronizedBatchNorm{1,2,3}d with the reference's constructor and state_dict (weight, bias, running_mean,
running_var, num_batches_tracked), computing with libmonkey_b200 kernels.

Semantics: `F.batch_norm` on the GLOBAL batch (what the reference computes on one device,
sync_batchnorm/batchnorm.py:50-53): biased variance + eps to normalise, unbiased variance into running_var,
momentum 0.1.  The reference's multi-device branch uses `clamp(var, eps) ** -0.5` (batchnorm.py:125); we keep
the single-device formula on every rank so that N-GPU == 1-GPU (deviation noted in SURVEY 7.3).
Inside `modules/*` the normalisation is fused with the activation / pooling kernels and these objects only hold the
parameters; calling the module directly normalises an (N,C,...) tensor on its own.
"""
import torch
from torch.nn.modules.batchnorm import _BatchNorm

from monkey_net_b200 import ops


class _SynchronizedBatchNorm(_BatchNorm):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True):
        super(_SynchronizedBatchNorm, self).__init__(num_features, eps=eps, momentum=momentum, affine=affine)
        if eps != 1e-5 or momentum != 0.1 or not affine:
            raise NotImplementedError('B200 path implements the configuration the reference uses: eps=1e-5, '
                                      'momentum=0.1, affine=True')

    def forward(self, input):
        self._check_input_dim(input)
        shape = input.shape
        x5 = input.reshape(shape[0], shape[1], 1, -1, 1) if input.dim() != 5 else input
        a = ops.to_nhwc(x5)
        y = ops.norm_act(a, self, mode='bn', slope=-1.0)
        out = ops.from_nhwc(y, x5.shape[0])
        return out.reshape(shape) if input.dim() != 5 else out


class SynchronizedBatchNorm1d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() not in (2, 3):
            raise ValueError('expected 2D or 3D input (got {}D input)'.format(input.dim()))


class SynchronizedBatchNorm2d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 4:
            raise ValueError('expected 4D input (got {}D input)'.format(input.dim()))


class SynchronizedBatchNorm3d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 5:
            raise ValueError('expected 5D input (got {}D input)'.format(input.dim()))
