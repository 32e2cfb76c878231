"""Drop-in facade for the reference's `sync_batchnorm` package (sync_batchnorm/__init__.py:11-12): the same names,
backed by the B200 kernels and - across GPUs - by ONE NCCL all-reduce of the packed BN statistics per layer per
direction instead of the reference's thread rendezvous + reduce/broadcast pair."""
from .batchnorm import SynchronizedBatchNorm1d, SynchronizedBatchNorm2d, SynchronizedBatchNorm3d
from .replicate import DataParallelWithCallback, patch_replication_callback

__all__ = ['SynchronizedBatchNorm1d', 'SynchronizedBatchNorm2d', 'SynchronizedBatchNorm3d',
           'DataParallelWithCallback', 'patch_replication_callback']
