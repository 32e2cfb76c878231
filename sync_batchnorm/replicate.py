"""`DataParallelWithCallback` facade (reference: sync_batchnorm/replicate.py:50-67, used at train.py:104-105,
transfer.py:101-102, reconstruction.py:45-46).

B200 design: ONE PROCESS PER GPU (torchrun), not one Python thread per GPU.  The wrapper therefore holds exactly one
replica - the wrapped module itself, on this rank's device - and
  * moves CPU tensors handed to it onto that device (DataParallel's scatter did the H2D copy, train.py:110);
  * when torch.distributed is initialised, registers gradient-averaging hooks on the wrapped parameters so the
    unchanged `loss.backward(); optimizer.step()` of train.py:117-136 trains data-parallel (the batch each rank
    feeds is its own shard); BN statistics are synchronised inside the norm kernels' host code
    (monkey_net_b200/dist.py);
  * `device_ids` with more than one entry in a single process is rejected with an explanation.
"""
import torch
from torch import nn
import torch.distributed as dist

_hooked = set()
_synced = set()
HOOKS_ENABLED = True   # monkey_net_b200.train_step.GraphedTrainer switches to one flat all-reduce per optimiser step


def _to_device(obj, device):
    if torch.is_tensor(obj):
        return obj.to(device, non_blocking=True) if obj.device != device else obj
    if isinstance(obj, dict):
        return {k: _to_device(v, device) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_device(v, device) for v in obj)
    return obj


def _avg_hook(p):
    if not HOOKS_ENABLED:
        return
    w = dist.get_world_size()
    dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
    p.grad.div_(w)


def broadcast_module_state(module, src=0):
    """rank `src`'s parameters and buffers -> every rank, in place (each tensor once per process)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            if id(t) in _synced:
                continue
            _synced.add(id(t))
            if t.dtype == torch.int64 and t.dim() == 0:   # num_batches_tracked: nccl/gloo want >= 1-D
                buf = t.detach().reshape(1).clone()
                dist.broadcast(buf, src=src)
                t.copy_(buf[0])
            else:
                dist.broadcast(t.data, src=src)


class DataParallelWithCallback(nn.Module):
    def __init__(self, module, device_ids=None, output_device=None, dim=0):
        super(DataParallelWithCallback, self).__init__()
        if device_ids is not None and len(device_ids) > 1:
            raise RuntimeError('monkey-net_b200 is one-process-per-GPU: launch with `torchrun --nproc-per-node N` '
                               'and pass a single device id per process instead of device_ids=%r' % (device_ids,))
        self.module = module
        self.device_ids = list(device_ids) if device_ids is not None else None
        self.dim = dim
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            # The reference's DataParallel re-broadcasts the weights of device 0 to every replica on every forward
            # (replicate.py:64-67) and neither run.py nor train.py seeds, so replicas may start from different random
            # initialisations: rank 0's parameters and buffers become everybody's once, as DDP does at construction.
            broadcast_module_state(module)
            for p in module.parameters():
                if p.requires_grad and id(p) not in _hooked:
                    _hooked.add(id(p))
                    p.register_post_accumulate_grad_hook(_avg_hook)

    def _device(self):
        for p in self.module.parameters():
            return p.device
        for b in self.module.buffers():
            return b.device
        return torch.device('cuda', torch.cuda.current_device())

    def forward(self, *inputs, **kwargs):
        dev = self._device()
        return self.module(*_to_device(inputs, dev), **_to_device(kwargs, dev))


def patch_replication_callback(data_parallel):
    """Kept for API compatibility (replicate.py:70-94); there is nothing to patch with one replica per process."""
    return data_parallel
