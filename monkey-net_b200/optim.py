"""Flat-group Adam for the CUDA-graph training step (optimiser / step machinery).

`FlatAdam(module.parameters(), lr, betas)` moves the parameters of one optimiser group into ONE flat fp32 buffer
(every `nn.Parameter` keeps its identity, name and shape - its `.data` becomes a view), gives them gradient views into
a second flat buffer and keeps both Adam moments flat as well.  `step()` is then a single `mk_adam_flat` launch that
also zeroes the gradients (the reference calls `optimizer.zero_grad()` straight after every `step()`,
train.py:118-136), the step count lives on the device so the launch can be captured in a CUDA graph, and a
data-parallel run needs exactly one all-reduce over the flat gradient per optimiser step.

Same update as `torch.optim.Adam(params, lr, betas)` (train.py:81-83; eps 1e-8, no weight decay, no amsgrad);
tests/test_gpu_4_graph.py compares the two step by step.  `torch.optim.Adam` itself keeps working on the drop-in
modules (the unchanged train.py uses it); this class is what `GraphedTrainer` runs.
"""
import torch
import torch.distributed as dist

from . import lib


class FlatAdam:
    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, 'empty parameter group'
        dev = self.params[0].device
        if dev.type != 'cuda':
            raise RuntimeError('monkey-net_b200: FlatAdam needs CUDA parameters - the B200 path has no CPU fallback')
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        # the learning rate lives in DEVICE memory: a captured graph replays fixed kernel arguments, so a host float
        # would freeze MultiStepLR (train.py:92-97,146-148) at its capture-time value
        self._lr = float(lr)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self.param_groups = [{'lr': float(lr), 'initial_lr': float(lr), 'betas': self.betas, 'eps': self.eps}]
        # every parameter starts on a 16-byte boundary so the kernel's float4 path never straddles two tensors' tails
        offs, total = [], 0
        for p in self.params:
            assert p.dtype == torch.float32 and p.device == dev
            offs.append(total)
            total += (p.numel() + 3) & ~3
        self.n = total
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self._ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.no_grad():
            for p, off in zip(self.params, offs):
                n = p.numel()
                self.flat_p[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + n].view(p.shape)
                p.grad = self.flat_g[off:off + n].view(p.shape)

    # ------------------------------------------------------------------ torch.optim-like surface
    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, value):
        """host-side schedule update: one tiny H2D fill outside the graph, the captured step reads the new value"""
        self._lr = float(value)
        self.param_groups[0]['lr'] = self._lr
        self.lr_dev.fill_(self._lr)

    def intact(self):
        lo, hi = self.flat_g.data_ptr(), self.flat_g.data_ptr() + 4 * self.n
        plo, phi = self.flat_p.data_ptr(), self.flat_p.data_ptr() + 4 * self.n
        return all(p.grad is not None and lo <= p.grad.data_ptr() < hi and plo <= p.data_ptr() < phi
                   for p in self.params)

    def sync_gradients(self):
        """data parallel: ONE all-reduce (average) of the whole group's gradient."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)
            self.flat_g.mul_(1.0 / dist.get_world_size())

    def step(self, zero_grad=True):
        lib.call('mk_adam_flat', self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(),
                 self.exp_avg_sq.data_ptr(), self.n, self.lr, self.betas[0], self.betas[1], self.eps,
                 self.step_count.data_ptr(), self._ticket.data_ptr(), 1 if zero_grad else 0,
                 self.lr_dev.data_ptr(), torch.cuda.current_stream().cuda_stream)
        from . import ops
        ops.note_parameters_changed()  # in-place update through raw pointers: invalidate cached inference packs

    def zero_grad(self, set_to_none=False):
        if set_to_none:
            raise ValueError('FlatAdam keeps gradient views: zero_grad(set_to_none=True) would orphan them')
        lib.call('mk_fill_zero', self.flat_g.data_ptr(), self.n * 4, torch.cuda.current_stream().cuda_stream)

    def state_dict(self):
        return {'exp_avg': self.exp_avg.clone(), 'exp_avg_sq': self.exp_avg_sq.clone(),
                'step': self.step_count.clone(), 'lr': self.lr, 'betas': self.betas, 'eps': self.eps}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd['exp_avg']); self.exp_avg_sq.copy_(sd['exp_avg_sq']); self.step_count.copy_(sd['step'])
        self.betas, self.eps = tuple(sd['betas']), float(sd['eps'])
        self.lr = float(sd['lr'])


class MultiStepLR:
    """torch.optim.lr_scheduler.MultiStepLR for FlatAdam (torch's class insists on a torch Optimizer): same schedule
    as train.py:92-97 - lr = initial_lr * gamma ** (number of milestones <= epoch), `step()` once per epoch
    (train.py:146-148).  The new rate is written into the optimiser's device scalar, so a captured training graph
    picks it up on its next replay."""

    def __init__(self, optimizer, milestones, gamma=0.1, last_epoch=-1):
        self.optimizer, self.milestones, self.gamma = optimizer, sorted(int(m) for m in milestones), float(gamma)
        self.base_lr = optimizer.param_groups[0]['initial_lr']
        self.last_epoch = last_epoch
        self.step()

    def get_last_lr(self):
        return [self.optimizer.lr]

    def step(self):
        self.last_epoch += 1
        k = sum(1 for m in self.milestones if m <= self.last_epoch)
        self.optimizer.lr = self.base_lr * self.gamma ** k
