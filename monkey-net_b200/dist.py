"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch; gloo in CPU tests).

Replaces the reference's single-process DataParallel + thread-rendezvous sync-BN (sync_batchnorm/batchnorm.py:90-125,
sync_batchnorm/comm.py) with
  * ONE all-reduce of the packed `[sum x | sum x^2]` (2*Cp floats) per BN layer in forward and ONE of
    `[sum dz | sum dz*xhat]` in backward, issued on the compute stream;
  * one bucketed gradient all-reduce (average) per optimiser step (`GradSync`).
The batch is sharded on dim 0 with equal shares per rank (train.py:99 uses drop_last=True), so the global pixel
count is world_size x the local one.  The data path itself needs no other collective.
"""
import torch
import torch.distributed as dist

_sync_bn = True


def set_sync_bn(enabled):
    global _sync_bn
    _sync_bn = bool(enabled)


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def stats_world():
    return world() if _sync_bn else 1


class _PeerStats:
    """One-shot all-reduce of the packed BN statistics over NVLink peer memory (csrc/p2p.cu) on a symmetric buffer of
    torch.distributed._symmetric_memory: ~3-5 us per call inside the captured graph instead of an NCCL launch."""

    def __init__(self, device):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        from . import lib
        nbytes = lib.load().mk_stats_allreduce_bytes()
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, dist.group.WORLD)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        assert len(ptrs) == self.world and ptrs[self.rank] == self.buf.data_ptr()
        self.peers = (ctypes.c_ulonglong * self.world)(*ptrs)
        self.seq = torch.zeros(1, dtype=torch.int64, device=device)
        self.lib = lib
        torch.cuda.synchronize(device)
        dist.barrier()            # every rank's buffer is zeroed before anybody pushes into it

    def all_reduce(self, t):
        assert t.is_contiguous() and t.dtype in (torch.float64, torch.float32)
        self.lib.call('mk_stats_allreduce', t.data_ptr(), t.numel(), 1 if t.dtype == torch.float64 else 0, self.peers,
                      self.rank, self.world, self.seq.data_ptr(), torch.cuda.current_stream().cuda_stream)


_peer = {'tried': False, 'obj': None, 'why': None}


def _peer_stats(t):
    """the peer-memory communicator, or None (CPU tensors / gloo tests, MONKEY_B200_BN_P2P=0, symmetric memory
    unavailable): the caller then uses torch.distributed.all_reduce (NCCL)."""
    import os
    if not t.is_cuda or os.environ.get('MONKEY_B200_BN_P2P', '1') == '0':
        return None
    if not _peer['tried']:
        _peer['tried'] = True
        try:
            _peer['obj'] = _PeerStats(t.device)
        except Exception as e:  # noqa: BLE001 - any failure (no P2P access, old torch) falls back to NCCL, loudly once
            _peer['why'] = repr(e)
            if rank() == 0:
                import sys
                sys.stderr.write('monkey-net_b200: peer-memory BN statistics unavailable (%s); using NCCL all-reduce\n' % e)
    return _peer['obj']


def stats_backend():
    return 'nvlink-peer-memory' if _peer['obj'] is not None else ('nccl' if world() > 1 else 'single')


def all_reduce_stats(t):
    """Sum-all-reduce a packed statistics tensor in place; returns the factor the local count must be scaled by."""
    w = stats_world()
    if w > 1:
        peer = _peer_stats(t)
        if peer is not None and t.numel() <= 4096:
            peer.all_reduce(t)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return w


def combine_stats(sum_x, sum_x2, count, eps=1e-5):
    """Host-side restatement of mk_norm_finalize for tests: mean, invstd (biased) and unbiased variance."""
    mean = sum_x / count
    var = (sum_x2 / count - mean * mean).clamp_min(0)
    return mean, (var + eps).rsqrt(), var * count / max(count - 1, 1)


class GradSync:
    """Bucketed gradient averaging across ranks (the DDP-style replacement of DataParallel's reduce-to-GPU-0).

    Parameters are flattened into buckets of ~`bucket_mb`; `sync()` all-reduces every bucket that has gradients and
    scatters the averages back.  Call it once between backward() and optimizer.step()."""

    def __init__(self, params, bucket_mb=64):
        self.params = [p for p in params if p.requires_grad]
        self.buckets, cur, size = [], [], 0
        limit = bucket_mb * (1 << 20) // 4
        for p in self.params:
            cur.append(p)
            size += p.numel()
            if size >= limit:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)

    def sync(self):
        w = world()
        if w == 1:
            return
        for bucket in self.buckets:
            grads = [p.grad for p in bucket if p.grad is not None]
            if not grads:
                continue
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.div_(w)
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n


class FlatGradSync:
    """Gradient averaging with ZERO copies: the `.grad` of every parameter of the group is a view into one flat
    buffer (like DDP's gradient_as_bucket_view), so data-parallel training needs exactly one all-reduce per optimiser
    step - `sync()` between `backward()` and `optimizer.step()` - instead of one small collective per parameter.
    Requires `optimizer.zero_grad(set_to_none=False)` (torch 0.4.1's zero_grad, which the reference was written for,
    zeroes in place too).  Deterministic layout: parameters in `module.parameters()` order."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def intact(self):
        base = self.flat.data_ptr()
        return all(p.grad is not None and base <= p.grad.data_ptr() < base + self.flat.numel() * 4 for p in self.params)

    def sync(self):
        w = world()
        if w == 1:
            return
        assert self.intact(), 'a gradient view was replaced (zero_grad(set_to_none=True)?)'
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.mul_(1.0 / w)


def shard_batch(x, dim=0):
    """This rank's contiguous share of a global batch (DataParallel scatter semantics, train.py:104-110)."""
    w, r = world(), rank()
    if w == 1:
        return x
    if isinstance(x, dict):
        return {k: shard_batch(v, dim) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(shard_batch(v, dim) for v in x)
    if torch.is_tensor(x):
        n = x.shape[dim]
        assert n % w == 0, 'global batch %d must divide by world size %d' % (n, w)
        return x.narrow(dim, r * (n // w), n // w)
    return x
