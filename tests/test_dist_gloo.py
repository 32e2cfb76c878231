"""N > 1 host-side logic on CPU with the gloo backend, world_size 2 (one process per rank, as on the GPU box):
batch sharding, the packed BN-statistics all-reduce (forward and backward sums), gradient averaging through the
DataParallelWithCallback facade and GradSync.  The invariant: 2 ranks on half batches == 1 process on the full batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _bn_stats(rank, world):
    from monkey_net_b200 import dist as mkdist
    torch.manual_seed(0)
    x = torch.randn(8, 5, 6, 6) * 2 + 1  # global batch, channels 5
    mine = mkdist.shard_batch(x)
    assert mine.shape[0] == 4
    sums = torch.cat([mine.sum((0, 2, 3)), (mine ** 2).sum((0, 2, 3))])
    factor = mkdist.all_reduce_stats(sums)
    count = mine.numel() / 5 * factor
    mean, invstd, unbiased = mkdist.combine_stats(sums[:5], sums[5:], count)
    return mean, invstd, unbiased


def test_bn_statistics_allreduce_equals_full_batch():
    torch.manual_seed(0)
    x = torch.randn(8, 5, 6, 6) * 2 + 1
    ref_mean = x.mean((0, 2, 3))
    ref_var = x.var((0, 2, 3), unbiased=False)
    for mean, invstd, unbiased in _spawn(_bn_stats):
        assert torch.allclose(mean, ref_mean, atol=1e-5)
        assert torch.allclose(invstd, (ref_var + 1e-5).rsqrt(), atol=1e-5)
        assert torch.allclose(unbiased, x.var((0, 2, 3), unbiased=True), atol=1e-4)


def _bn_backward(rank, world):
    """dx of batch norm on a shard using all-reduced [sum dz | sum dz*xhat] == the full-batch autograd dx."""
    from monkey_net_b200 import dist as mkdist
    torch.manual_seed(1)
    x, g = torch.randn(8, 3, 4, 4), torch.randn(8, 3, 4, 4)
    xs, gs = mkdist.shard_batch(x), mkdist.shard_batch(g)
    sums = torch.cat([xs.sum((0, 2, 3)), (xs ** 2).sum((0, 2, 3))])
    w = mkdist.all_reduce_stats(sums)
    count = xs.numel() / 3 * w
    mean, invstd, _ = mkdist.combine_stats(sums[:3], sums[3:], count)
    xhat = (xs - mean[None, :, None, None]) * invstd[None, :, None, None]
    bsum = torch.cat([gs.sum((0, 2, 3)), (gs * xhat).sum((0, 2, 3))])
    mkdist.all_reduce_stats(bsum)
    dx = invstd[None, :, None, None] * (gs - bsum[:3][None, :, None, None] / count
                                        - xhat * bsum[3:][None, :, None, None] / count)
    return dx


def test_bn_backward_allreduce_equals_full_batch():
    torch.manual_seed(1)
    x, g = torch.randn(8, 3, 4, 4), torch.randn(8, 3, 4, 4)
    xr = x.clone().requires_grad_()
    y = torch.nn.functional.batch_norm(xr, None, None, None, None, True, 0.1, 1e-5)
    (y * g).sum().backward()
    outs = _spawn(_bn_backward)
    assert torch.allclose(torch.cat(outs, 0), xr.grad, atol=1e-5)


def _facade_grads(rank, world):
    from sync_batchnorm import DataParallelWithCallback
    from monkey_net_b200 import dist as mkdist
    torch.manual_seed(2)
    net = torch.nn.Linear(6, 3)
    par = DataParallelWithCallback(net)
    torch.manual_seed(3)
    x = torch.randn(8, 6)
    out = par(mkdist.shard_batch(x))
    out.pow(2).mean().backward()  # local mean; hooks average over ranks -> global mean gradient
    return net.weight.grad.clone(), net.bias.grad.clone()


def test_facade_averages_gradients_like_full_batch():
    torch.manual_seed(2)
    net = torch.nn.Linear(6, 3)
    torch.manual_seed(3)
    x = torch.randn(8, 6)
    net(x).pow(2).mean().backward()
    for gw, gb in _spawn(_facade_grads):
        assert torch.allclose(gw, net.weight.grad, atol=1e-6) and torch.allclose(gb, net.bias.grad, atol=1e-6)


def _gradsync(rank, world):
    from monkey_net_b200 import dist as mkdist
    torch.manual_seed(4)
    net = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.Linear(4, 2))
    sync = mkdist.GradSync(net.parameters(), bucket_mb=1)
    x = mkdist.shard_batch(torch.arange(32.).reshape(8, 4) / 10)
    net(x).sum(1).mean().backward()
    sync.sync()
    return [p.grad.clone() for p in net.parameters()]


def test_gradsync_buckets():
    torch.manual_seed(4)
    net = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.Linear(4, 2))
    net(torch.arange(32.).reshape(8, 4) / 10).sum(1).mean().backward()
    for grads in _spawn(_gradsync):
        for g, p in zip(grads, net.parameters()):
            assert torch.allclose(g, p.grad, atol=1e-6)


def _flat_gradsync(rank, world):
    from monkey_net_b200 import dist as mkdist
    torch.manual_seed(4)
    net = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.Linear(4, 2))
    sync = mkdist.FlatGradSync(net.parameters())
    opt = torch.optim.SGD(net.parameters(), lr=0.0)
    x = mkdist.shard_batch(torch.arange(32.).reshape(8, 4) / 10)
    net(x).sum(1).mean().backward()
    opt.zero_grad(set_to_none=False)          # torch-0.4.1-style zeroing keeps the views
    assert sync.intact() and float(sync.flat.abs().sum()) == 0.0
    net(x).sum(1).mean().backward()           # autograd accumulates INTO the flat buffer
    sync.sync()                               # one all-reduce for all parameters
    return [p.grad.clone() for p in net.parameters()], sync.intact()


def test_flat_gradsync_is_one_allreduce_over_views():
    torch.manual_seed(4)
    net = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.Linear(4, 2))
    net(torch.arange(32.).reshape(8, 4) / 10).sum(1).mean().backward()
    for grads, intact in _spawn(_flat_gradsync):
        assert intact
        for g, p in zip(grads, net.parameters()):
            assert torch.allclose(g, p.grad, atol=1e-6)


def _facade_initial_state(rank, world):
    from sync_batchnorm import DataParallelWithCallback
    torch.manual_seed(100 + rank)                      # run.py / train.py do not seed: every rank draws its own init
    net = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.BatchNorm1d(4))
    with torch.no_grad():
        net[1].running_mean.normal_()
        net[1].num_batches_tracked.fill_(7 + rank)
    DataParallelWithCallback(net)
    return [t.detach().clone() for t in list(net.parameters()) + list(net.buffers())]


def test_facade_broadcasts_rank0_parameters_and_buffers():
    """replicas must start from rank 0's weights (DataParallel re-broadcasts them every forward in the
    reference, sync_batchnorm/replicate.py:64-67); without it averaged gradients are applied to diverging replicas."""
    a, b = _spawn(_facade_initial_state)
    torch.manual_seed(100)
    ref = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.BatchNorm1d(4))
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    for x, p in zip(a, ref.parameters()):
        assert torch.equal(x, p.detach())
    assert int(a[-1]) == int(b[-1]) == 7


def test_facade_rejects_single_process_multi_device():
    from sync_batchnorm import DataParallelWithCallback
    with pytest.raises(RuntimeError, match='one-process-per-GPU'):
        DataParallelWithCallback(torch.nn.Linear(2, 2), device_ids=[0, 1])
