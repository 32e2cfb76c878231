"""Device-side anatomy of one transfer_one call (BASELINE configs[2]: moving-gif nets @256, 16 sources x 2 driving
frames) replayed as a CUDA graph: time per kernel family (CUPTI records via torch.profiler)."""
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402


def short(name):
    name = name.replace('(anonymous namespace)::', '').replace('void ', '')
    name = re.sub(r'\(.*', '', name)
    return re.sub(r'<.*', '', name)


def main():
    import bench
    from monkey_net_b200 import transfer_step, ops
    from torch.profiler import profile, ProfilerActivity
    dev = torch.device('cuda', 0)
    cfg = bench.load_config('moving-gif')
    gen, disc, kp = bench.build_nets(cfg, dev)
    x = {'source': torch.rand(16, 3, 1, 256, 256, device=dev), 'driving': torch.rand(16, 3, 2, 256, 256, device=dev)}
    with torch.no_grad():
        for m in (gen, kp):
            m.train()
        kj = kp(torch.cat([x['source'][:2], x['driving'][:2, :, :1]], 2))
        gen(x['source'][:2], {k: v[:, 1:] for k, v in kj.items()}, {k: v[:, :1] for k, v in kj.items()})
    for m in (gen, kp):
        m.eval()
    runner = transfer_step.GraphedTransfer(gen, kp, cfg['transfer_params'], use_graph=True)
    for _ in range(3):
        runner.run(x['source'], x['driving'])
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        runner.run(x['source'], x['driving'])
    t1.record()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        runner.run(x['source'], x['driving'])
        torch.cuda.synchronize()
    fam = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and 'Memcpy' not in e.name:
            n = short(e.name)
            fam[n][0] += 1
            fam[n][1] += e.time_range.end - e.time_range.start
    busy = sum(v[1] for v in fam.values())
    print('# moving-gif@256 transfer_one 16 x 2, conv mode %s: %.3f ms per call, %d kernels, summed kernel time %.3f ms'
          % (ops.CONV_MODE, t0.elapsed_time(t1) / 5, sum(v[0] for v in fam.values()), busy / 1e3))
    print('| kernel | launches | us | share |\n|---|---:|---:|---:|')
    for n, (c, t) in sorted(fam.items(), key=lambda kv: -kv[1][1])[:25]:
        print('| `%s` | %d | %.1f | %.1f %% |' % (n[:70], c, t, 100 * t / busy))


if __name__ == '__main__':
    main()
