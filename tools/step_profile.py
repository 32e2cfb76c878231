"""Device-side anatomy of one training iteration replayed as a CUDA graph (no ncu, no CPU launch gaps).

    python tools/step_profile.py --config shapes --res 64 --batch 32 [--out gpurun_out/step_shapes64.md]

Captures the iteration with GraphedTrainer, replays it under torch.profiler (CUPTI kernel activity records = true
device durations, warm L2) and prints (1) time per kernel family, (2) the convolution launches matched, in launch
order, to the layer shapes recorded at capture time with count / mean us / TFLOP/s.  These are the numbers
DESIGN.md quotes for per-layer efficiency; ncu launch lists under profiles/ remain the cold-cache cross-check.
"""
import argparse
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import torch  # noqa: E402
import yaml  # noqa: E402

KERNEL_OF = {'mk_conv2d_tc': 'k_conv_tc', 'mk_conv2d_wgrad_tc': 'k_wgrad_tc', 'mk_conv2d': 'k_conv_ffma',
             'mk_conv2d_wgrad': 'k_conv_wgrad', 'mk_conv2d_tc_x3': 'k_conv_tc', 'mk_conv2d_wgrad_tc_x3': 'k_wgrad_tc',
             'mk_conv2d_tc_halo': 'k_conv_halo', 'mk_conv2d_tc_halo_x3': 'k_conv_halo',
             'mk_conv2d_wgrad_halo': 'k_wgrad_halo', 'mk_conv2d_wgrad_halo_x3': 'k_wgrad_halo',
             'mk_conv2d_tc_halo_ups': 'k_conv_halo', 'mk_conv2d_tc_halo_ups_x3': 'k_conv_halo',
             'mk_conv2d_wgrad_halo_ups': 'k_wgrad_halo', 'mk_conv2d_wgrad_halo_ups_x3': 'k_wgrad_halo'}


def short(name):
    name = name.replace('(anonymous namespace)::', '').replace('void ', '')
    name = re.sub(r'\(.*', '', name)
    return re.sub(r'<.*', '', name) if (name.startswith('at::') or name.startswith('k_conv_halo') or name.startswith('k_wgrad_halo')) else name


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='shapes')
    ap.add_argument('--res', type=int, default=64)
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--top', type=int, default=30)
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    import bench
    import conv_bench
    from monkey_net_b200 import lib, train_step, ops
    from torch.profiler import profile, ProfilerActivity
    lib.load()
    dev = torch.device('cuda', 0)
    cfg = yaml.safe_load(open(os.path.join(ROOT, 'config', args.config + '.yaml')))
    gen, disc, kp = bench.build_nets(cfg, dev)
    for m in (gen, disc, kp):
        m.train()
    tr = train_step.GraphedTrainer(kp, gen, disc, cfg['train_params'], use_graph=True)
    torch.manual_seed(0)
    x = {'source': torch.rand(args.batch, 3, 1, args.res, args.res, device=dev),
         'video': torch.rand(args.batch, 3, 1, args.res, args.res, device=dev)}
    # record the conv call sequence while the graph is captured (same order as the kernels in the graph)
    calls = []
    orig, orig_soft = lib.call, lib.call_soft

    def traced(name, *a):
        if name in KERNEL_OF:
            calls.append((name, conv_bench.signature(name, a)))
        orig(name, *a)

    def traced_soft(name, soft, *a):
        rc = orig_soft(name, soft, *a)
        if rc == 0 and name in KERNEL_OF:
            sig, fl = conv_bench.signature(name, a)
            reps = 4 if 'halo_ups' in name else 1   # four sub-pixel passes = four kernel launches per call
            calls.extend([(name, (sig, fl / reps))] * reps)
        return rc
    lib.call, lib.call_soft = traced, traced_soft
    tr.step(x)  # warm-up iterations + capture; `calls` keeps growing, the capture is the LAST iteration
    lib.call, lib.call_soft = orig, orig_soft
    per_iter = len(calls) // (tr.warmup + 1)
    calls = calls[-per_iter:]
    for _ in range(3):
        tr.step(x)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        tr.step(x)
    t1.record()
    torch.cuda.synchronize()
    step_ms = t0.elapsed_time(t1) / 5
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tr.step(x)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and 'Memcpy' not in e.name]
    evs.sort(key=lambda e: e.time_range.start)
    fam = collections.defaultdict(lambda: [0, 0.0])
    seq = collections.defaultdict(list)
    for e in evs:
        n = short(e.name)
        dur = e.time_range.end - e.time_range.start  # us
        fam[n][0] += 1
        fam[n][1] += dur
        seq[n.split('<')[0]].append(dur)
    busy = sum(v[1] for v in fam.values())
    lines = ['# %s@%d B=%d, conv mode %s: one graph replay' % (args.config, args.res, args.batch, ops.CONV_MODE), '',
             'step (CUDA events, 5 replays): %.3f ms; kernels in the graph: %d; summed kernel time %.3f ms'
             % (step_ms, len(evs), busy / 1e3), '', '| kernel | launches | us | share of kernel time |',
             '|---|---:|---:|---:|']
    for n, (c, t) in sorted(fam.items(), key=lambda kv: -kv[1][1])[:args.top]:
        lines.append('| `%s` | %d | %.1f | %.1f %% |' % (n[:80], c, t, 100 * t / busy))
    # per-layer conv table
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    idx = collections.defaultdict(int)
    ok = True
    for name, (sig, fl) in calls:
        k = KERNEL_OF[name]
        durs = seq.get(k, [])
        if idx[k] >= len(durs):
            ok = False
            break
        d = durs[idx[k]]
        idx[k] += 1
        a = agg[(name, sig)]
        a[0] += 1; a[1] += d; a[2] = fl
    lines += ['', 'conv launches matched to layers: %s' % ('yes' if ok and all(idx[k] == len(seq.get(k, [])) for k in idx) else 'PARTIAL'),
              '', '| entry | shape | n | us/launch | us/step | TFLOP/s |', '|---|---|---:|---:|---:|---:|']
    tot_fl = tot_us = 0.0
    for (name, sig), (n, us, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        tot_fl += fl * n
        tot_us += us
    for (name, sig), (n, us, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
        lines.append('| %s | %s | %d | %.1f | %.1f | %.1f |' % (name, sig, n, us / n, us, fl / (us / n) / 1e6))
    if tot_us:
        lines += ['', 'all conv launches: %.3f ms/step, %.1f GFLOP issued -> %.1f TFLOP/s'
                  % (tot_us / 1e3, tot_fl / 1e9, tot_fl / tot_us / 1e6)]
    text = '\n'.join(lines)
    print(text)
    if args.out:
        open(args.out, 'w').write(text + '\n')


if __name__ == '__main__':
    main()
