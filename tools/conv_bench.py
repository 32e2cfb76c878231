"""Per-layer timing of the convolution kernels inside one real training iteration.

    python tools/conv_bench.py --config shapes --res 64 --batch 32 [--reps 3]

Runs the eager iteration (train.py:110-136 body) with a CUDA-event pair around every convolution launch on the
launching stream, groups the launches by (entry point, shape signature) and prints count / mean us / total / TFLOP/s,
sorted by total time.  The shapes are the ones the network really issues (forward, dgrad and wgrad of every layer).
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402
import yaml  # noqa: E402

CONV = ('mk_conv2d', 'mk_conv2d_wgrad', 'mk_conv2d_tc', 'mk_conv2d_wgrad_tc', 'mk_conv2d_tc_x3', 'mk_conv2d_wgrad_tc_x3',
        'mk_conv2d_tc_halo', 'mk_conv2d_tc_halo_x3', 'mk_conv2d_wgrad_halo', 'mk_conv2d_wgrad_halo_x3',
        'mk_conv2d_tc_halo_ups', 'mk_conv2d_tc_halo_ups_x3', 'mk_conv2d_wgrad_halo_ups', 'mk_conv2d_wgrad_halo_ups_x3')


def signature(name, a):
    if name in ('mk_conv2d_wgrad_halo_ups', 'mk_conv2d_wgrad_halo_ups_x3'):
        N, H, W, Ci, Co = a[1], a[2], a[3], a[4], a[7]
        return 'N%d %dx%d ci%d co%d k3 p1 ups' % (N, H, W, Ci, Co), 2.0 * N * 4 * H * W * Ci * Co * 9
    if name in ('mk_conv2d_tc_halo_ups', 'mk_conv2d_tc_halo_ups_x3'):
        N, H, W, Ci, Co = a[1], a[2], a[3], a[4], a[12]
        return 'N%d %dx%d ci%d co%d k3 p1 ups' % (N, H, W, Ci, Co), 2.0 * N * 4 * H * W * Ci * Co * 9
    if name in ('mk_conv2d_tc_halo', 'mk_conv2d_tc_halo_x3'):
        N, H, W, Ci, R, pad, Co = a[1], a[2], a[3], a[4], a[7], a[9], a[17]
        Ho, Wo = H + 2 * pad - R + 1, W + 2 * pad - R + 1
        return 'N%d %dx%d ci%d co%d k%d p%d' % (N, H, W, Ci, Co, R, pad), 2.0 * N * Ho * Wo * Ci * Co * R * R
    if name in ('mk_conv2d_tc', 'mk_conv2d', 'mk_conv2d_tc_x3'):
        N, H, W, Ci, ups, R, pad, Co = a[1], a[2], a[3], a[4], a[6], a[8], a[10], a[18]
        Ho, Wo = (H << ups) + 2 * pad - R + 1, (W << ups) + 2 * pad - R + 1
        fl = 2.0 * N * Ho * Wo * Ci * Co * R * R
        return 'N%d %dx%d ci%d co%d k%d p%d%s' % (N, H, W, Ci, Co, R, pad, ' ups' if ups else ''), fl
    if name in ('mk_conv2d_wgrad_tc', 'mk_conv2d_wgrad_tc_x3', 'mk_conv2d_wgrad_halo', 'mk_conv2d_wgrad_halo_x3'):
        N, H, W, Ci, Co, R, pad = a[1], a[2], a[3], a[4], a[7], a[9], a[11]
    else:
        N, H, W, Ci, ups, Co, R, pad = a[1], a[2], a[3], a[4], a[6], a[8], a[10], a[12]
        H, W = H << ups, W << ups
    Ho, Wo = H + 2 * pad - R + 1, W + 2 * pad - R + 1
    return 'N%d %dx%d ci%d co%d k%d p%d' % (N, H, W, Ci, Co, R, pad), 2.0 * N * Ho * Wo * Ci * Co * R * R


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='shapes')
    ap.add_argument('--res', type=int, default=64)
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--top', type=int, default=40)
    args = ap.parse_args()
    import bench
    from monkey_net_b200 import lib, train_step, ops
    lib.load()
    dev = torch.device('cuda', 0)
    cfg = yaml.safe_load(open(os.path.join(ROOT, 'config', args.config + '.yaml')))
    gen, disc, kp = bench.build_nets(cfg, dev)
    for m in (gen, disc, kp):
        m.train()
    tr = train_step.GraphedTrainer(kp, gen, disc, cfg['train_params'], use_graph=False)
    torch.manual_seed(0)
    x = {'source': torch.rand(args.batch, 3, 1, args.res, args.res, device=dev),
         'video': torch.rand(args.batch, 3, 1, args.res, args.res, device=dev)}
    for _ in range(2):
        tr.step(x)
    spans = []
    orig = lib.call

    def traced(name, *a):
        if name in CONV:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); orig(name, *a); e.record()
            spans.append((name, signature(name, a), s, e))
        else:
            orig(name, *a)
    lib.call = traced
    ops.lib.call = traced
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(args.reps):
        tr.step(x)
    t1.record()
    torch.cuda.synchronize()
    lib.call = orig
    ops.lib.call = orig
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for name, (sig, fl), s, e in spans:
        k = (name, sig)
        agg[k][0] += 1
        agg[k][1] += s.elapsed_time(e) * 1e3
        agg[k][2] = fl
    tot = sum(v[1] for v in agg.values()) / args.reps
    print('%s@%d B=%d mode=%s: eager step %.2f ms, conv launches/step %d, conv time/step %.2f ms'
          % (args.config, args.res, args.batch, ops.CONV_MODE, t0.elapsed_time(t1) / args.reps,
             len(spans) // args.reps, tot / 1e3))
    print('%-20s %-34s %5s %9s %9s %8s' % ('entry', 'shape', 'n', 'us/launch', 'us/step', 'TFLOP/s'))
    for (name, sig), (n, us, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
        print('%-20s %-34s %5d %9.1f %9.1f %8.1f' % (name, sig, n // args.reps, us / n, us / args.reps,
                                                     fl / (us / n) / 1e6))


if __name__ == '__main__':
    main()
