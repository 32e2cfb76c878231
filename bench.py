#!/usr/bin/env python
"""bench.py - frames/sec of the Monkey-Net training step (BASELINE.json configs[1]: config/shapes.yaml nets, batch 32
synthetic 64x64 frame pairs per GPU, full train.py:110-136 iteration body: G fwd+bwd + Adam(G,KP) + D fwd+bwd +
Adam(D)) on N x B200, one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU restatement of the reference (oracle) on the host cores

Prints ONE JSON line (rank 0).  `value` = whole-job frames/s with inputs resident in HBM; `e2e` = same metric through
the public API (DataParallelWithCallback(GeneratorFullModel)(x)) with pinned HOST inputs, H2D copy and D2H loss read
inside the timed region; `roofline` = conv kernels' algorithmic FLOP/s vs the measured tensor peak;
`cpu_baseline` = the oracle port timed on this box's host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import torch  # noqa: E402
import yaml  # noqa: E402

METRIC = 'frames/sec (training step: G fwd+bwd+Adam, D fwd+bwd+Adam)'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='shapes')
    ap.add_argument('--res', type=int, default=64)
    ap.add_argument('--batch', type=int, default=32, help='frame pairs PER GPU (weak scaling)')
    ap.add_argument('--cpu-batch', type=int, default=None, help='batch of the CPU baseline sample')
    ap.add_argument('--graph', default='auto', choices=['auto', 'on', 'off'],
                    help='run the iteration as one CUDA graph launch (monkey_net_b200.train_step.GraphedTrainer)')
    ap.add_argument('--workload', default='train', choices=['train', 'transfer'],
                    help="train = configs[1] (the headline); transfer = configs[2]: moving-gif nets, 256x256, "
                         "transfer_one on 16 sources x 2 driving frames (use --config moving-gif --res 256 --batch 16)")
    ap.add_argument('--no-transfer', action='store_true', help='skip the 256x256 transfer side measurement')
    ap.add_argument('--adam', default='flat', choices=['flat', 'torch'],
                    help='flat = monkey_net_b200.optim.FlatAdam (one fused launch per optimiser step); torch = torch.optim.Adam')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-kernel-bench', action='store_true')
    return ap.parse_args()


def load_config(name):
    with open(os.path.join(ROOT, 'config', name + '.yaml')) as f:
        return yaml.safe_load(f)


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'],
                'bf16_tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


# ------------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == 'Active' for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------- our arm
def build_nets(cfg, device):
    from modules.generator import MotionTransferGenerator
    from modules.discriminator import Discriminator
    from modules.keypoint_detector import KPDetector
    mp = cfg['model_params']
    torch.manual_seed(0)
    gen = MotionTransferGenerator(**mp['generator_params'], **mp['common_params'])
    disc = Discriminator(**mp['discriminator_params'], **mp['common_params'])
    kp = KPDetector(**mp['kp_detector_params'], **mp['common_params'])
    with torch.no_grad():  # non-trivial flow field
        g = torch.Generator().manual_seed(1)
        w = gen.dense_motion_module.hourglass.decoder.conv.weight
        w.copy_(torch.randn(w.shape, generator=g) * 0.05)
    return gen.to(device), disc.to(device), kp.to(device)


def conv_flops_per_step(cfg, res, batch):
    """Algorithmic conv FLOPs of one training iteration for `batch` samples: 3*KP2 + 3*G + 12*D, from
    hooking every conv of the oracle (2*MACs, forward)."""
    from oracle import monkey_oracle as mo
    og, od, ok = mo.build_from_config(cfg)
    for m in (og, od, ok):
        for p in m.parameters():
            torch.nn.init.normal_(p, std=0.01)
        m.eval()
    x = torch.rand(1, 3, 1, res, res)
    with torch.no_grad():
        kpj = ok(torch.cat([x, x], 2))
    kd = {k: v[:, 1:] for k, v in kpj.items()}
    ks = {k: v[:, :1] for k, v in kpj.items()}
    f_kp = mo.conv_flops(ok, torch.cat([x, x], 2))
    f_g = mo.conv_flops(og, x, kd, ks)
    f_d = mo.conv_flops(od, x, kd, ks)
    return {'kp2': f_kp, 'g': f_g, 'd': f_d, 'train_step_per_sample': 3 * f_kp + 3 * f_g + 12 * f_d,
            'train_step': batch * (3 * f_kp + 3 * f_g + 12 * f_d)}


def run_ours(args):
    import torch.distributed as dist
    from monkey_net_b200 import lib, train_step
    from sync_batchnorm import DataParallelWithCallback

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py (impl=ours) needs a CUDA device; there is no CPU fallback')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        import datetime
        # rendezvous of 8 cold-starting ranks can take minutes on a fresh box (first `import torch`); collectives
        # themselves are milliseconds, so a diverged rank still fails the run well inside the driver's patience
        dist.init_process_group('nccl', device_id=device, timeout=datetime.timedelta(seconds=480))
    lib.load()
    from monkey_net_b200 import ops as mkops
    conv_mode = mkops.CONV_MODE
    cfg = load_config(args.config)
    tp = cfg['train_params']
    gen, disc, kp = build_nets(cfg, device)
    for m in (gen, disc, kp):
        m.train()
    use_graph = args.graph == 'on' or (args.graph == 'auto')
    trainer = train_step.GraphedTrainer(kp, gen, disc, tp, use_graph=use_graph, fused_adam=(args.adam == 'flat'))
    B = args.batch
    torch.manual_seed(100 + rank)
    host = {'source': torch.rand(B, 3, 1, args.res, args.res).pin_memory(),
            'video': torch.rand(B, 3, 1, args.res, args.res).pin_memory()}
    resident = {k: v.to(device) for k, v in host.items()}
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(step_fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for s, e in evs:
            lib.call('mk_fill_zero', flush.data_ptr(), flush.numel(), torch.cuda.current_stream().cuda_stream)
            lib.call('mk_l2_evict', flush.data_ptr(), flush.numel(), torch.cuda.current_stream().cuda_stream)
            s.record()
            step_fn()
            e.record()
        barrier()
        ms = sum(s.elapsed_time(e) for s, e in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_resident():
        trainer.step(resident)

    last = {}

    def step_e2e():
        last['loss'] = trainer.step(host).cpu()  # H2D copy of the pinned batch + step + D2H read of the losses

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.launches()
    ms = timed(step_resident, args.steps)
    if trainer.graph is not None:   # kernels recorded into the CUDA graph at capture time, replayed every step
        launches = trainer.kernels_per_step * args.steps
    else:
        launches = lib.launches() - n0 - 2 * args.steps  # minus the L2-flush memset + read pass
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # ---- roofline of the dominant kernel family (implicit-GEMM convolutions): device time by CUDA events around
    # every conv launch on the launching stream, algorithmic FLOPs from the oracle's conv hooks
    conv_ms = None
    n_conv = 0
    if True:  # EVERY rank runs this pass: the iteration contains collectives (BN statistics, gradient all-reduce)
        names = ('mk_conv2d', 'mk_conv2d_wgrad', 'mk_conv2d_tc', 'mk_conv2d_wgrad_tc')
        spans = []
        orig = lib.call

        def traced(name, *a):
            if name in names:
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); orig(name, *a); e.record()
                spans.append((s, e))
            else:
                orig(name, *a)
        lib.call = traced
        from monkey_net_b200 import ops as _ops
        _ops.lib.call = traced
        torch.cuda.synchronize()
        # The eager loop is CPU-bound at this size (host launch cost > kernel time): an event pair would then
        # measure the host gap, not the kernel.  Park the stream behind a ~40 ms spin so the host enqueues the whole
        # iteration ahead of the device; every pair then brackets back-to-back device execution.
        torch.cuda._sleep(int(0.040 * 1.9e9))
        trainer._iteration(resident)  # eager (ungraphed) pass so every conv launch can be bracketed by events
        torch.cuda.synchronize()
        lib.call = orig
        _ops.lib.call = orig
        conv_ms = sum(s.elapsed_time(e) for s, e in spans)
        n_conv = len(spans)
    if world > 1:
        dist.barrier()

    if rank != 0:
        _exit_rank()
    pk = peaks()
    flops = conv_flops_per_step(cfg, args.res, B)
    frames = B * world * args.steps
    out = {
        'metric': METRIC, 'value': frames / (ms / 1e3), 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None,
        'dtype': 'tf32 tensor-core convs (fp32 accumulate), fp32 elsewhere' if conv_mode == 'tf32' else 'f32',
        'data': 'synthetic (torch.rand frames, seeded default-init weights)',
        'config': {'workload': 'config/%s.yaml training step, %d synthetic %dx%d frame pairs per GPU, 1 driving '
                               'frame, fwd+bwd+Adam for G, KP and D' % (args.config, B, args.res, args.res),
                   'global_batch': B * world, 'parallelism': 'dp%d' % world,
                   'cuda_graph': trainer.graph is not None, 'conv_mode': conv_mode,
                   'optimizer': 'FlatAdam (mk_adam_flat, fused zero_grad)' if args.adam == 'flat' else 'torch.optim.Adam',
                   'l2': 'flushed between timed steps: 256 MiB memset (write) followed by a read pass over the same buffer so no dirty lines are left to be written back inside the timed region; both outside the per-step event pairs',
                   'conv_gflop_per_sample': flops['train_step_per_sample'] / 1e9},
        'e2e': {'value': frames / (ms_e2e / 1e3), 'unit': 'frames/s',
                'h2d_bytes_per_step': sum(v.numel() * 4 for v in host.values()),
                'd2h_bytes_per_step': int(last['loss'].numel() * 4)},
        'gpu_launches': launches,
        'clocks': clocks,
    }
    achieved = flops['train_step'] / (conv_ms / 1e3) / 1e12
    out['roofline'] = {'bound': 'tensor', 'kernel': 'k_conv_tc/k_wgrad_tc (tcgen05 tf32) + k_conv_ffma/k_conv_wgrad (fp32) - implicit-GEMM conv fwd, dgrad, wgrad',
                       'achieved': achieved, 'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                       'frac': achieved / pk['bf16_tflops_sustained'], 'traffic': None,
                       'frac_of_tf32_peak': achieved / (pk['bf16_tflops_sustained'] / 2),
                       'tf32_note': 'kind::tf32 issues at half the bf16 rate: the TF32 ceiling is peak/2; at 64x64 with '
                                    '16-128 channels the convs are L2->SMEM-bandwidth bound, not tensor bound '
                                    '(profiles/r1_ncu_conv_tc.md)',
                       'peak_source': pk['source'] + ' bf16 sustained (kernel timed inside a long step)',
                       'launches_per_step': n_conv, 'conv_ms_per_step': conv_ms,
                       'conv_share_of_step': conv_ms / (ms / args.steps),
                       'note': 'conv kernels timed in an eager pass (stream parked behind a spin so launches are back to back) with CUDA events around every launch; algorithmic '
                               'FLOPs = 3*KP2 + 3*G + 12*D conv FLOPs (2*MACs of the reference convs)'}
    if world == 1 and not args.no_kernel_bench:
        out['kernels'] = kernel_bench(device, pk)
    if world == 1 and not args.no_cpu_baseline:  # 'on rank 0 at N=1 only'
        out['cpu_baseline'] = cpu_baseline(args, cfg)
    if world == 1 and conv_mode == 'tf32' and not args.no_kernel_bench:
        # the same step with the EXACT fp32 (FFMA) convolutions, for readers who want the no-TF32 number
        mkops.set_conv_mode('fp32')
        try:
            g2, d2, k2 = build_nets(cfg, device)
            for m in (g2, d2, k2):
                m.train()
            t2 = train_step.GraphedTrainer(k2, g2, d2, tp, use_graph=use_graph)
            for _ in range(3):
                t2.step(resident)
            n2 = max(3, args.steps // 2)
            ms2 = timed(lambda: t2.step(resident), n2)
            out['fp32_exact'] = {'value': B * n2 / (ms2 / 1e3), 'unit': 'frames/s', 'ms_per_step': ms2 / n2,
                                 'note': 'MONKEY_B200_CONV=fp32: every convolution on the exact FFMA kernels'}
            del t2, g2, d2, k2
        finally:
            mkops.set_conv_mode(conv_mode)
    if world == 1 and not args.no_transfer:
        del trainer, gen, disc, kp
        torch.cuda.empty_cache()
        out['transfer_256'] = transfer_bench(device, 'moving-gif', 256, 16, 2, args.steps, args.warmup,
                                             cpu=not args.no_cpu_baseline)
    print(json.dumps(out))
    if world > 1:
        _exit_rank()


def _exit_rank():
    """Multi-rank exit without NCCL teardown ordering: rank 0 goes on to CPU-side work (FLOP count, CPU baseline) for
    up to a minute after the other ranks are done, and ncclCommDestroy on one side waiting for a peer that has already
    left hung the launcher (observed at N=2).  All collectives are complete and synchronised at this point, so the
    processes flush and leave; the driver reads rank 0's JSON line."""
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def transfer_bench(device, config, res, batch, d, steps, warmup, cpu=True):
    """BASELINE.json configs[2] / north-star 256x256 transfer: `config` nets at res x res, eval mode, transfer_one
    (transfer.py:65-79) on `batch` sources x `d` driving frames = batch*d generated frames per call.  `value`: inputs
    resident in HBM; `e2e`: pinned host inputs -> H2D -> graph -> D2H of the predicted frames (what transfer.py:116
    does with `.data.cpu().numpy()`); `cpu_baseline`: the oracle's transfer_one on a bounded sample."""
    from monkey_net_b200 import lib, transfer_step
    from monkey_net_b200 import ops as mkops
    cfg = load_config(config)
    gen, disc, kp = build_nets(cfg, device)
    del disc
    x = {'source': torch.rand(batch, 3, 1, res, res), 'driving': torch.rand(batch, 3, d, res, res)}
    with torch.no_grad():  # populate the BN running statistics the way a trained checkpoint would carry them
        for m in (gen, kp):
            m.train()
        kj = kp(torch.cat([x['source'][:2], x['driving'][:2, :, :1]], 2).to(device))
        gen(x['source'][:2].to(device), {k: v[:, 1:] for k, v in kj.items()}, {k: v[:, :1] for k, v in kj.items()})
    for m in (gen, kp):
        m.eval()
    tparams = cfg['transfer_params']
    runner = transfer_step.GraphedTransfer(gen, kp, tparams, use_graph=True)
    host = {k: v.pin_memory() for k, v in x.items()}
    resident = {k: v.to(device) for k, v in x.items()}
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=device)
    out_host = torch.empty(batch, 3, d, res, res).pin_memory()

    def timed(fn, n):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        torch.cuda.synchronize()
        for s_, e_ in evs:
            lib.call('mk_fill_zero', flush.data_ptr(), flush.numel(), torch.cuda.current_stream().cuda_stream)
            lib.call('mk_l2_evict', flush.data_ptr(), flush.numel(), torch.cuda.current_stream().cuda_stream)
            s_.record(); fn(); e_.record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / n

    def call_resident():
        runner.run(resident['source'], resident['driving'])

    def call_e2e():
        o = runner.run(host['source'], host['driving'])
        out_host.copy_(o['video_prediction'], non_blocking=True)

    for _ in range(max(2, warmup)):
        call_resident()
    ms = timed(call_resident, steps)
    call_e2e()
    ms_e2e = timed(call_e2e, steps)
    frames = batch * d
    r = {'workload': 'config/%s.yaml nets @%dx%d, eval, transfer_one: %d sources x %d driving frames per call, CUDA '
                     'graph, all driving frames batched into one KP + one generator pass' % (config, res, res, batch, d),
         'metric': 'generated frames/sec (transfer)', 'value': frames / (ms / 1e3), 'unit': 'frames/s',
         'ms_per_call': ms, 'conv_mode': mkops.CONV_MODE, 'kernels_per_call': runner.kernels_per_call,
         'e2e': {'value': frames / (ms_e2e / 1e3), 'unit': 'frames/s',
                 'h2d_bytes_per_step': sum(v.numel() * 4 for v in host.values()),
                 'd2h_bytes_per_step': out_host.numel() * 4}}
    if cpu:
        r['cpu_baseline'] = cpu_transfer_baseline(cfg, config, res, 2, d)
        r['e2e_speedup_vs_cpu'] = r['e2e']['value'] / r['cpu_baseline']['value']
    del runner, gen, kp
    torch.cuda.empty_cache()
    return r


def cpu_transfer_baseline(cfg, config, res, batch, d):
    from oracle import monkey_oracle as mo
    og, od, ok = mo.build_from_config(cfg)
    torch.manual_seed(0)
    for m in (og, ok):
        m.eval()
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else os.cpu_count()
    x = {'source': torch.rand(batch, 3, 1, res, res), 'driving': torch.rand(batch, 3, d, res, res)}
    norm = cfg['transfer_params']['normalization_params']
    best = None
    probe = {}
    for c in sorted({c for c in (16, 32, 64, ncpu) if c <= ncpu}):
        torch.set_num_threads(c)
        with torch.no_grad():
            mo.transfer_one(og, ok, x['source'], x['driving'], norm)  # warm-up (oneDNN primitive creation)
            t0 = time.perf_counter()
            mo.transfer_one(og, ok, x['source'], x['driving'], norm)
            t = time.perf_counter() - t0
        probe[c] = round(batch * d / t, 2)
        if best is None or t < best[1]:
            best = (c, t)
    torch.set_num_threads(best[0])
    ts = []
    with torch.no_grad():
        for _ in range(3):
            t0 = time.perf_counter()
            mo.transfer_one(og, ok, x['source'], x['driving'], norm)
            ts.append(time.perf_counter() - t0)
    ts.sort()
    return {'value': batch * d / ts[1], 'unit': 'frames/s', 'cores': best[0], 'kind': 'port',
            'sample': 'oracle transfer_one (per-frame loop of transfer.py:65-79), %s.yaml @%dx%d, %d sources x %d '
                      'driving frames, median of 3 after warm-up; %d of %d host threads (fastest of %s)'
                      % (config, res, res, batch, d, best[0], ncpu, probe)}


def kernel_bench(device, pk):
    """grid_sample HBM roofline on the large vox-full pyramid levels: algorithmic bytes
    4*(B*C*h*w + 2*B*d*h*w + B*d*C*h*w) / CUDA-event time, L2 flushed between launches."""
    from monkey_net_b200 import lib
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import prof_kernels
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=device)
    res = {'deformation': 'identity + smooth displacement (8x8 noise, bicubic, amplitude 0.3) at 256x256, nearest '
                          'resize to the level'}
    B, d = 16, 1
    for (C, h) in ((64, 128), (128, 64), (4, 256)):
        inp = torch.rand(B, h, h, C, device=device)
        deform = prof_kernels.smooth_deformation(B * d, 256, device)  # smooth, network-like flow field
        out = torch.empty(B * d, h, h, C, device=device)
        times = []
        for it in range(8):
            lib.call('mk_fill_zero', flush.data_ptr(), flush.numel(), st)
            lib.call('mk_l2_evict', flush.data_ptr(), flush.numel(), st)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            lib.call('mk_grid_sample_fwd', inp.data_ptr(), B, h, h, C, C, deform.data_ptr(), d, 256, 256, 0,
                     out.data_ptr(), C, st)
            e.record()
            torch.cuda.synchronize()
            if it >= 3:
                times.append(s.elapsed_time(e))
        ms = sum(times) / len(times)
        logical_c = 3 if C == 4 else C
        byts = 4 * (B * logical_c * h * h + 2 * B * d * h * h + B * d * logical_c * h * h)
        gbs = byts / (ms / 1e3) / 1e9
        res['grid_sample_fwd_%dx%d2' % (logical_c, h)] = {'ms': ms, 'achieved_gbs': gbs, 'frac_hbm': gbs / pk['hbm_gbs']}
    return res


# ------------------------------------------------------------------------------------------------- CPU arms
def oracle_step_time(cfg, res, batch, steps, warmup):
    from oracle import monkey_oracle as mo
    from modules.generator import MotionTransferGenerator
    from modules.discriminator import Discriminator
    from modules.keypoint_detector import KPDetector
    mp, tp = cfg['model_params'], cfg['train_params']
    torch.manual_seed(0)
    pg = MotionTransferGenerator(**mp['generator_params'], **mp['common_params'])
    pd = Discriminator(**mp['discriminator_params'], **mp['common_params'])
    pk = KPDetector(**mp['kp_detector_params'], **mp['common_params'])
    og, od, ok = mo.build_from_config(cfg)
    og.load_state_dict(pg.state_dict()); od.load_state_dict(pd.state_dict()); ok.load_state_dict(pk.state_dict())
    with torch.no_grad():
        g = torch.Generator().manual_seed(1)
        w = og.dense_motion_module.hourglass.decoder.conv.weight
        w.copy_(torch.randn(w.shape, generator=g) * 0.05)
    for m in (og, od, ok):
        m.train()
    opts = mo.make_optimizers(og, od, ok, tp['lr'])
    torch.manual_seed(100)
    x = {'source': torch.rand(batch, 3, 1, res, res), 'video': torch.rand(batch, 3, 1, res, res)}
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        mo.train_iteration(ok, og, od, opts, tp, x)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    times.sort()
    return times[len(times) // 2]


def pick_threads(cfg, res):
    """The CPU arm gets the thread count that is FASTEST on this host (more threads than the op sizes can feed makes
    oneDNN slower, e.g. 128 threads on 64x64 frames): probe a few counts on a small batch, keep the best."""
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else os.cpu_count()
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    best, probe = None, {}
    for c in cands:
        torch.set_num_threads(c)
        t = oracle_step_time(cfg, res, 4, steps=1, warmup=1)
        probe[c] = round(4 / t, 2)
        if best is None or t < best[1]:
            best = (c, t)
        if t > 4 * best[1]:
            break
    torch.set_num_threads(best[0])
    return best[0], probe, ncpu


def cpu_baseline(args, cfg):
    threads, probe, ncpu = pick_threads(cfg, args.res)
    b = args.cpu_batch or (args.batch if args.res <= 64 else 2)
    t = oracle_step_time(cfg, args.res, b, steps=5, warmup=2)
    return {'value': b / t, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
            'sample': 'oracle/monkey_oracle.py (plain-PyTorch CPU restatement of the reference step), %s.yaml, '
                      'batch %d @%dx%d, median of 5 steps after 2 warm-up; %d of %d host threads (fastest of the '
                      'probed counts, frames/s at batch 4: %s)' % (args.config, b, args.res, args.res, threads, ncpu,
                                                                   probe)}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path = the pinned oracle port (the reference itself
    cannot travel to the GPU box and has no installable package), all host threads, bounded sample per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = load_config(args.config)
    threads, probe, ncpu = pick_threads(cfg, args.res)
    b = args.cpu_batch or (args.batch if args.res <= 64 else 2)
    steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    t = oracle_step_time(cfg, args.res, b, steps=steps, warmup=warm)
    val = b / t
    cb = {'value': val, 'unit': 'frames/s', 'cores': torch.get_num_threads(), 'kind': 'port',
          'sample': 'oracle port, %s.yaml, batch %d @%dx%d, median of %d steps; %d of %d host threads (fastest of '
                    'the probed counts, frames/s at batch 4: %s)' % (args.config, b, args.res, args.res, steps, threads,
                                                                     ncpu, probe)}
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warm, 'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'config/%s.yaml training step, CPU, batch %d @%dx%d' % (args.config, b, args.res,
                                                                                        args.res)},
        'cpu_baseline': cb,
        'e2e': {'value': val, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def run_transfer(args):
    """--workload transfer: the 256x256 transfer configuration as the bench line itself (replicas only for N > 1:
    eval-mode inference has no exchange step)."""
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    cfgname = args.config if args.config != 'shapes' else 'moving-gif'
    res = args.res if args.res != 64 else 256
    batch = args.batch if args.batch != 32 else 16
    if args.impl == 'reference':
        if rank == 0:
            cb = cpu_transfer_baseline(load_config(cfgname), cfgname, res, 2, 2)
            print(json.dumps({'impl': 'reference', 'metric': 'generated frames/sec (transfer)', 'value': cb['value'],
                              'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': 3, 'warmup': 1,
                              'ms_per_step': 4 / cb['value'] * 1e3, 'higher_is_better': True, 'scaling': 'weak',
                              'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                              'config': {'workload': cb['sample']}, 'cpu_baseline': cb,
                              'e2e': {'value': cb['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0,
                                      'd2h_bytes_per_step': 0}}))
        return
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=device, timeout=datetime.timedelta(seconds=480))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    r = transfer_bench(device, cfgname, res, batch, 2, args.steps, args.warmup,
                       cpu=(rank == 0 and not args.no_cpu_baseline))
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([r['ms_per_call'], 1e3 * batch * 2 / r['e2e']['value']], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        frames = batch * 2 * world
        out = {'metric': r['metric'], 'value': frames / (float(t[0]) / 1e3), 'unit': 'frames/s', 'n_gpus': world,
               'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': float(t[0]), 'higher_is_better': True,
               'scaling': 'weak', 'vs_baseline': None,
               'dtype': 'tf32 tensor-core convs (fp32 accumulate), fp32 elsewhere' if r['conv_mode'] == 'tf32' else 'f32',
               'data': 'synthetic (torch.rand frames, seeded default-init weights)',
               'config': {'workload': r['workload'], 'parallelism': 'replicas x%d' % world,
                          'l2': 'flushed between timed calls'},
               'e2e': dict(r['e2e'], value=frames / (float(t[1]) / 1e3)),
               'gpu_launches': r['kernels_per_call'] * args.steps, 'clocks': clocks}
        if 'cpu_baseline' in r:
            out['cpu_baseline'] = r['cpu_baseline']
        print(json.dumps(out))
    if world > 1:
        _exit_rank()


if __name__ == '__main__':
    a = parse()
    if a.workload == 'transfer':
        run_transfer(a)
    elif a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
