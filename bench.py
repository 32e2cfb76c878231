#!/usr/bin/env python
"""bench.py - frames/sec of the Monkey-Net training step on N x B200, one process per GPU.

Default workload (every N, weak scaling) = BASELINE.json configs[3]'s per-GPU share: config/taichi.yaml nets run at
256x256, 8 synthetic frame pairs per GPU, the full train.py:110-136 iteration body (G fwd+bwd + Adam(G, KP) + D
fwd+bwd + Adam(D)), data-parallel with NCCL sync-BN - at N = 8 that IS configs[3] (global batch 64).  At N = 1 the same
line also carries, as extra objects, the other single-GPU configurations of BASELINE.json:
  `configs1_shapes64`  configs[1]  shapes.yaml training step, 32 pairs @64x64
  `transfer_256`       configs[2]  moving-gif.yaml nets @256x256, transfer_one 16 sources x 2 driving frames, with the
                                   max-abs error of the timed path against the CPU oracle on the same weights / inputs
  `voxfull_256`        configs[4]  per-GPU share: vox-full.yaml nets @256x256, 16 pairs, + grid_sample HBM GB/s

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU implementation on the host cores

Prints ONE JSON line (rank 0).  `value` = whole-job frames/s with inputs resident in HBM; `e2e` = same metric through
the public API with pinned HOST inputs, H2D copy and D2H loss read inside the timed region; `roofline` = the
convolution kernels' algorithmic FLOP/s vs the measured tensor peak; `cpu_baseline` = the reference's modules (from
the byte-compiled oracle/_ref build; the oracle port when that is absent) timed on this box's host cores on a bounded
sample of the same workload.  Training convolutions run in the 'auto' arithmetic: split-operand (TF32 + BF16 cross terms) tensor-core kernels
(fp32-accurate); `tf32_fast` reports the same step with 1xTF32 for readers who train at TF32 like stock PyTorch.
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
DEFAULT = {'config': 'taichi', 'res': 256, 'batch': 8}
CONV_ENTRIES = ('mk_conv2d', 'mk_conv2d_wgrad', 'mk_conv2d_tc', 'mk_conv2d_wgrad_tc', 'mk_conv2d_tc_x3',
                'mk_conv2d_wgrad_tc_x3', 'mk_conv2d_tc_halo', 'mk_conv2d_tc_halo_x3', 'mk_conv2d_wgrad_halo',
                'mk_conv2d_wgrad_halo_x3', 'mk_conv2d_tc_halo_ups', 'mk_conv2d_tc_halo_ups_x3', 'mk_conv2d_wgrad_halo_ups',
                'mk_conv2d_wgrad_halo_ups_x3')


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default=DEFAULT['config'])
    ap.add_argument('--res', type=int, default=DEFAULT['res'])
    ap.add_argument('--batch', type=int, default=DEFAULT['batch'], help='frame pairs PER GPU (weak scaling)')
    ap.add_argument('--cpu-batch', type=int, default=None, help='batch of the CPU baseline sample')
    ap.add_argument('--graph', default='auto', choices=['auto', 'on', 'off'],
                    help='run the iteration as one CUDA graph launch (monkey_net_b200.train_step.GraphedTrainer)')
    ap.add_argument('--workload', default='train', choices=['train', 'transfer'],
                    help="train = the headline; transfer = configs[2] as the bench line itself (replicas for N > 1)")
    ap.add_argument('--adam', default='flat', choices=['flat', 'torch'],
                    help='flat = monkey_net_b200.optim.FlatAdam (one fused launch per optimiser step); torch = torch.optim.Adam')
    ap.add_argument('--no-extras', action='store_true', help='N = 1: skip configs1_shapes64 / transfer_256 / voxfull_256')
    ap.add_argument('--no-transfer', action='store_true')
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


class Harness:
    """device, process group, L2 flush and the timing rule of the contract (CUDA events per step, barrier +
    synchronize on both sides, max over ranks)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise RuntimeError('bench.py (impl=ours) needs a CUDA device; there is no CPU fallback')
        torch.cuda.set_device(self.local)
        self.device = torch.device('cuda', self.local)
        if self.world > 1:
            import datetime
            # rendezvous of 8 cold-starting ranks can take minutes on a fresh box (first `import torch`)
            dist.init_process_group('nccl', device_id=self.device, timeout=datetime.timedelta(seconds=480))
        from monkey_net_b200 import lib
        lib.load()
        self.lib = lib
        self.flush = torch.zeros(256 << 20, dtype=torch.uint8, device=self.device)  # > 126 MB L2

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            torch.cuda.synchronize()

    def timed(self, step_fn, steps):
        """total ms of `steps` calls, L2 flushed (write + read pass, outside the event pairs) before each"""
        st = torch.cuda.current_stream().cuda_stream
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        for s, e in evs:
            self.lib.call('mk_fill_zero', self.flush.data_ptr(), self.flush.numel(), st)
            self.lib.call('mk_l2_evict', self.flush.data_ptr(), self.flush.numel(), st)
            s.record()
            step_fn()
            e.record()
        self.barrier()
        ms = sum(s.elapsed_time(e) for s, e in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def measure_train(h, cfg_name, res, batch, steps, warmup, adam='flat', graph=True, conv_pass=True, e2e=True,
                  conv_mode=None, sample_clocks=False):
    """One training workload: resident + e2e frames/s and the convolution timing pass.  Every rank runs everything
    that contains collectives."""
    from monkey_net_b200 import lib, train_step
    from monkey_net_b200 import ops as mkops
    prev_mode = mkops.CONV_MODE
    if conv_mode:
        mkops.set_conv_mode(conv_mode)
    try:
        cfg = load_config(cfg_name)
        tp = cfg['train_params']
        gen, disc, kp = build_nets(cfg, h.device)
        for m in (gen, disc, kp):
            m.train()
        trainer = train_step.GraphedTrainer(kp, gen, disc, tp, use_graph=graph, fused_adam=(adam == 'flat'))
        torch.manual_seed(100 + h.rank)
        host = {'source': torch.rand(batch, 3, 1, res, res).pin_memory(),
                'video': torch.rand(batch, 3, 1, res, res).pin_memory()}
        resident = {k: v.to(h.device) for k, v in host.items()}
        last = {}

        def step_resident():
            trainer.step(resident)

        def step_e2e():
            last['loss'] = trainer.step(host).cpu()  # H2D copy of the pinned batch + step + D2H read of the losses

        for _ in range(warmup):
            step_resident()
        sampler = ClockSampler(h.local) if sample_clocks and h.rank == 0 else None
        if sampler:
            sampler.start()
        n0 = lib.launches()
        ms = h.timed(step_resident, steps)
        launches = trainer.kernels_per_step * steps if trainer.graph is not None else lib.launches() - n0 - 2 * steps
        out = {'ms_per_step': ms / steps, 'value': batch * h.world * steps / (ms / 1e3), 'gpu_launches': launches,
               'cuda_graph': trainer.graph is not None, 'conv_mode': mkops.CONV_MODE,
               'kernels_per_step': trainer.kernels_per_step}
        if sampler:
            out['clocks'] = sampler.stop()
        if e2e:
            for _ in range(2):
                step_e2e()
            ms_e2e = h.timed(step_e2e, steps)
            out['e2e'] = {'value': batch * h.world * steps / (ms_e2e / 1e3), 'unit': 'frames/s',
                          'h2d_bytes_per_step': sum(v.numel() * 4 for v in host.values()),
                          'd2h_bytes_per_step': int(last['loss'].numel() * 4)}
        if conv_pass:
            # device time of every convolution launch: CUDA events around each launch on the launching stream in an
            # eager (ungraphed) pass; the stream is parked behind a spin so the host enqueues the whole iteration ahead
            # of the device and every event pair brackets back-to-back device execution, not a host gap
            spans = []
            orig, orig_soft = lib.call, lib.call_soft

            def traced(name, *a):
                if name in CONV_ENTRIES:
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record(); orig(name, *a); e.record()
                    spans.append((s, e))
                else:
                    orig(name, *a)

            def traced_soft(name, soft, *a):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                rc = orig_soft(name, soft, *a)
                e.record()
                if rc == 0 and name in CONV_ENTRIES:
                    spans.append((s, e))
                return rc
            lib.call, lib.call_soft = traced, traced_soft
            try:
                torch.cuda.synchronize()
                torch.cuda._sleep(int(0.060 * 1.9e9))
                trainer._iteration(resident)
                torch.cuda.synchronize()
            finally:
                lib.call, lib.call_soft = orig, orig_soft
            out['conv_ms_per_step'] = sum(s.elapsed_time(e) for s, e in spans)
            out['conv_launches_per_step'] = len(spans)
        if h.world > 1:
            h.dist.barrier()
        del trainer, gen, disc, kp
        torch.cuda.empty_cache()
        return out
    finally:
        mkops.set_conv_mode(prev_mode)


def roofline(m, flops, pk, steps_note=''):
    achieved = flops['train_step'] / (m['conv_ms_per_step'] / 1e3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp))
        except Exception:
            traffic = None
    x3 = m['conv_mode'] in ('auto', 'tf32x3')
    return {'bound': 'tensor',
            'kernel': 'implicit-GEMM convolutions, fwd + dgrad + wgrad: k_conv_halo / k_conv_tc / k_wgrad_tc (tcgen05 '
                      'kind::tf32%s)' % (' + one kind::f16 BF16 MMA for the cross terms = 2 MMAs per product' if x3 else ', 1xTF32'),
            'achieved': achieved, 'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
            'frac': achieved / pk['bf16_tflops_sustained'],
            'traffic': traffic,
            'frac_of_tf32_peak': achieved / (pk['bf16_tflops_sustained'] / 2),
            'frac_of_reference_precision_ceiling': achieved / (pk['bf16_tflops_sustained'] / 4) if x3 else None,
            'note': 'achieved = ALGORITHMIC conv FLOPs of the step (3*KP2 + 3*G + 12*D, 2*MACs of the reference convs) / '
                    'summed device time of the conv launches (CUDA events around every launch, eager pass, stream parked '
                    'so launches are back to back).  kind::tf32 issues at half the bf16 rate; the reference-precision mode adds '
                    'one BF16 MMA (K = 16, same duration) for the two cross terms, so its ceiling is peak/4 (3xTF32 until '
                    'visit 16: peak/6); `frac` stays against the measured bf16 peak as the contract asks.',
            'peak_source': pk['source'] + ' bf16 sustained (kernel timed inside a long step)',
            'launches_per_step': m['conv_launches_per_step'], 'conv_ms_per_step': m['conv_ms_per_step'],
            'conv_share_of_step': m['conv_ms_per_step'] / m['ms_per_step']}


def run_ours(args):
    h = Harness()
    from monkey_net_b200 import ops as mkops
    use_graph = args.graph in ('on', 'auto')
    dist_check = None
    if h.world > 1:
        from monkey_net_b200 import dist_check as dc
        dist_check = dc.run(h.device)   # N ranks on shards == 1 rank on the full batch, through the kernels + NCCL
    main = measure_train(h, args.config, args.res, args.batch, args.steps, args.warmup, adam=args.adam, graph=use_graph,
                         sample_clocks=True)
    if h.rank != 0:
        _exit_rank()
    pk = peaks()
    cfg = load_config(args.config)
    flops = conv_flops_per_step(cfg, args.res, args.batch)
    mode = main['conv_mode']
    out = {
        'metric': METRIC, 'value': main['value'], 'unit': 'frames/s', 'n_gpus': h.world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': main['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None,
        'dtype': {'auto': 'f32 (split-operand tensor-core convolutions: TF32 main term + BF16 cross terms = fp32-accurate products, fp32 accumulate; fp32 elsewhere)',
                  'tf32x3': 'f32 (split-operand tensor-core convolutions: TF32 main term + BF16 cross terms)', 'tf32': 'tf32 tensor-core convs (fp32 accumulate)',
                  'fp32': 'f32 (FFMA convolutions)'}[mode],
        'data': 'synthetic (torch.rand frames, seeded default-init weights)',
        'config': {'workload': 'config/%s.yaml nets @%dx%d, training step (train.py:110-136: G fwd+bwd+Adam(G,KP), D '
                               'fwd+bwd+Adam(D)), %d synthetic frame pairs per GPU%s'
                               % (args.config, args.res, args.res, args.batch,
                                  ' = BASELINE.json configs[3] (taichi.yaml @256, batch 64 over 8 GPUs) per-GPU share'
                                  if (args.config, args.res, args.batch) == ('taichi', 256, 8) else ''),
                   'global_batch': args.batch * h.world, 'parallelism': 'dp%d' % h.world,
                   'cuda_graph': main['cuda_graph'], 'conv_mode': mode,
                   'optimizer': 'FlatAdam (mk_adam_flat, fused zero_grad)' if args.adam == 'flat' else 'torch.optim.Adam',
                   'l2': 'flushed between timed steps: 256 MiB memset (write) followed by a read pass over the same '
                         'buffer so no dirty lines are written back inside the timed region; outside the event pairs',
                   'conv_gflop_per_sample': flops['train_step_per_sample'] / 1e9},
        'e2e': main['e2e'], 'gpu_launches': main['gpu_launches'], 'clocks': main.get('clocks'),
        'roofline': roofline(main, flops, pk),
    }
    if dist_check is not None:
        out['dist_check'] = dist_check
        from monkey_net_b200 import dist as mkdist
        out['config']['bn_statistics'] = mkdist.stats_backend() + ' (one sum per BN layer per direction; NCCL flat ' \
            'gradient all-reduce per optimiser step, G / KP overlapped with the discriminator step)'
    if h.world == 1:
        if not args.no_cpu_baseline:
            out['cpu_baseline'] = cpu_baseline(args, cfg)
        if not args.no_extras:
            # the same step with every convolution 1xTF32 (what stock PyTorch/cuDNN trains with on this GPU)
            fast = measure_train(h, args.config, args.res, args.batch, max(3, args.steps // 2), 2, conv_pass=True,
                                 e2e=False, conv_mode='tf32', graph=use_graph)
            out['tf32_fast'] = {'value': fast['value'], 'unit': 'frames/s', 'ms_per_step': fast['ms_per_step'],
                                'conv_tflops': flops['train_step'] / (fast['conv_ms_per_step'] / 1e3) / 1e12,
                                'note': 'MONKEY_B200_CONV=tf32: 1xTF32 convolutions; gradient cosine vs the fp32 oracle '
                                        '0.94-0.98 (tests/test_gpu_3_tc.py), not the reference-precision headline'}
            s_cfg = load_config('shapes')
            s = measure_train(h, 'shapes', 64, 32, args.steps, args.warmup, graph=use_graph)
            sf = conv_flops_per_step(s_cfg, 64, 32)
            out['configs1_shapes64'] = {
                'workload': 'BASELINE.json configs[1]: config/shapes.yaml training step, 32 synthetic 64x64 frame pairs',
                'value': s['value'], 'unit': 'frames/s', 'ms_per_step': s['ms_per_step'], 'e2e': s['e2e'],
                'kernels_per_step': s['kernels_per_step'], 'conv_mode': s['conv_mode'],
                'conv_tflops': sf['train_step'] / (s['conv_ms_per_step'] / 1e3) / 1e12,
                'conv_share_of_step': s['conv_ms_per_step'] / s['ms_per_step']}
            if not args.no_cpu_baseline:
                a2 = argparse.Namespace(**vars(args))
                a2.config, a2.res, a2.batch, a2.cpu_batch = 'shapes', 64, 32, None
                out['configs1_shapes64']['cpu_baseline'] = cpu_baseline(a2, s_cfg)
            v = measure_train(h, 'vox-full', 256, 16, max(3, args.steps // 2), 2, graph=use_graph, e2e=False)
            vf = conv_flops_per_step(load_config('vox-full'), 256, 16)
            out['voxfull_256'] = {
                'workload': 'BASELINE.json configs[4] per-GPU share: config/vox-full.yaml nets @256x256 (trilinear '
                            'grid resize, 7-block generator), training step, 16 synthetic frame pairs',
                'value': v['value'], 'unit': 'frames/s', 'ms_per_step': v['ms_per_step'],
                'kernels_per_step': v['kernels_per_step'],
                'conv_tflops': vf['train_step'] / (v['conv_ms_per_step'] / 1e3) / 1e12,
                'conv_share_of_step': v['conv_ms_per_step'] / v['ms_per_step']}
        if not args.no_kernel_bench:
            out['kernels'] = kernel_bench(h.device, pk)
            if 'voxfull_256' in out:
                out['voxfull_256']['grid_sample'] = out['kernels']
        if not args.no_transfer and not args.no_extras:
            out['transfer_256'] = transfer_bench(h.device, 'moving-gif', 256, 16, 2, args.steps, args.warmup,
                                                 cpu=not args.no_cpu_baseline)
    print(json.dumps(out))
    if h.world > 1:
        _exit_rank()


def _exit_rank():
    """Multi-rank exit without NCCL teardown ordering: rank 0 goes on to CPU-side work (FLOP count) after the other
    ranks are done, and ncclCommDestroy on one side waiting for a peer that has already left hung the launcher
    (observed at N=2).  All collectives are complete and synchronised at this point, so the processes flush and leave;
    the driver reads rank 0's JSON line."""
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def transfer_bench(device, config, res, batch, d, steps, warmup, cpu=True):
    """BASELINE.json configs[2] / north-star 256x256 transfer: `config` nets at res x res, eval mode, transfer_one
    (transfer.py:65-79) on `batch` sources x `d` driving frames = batch*d generated frames per call.  `value`: inputs
    resident in HBM; `e2e`: pinned host inputs -> H2D -> graph -> D2H of the predicted frames (what transfer.py:116
    does with `.data.cpu().numpy()`); `parity`: the TIMED path's frames / keypoints against the CPU oracle loaded with
    the same weights on the first 2 sources; `cpu_baseline`: the reference's transfer_one on a bounded sample."""
    from monkey_net_b200 import lib, transfer_step
    from monkey_net_b200 import ops as mkops
    cfg = load_config(config)
    gen, disc, kp = build_nets(cfg, device)
    del disc
    torch.manual_seed(7)
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
        return runner.run(resident['source'], resident['driving'])

    def call_e2e():
        o = runner.run(host['source'], host['driving'])
        out_host.copy_(o['video_prediction'], non_blocking=True)

    for _ in range(max(2, warmup)):
        call_resident()
    ms = timed(call_resident, steps)
    call_e2e()
    ms_e2e = timed(call_e2e, steps)
    frames = batch * d
    r = {'workload': 'BASELINE.json configs[2]: config/%s.yaml nets @%dx%d, eval, transfer_one: %d sources x %d driving '
                     'frames per call, CUDA graph, all driving frames batched into one KP + one generator pass'
                     % (config, res, res, batch, d),
         'metric': 'generated frames/sec (transfer)', 'value': frames / (ms / 1e3), 'unit': 'frames/s',
         'ms_per_call': ms, 'conv_mode': mkops.CONV_MODE, 'kernels_per_call': runner.kernels_per_call,
         'e2e': {'value': frames / (ms_e2e / 1e3), 'unit': 'frames/s',
                 'h2d_bytes_per_step': sum(v.numel() * 4 for v in host.values()),
                 'd2h_bytes_per_step': out_host.numel() * 4}}
    # ---- parity of THIS path (graph replay, fused inference kernels) against the CPU oracle, same weights, same inputs
    try:
        from oracle import monkey_oracle as mo
        og, od, ok = mo.build_from_config(cfg)
        og.load_state_dict({k: v.cpu() for k, v in gen.state_dict().items()})
        ok.load_state_dict({k: v.cpu() for k, v in kp.state_dict().items()})
        for m in (og, ok):
            m.eval()
        got = call_resident()
        with torch.no_grad():
            want = mo.transfer_one(og, ok, x['source'][:2], x['driving'][:2], tparams['normalization_params'])
        g_mean = got['kp_driving']['mean'][:2].cpu()
        px = lambda m_: torch.round(res * (m_ + 1) / 2)
        r['parity'] = {
            'frames_max_abs': float((got['video_prediction'][:2].cpu() - want['video_prediction']).abs().max()),
            'kp_mean_max_abs': float((g_mean - want['kp_driving']['mean']).abs().max()),
            'kp_pixel_indices_identical': bool(torch.equal(px(g_mean), px(want['kp_driving']['mean']))),
            'note': 'timed path vs oracle/monkey_oracle.py (CPU fp32) with the same state_dict on sources 0-1, all %d '
                    'driving frames; north-star bar 1e-3 on the frames, identical keypoint pixel indices' % d}
    except Exception as ex:  # the checker must never take the measurement down
        r['parity'] = {'error': repr(ex)}
    if cpu:
        r['cpu_baseline'] = cpu_transfer_baseline(cfg, config, res, 2, d)
        r['e2e_speedup_vs_cpu'] = r['e2e']['value'] / r['cpu_baseline']['value']
    del runner, gen, kp
    torch.cuda.empty_cache()
    return r


def _reference_or_port(cfg):
    """(generator, discriminator, kp_detector, kind, module namespace): the reference's OWN modules from the
    byte-compiled oracle/_ref build (or the source tree in the build container) when present - cpu_baseline.kind
    'reference' - else the oracle port."""
    from oracle import monkey_oracle as mo
    try:
        from oracle import ref_shim
        if ref_shim.available():
            rg, rd, rk = ref_shim.build_from_config(cfg)
            return rg, rd, rk, 'reference', ref_shim
    except Exception:
        pass
    og, od, ok = mo.build_from_config(cfg)
    return og, od, ok, 'port', None


def cpu_transfer_baseline(cfg, config, res, batch, d):
    from oracle import monkey_oracle as mo
    torch.manual_seed(0)
    og, od, ok, kind, shim = _reference_or_port(cfg)
    for m in (og, ok):
        m.eval()
    if kind == 'reference':
        tmod = shim.load_driver('transfer', modules='reference')
        run = lambda s, v, norm: tmod.transfer_one(og, ok, s, v, {'normalization_params': norm})
    else:
        run = lambda s, v, norm: mo.transfer_one(og, ok, s, v, norm)
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else os.cpu_count()
    x = {'source': torch.rand(batch, 3, 1, res, res), 'driving': torch.rand(batch, 3, d, res, res)}
    norm = cfg['transfer_params']['normalization_params']
    best = None
    probe = {}
    for c in sorted({c for c in (16, 32, 64, ncpu) if c <= ncpu}):
        torch.set_num_threads(c)
        with torch.no_grad():
            run(x['source'], x['driving'], norm)  # warm-up (oneDNN primitive creation)
            t0 = time.perf_counter()
            run(x['source'], x['driving'], norm)
            t = time.perf_counter() - t0
        probe[c] = round(batch * d / t, 2)
        if best is None or t < best[1]:
            best = (c, t)
    torch.set_num_threads(best[0])
    ts = []
    with torch.no_grad():
        for _ in range(3):
            t0 = time.perf_counter()
            run(x['source'], x['driving'], norm)
            ts.append(time.perf_counter() - t0)
    ts.sort()
    return {'value': batch * d / ts[1], 'unit': 'frames/s', 'cores': best[0], 'kind': kind,
            'sample': '%s transfer_one (per-frame loop of transfer.py:65-79), %s.yaml @%dx%d, %d sources x %d driving '
                      'frames, median of 3 after warm-up; %d of %d host threads (fastest of %s)'
                      % ("the reference's own transfer.py + modules (oracle/_ref)" if kind == 'reference' else 'oracle port',
                         config, res, res, batch, d, best[0], ncpu, probe)}


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
def cpu_step_time(cfg, res, batch, steps, warmup):
    """median seconds of one training iteration on the host CPU: the reference's own train.py FullModels + modules
    (oracle/_ref) when available, else the oracle port; returns (seconds, kind)."""
    from oracle import monkey_oracle as mo
    mp, tp = cfg['model_params'], cfg['train_params']
    torch.manual_seed(0)
    og, od, ok, kind, shim = _reference_or_port(cfg)
    with torch.no_grad():
        g = torch.Generator().manual_seed(1)
        w = og.dense_motion_module.hourglass.decoder.conv.weight
        w.copy_(torch.randn(w.shape, generator=g) * 0.05)
    for m in (og, od, ok):
        m.train()
    torch.manual_seed(100)
    x = {'source': torch.rand(batch, 3, 1, res, res), 'video': torch.rand(batch, 3, 1, res, res)}
    if kind == 'reference':
        tr = shim.load_driver('train', modules='reference')
        g_full = tr.GeneratorFullModel(ok, og, od, tp)
        d_full = tr.DiscriminatorFullModel(ok, og, od, tp)
        mk = lambda m: torch.optim.Adam(m.parameters(), lr=tp['lr'], betas=(0.5, 0.999))
        o_g, o_d, o_k = mk(og), mk(od), mk(ok)

        def iteration():   # train.py:110-136, the loop body verbatim (logging excluded)
            out = g_full(x)
            loss_values = [val.mean() for val in out[:-2]]
            generated, kp_joined = out[-2], out[-1]
            sum(loss_values).backward(retain_graph=not tp['detach_kp_discriminator'])
            o_g.step(); o_g.zero_grad(); o_d.zero_grad()
            if tp['detach_kp_discriminator']:
                o_k.step(); o_k.zero_grad()
            loss_values = [val.mean() for val in d_full(x, kp_joined, generated)]
            sum(loss_values).backward()
            o_d.step(); o_d.zero_grad()
            if not tp['detach_kp_discriminator']:
                o_k.step(); o_k.zero_grad()
    else:
        opts = mo.make_optimizers(og, od, ok, tp['lr'])
        iteration = lambda: mo.train_iteration(ok, og, od, opts, tp, x)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        iteration()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    times.sort()
    return times[len(times) // 2], kind


def pick_threads(cfg, res):
    """The CPU arm gets the thread count that is FASTEST on this host (more threads than the op sizes can feed makes
    oneDNN slower, e.g. 128 threads on 64x64 frames): probe a few counts on a small batch, keep the best."""
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else os.cpu_count()
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    pb = 4 if res <= 64 else 1
    best, probe = None, {}
    for c in cands:
        torch.set_num_threads(c)
        t, _ = cpu_step_time(cfg, res, pb, steps=1, warmup=1)
        probe[c] = round(pb / t, 2)
        if best is None or t < best[1]:
            best = (c, t)
        if t > 4 * best[1]:
            break
    torch.set_num_threads(best[0])
    return best[0], probe, ncpu, pb


def cpu_baseline(args, cfg, steps=5, warmup=2):
    threads, probe, ncpu, pb = pick_threads(cfg, args.res)
    b = args.cpu_batch or (args.batch if args.res <= 64 else 2)
    if args.res > 64:
        steps, warmup = min(steps, 3), 1
    t, kind = cpu_step_time(cfg, args.res, b, steps=steps, warmup=warmup)
    what = "the reference's own train.py FullModels + modules (byte-compiled oracle/_ref), unmodified" \
        if kind == 'reference' else 'oracle/monkey_oracle.py (plain-PyTorch CPU restatement of the reference step)'
    return {'value': b / t, 'unit': 'frames/s', 'cores': threads, 'kind': kind,
            'sample': '%s, %s.yaml, batch %d @%dx%d, median of %d steps after %d warm-up; %d of %d host threads (fastest '
                      'of the probed counts, frames/s at batch %d: %s)'
                      % (what, args.config, b, args.res, args.res, steps, warmup, threads, ncpu, pb, probe)}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path on the host cores - its own modules and train.py
    FullModels from the byte-compiled oracle/_ref build (cpu_baseline.kind 'reference'), the pinned oracle port when
    that build is absent - all useful host threads, bounded sample per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = load_config(args.config)
    steps, warm = max(1, min(args.steps, 5 if args.res <= 64 else 3)), max(1, min(args.warmup, 2 if args.res <= 64 else 1))
    cb = cpu_baseline(args, cfg, steps=steps, warmup=warm)
    val = cb['value']
    b = args.cpu_batch or (args.batch if args.res <= 64 else 2)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warm, 'ms_per_step': b / val * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'config/%s.yaml training step on the host CPU, batch %d @%dx%d (bounded sample of the '
                               'GPU arm\'s workload; CPU cost per frame is flat in the batch)'
                               % (args.config, b, args.res, args.res)},
        'cpu_baseline': cb,
        'e2e': {'value': val, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def run_transfer(args):
    """--workload transfer: the 256x256 transfer configuration as the bench line itself (replicas only for N > 1:
    eval-mode inference has no exchange step)."""
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    cfgname = args.config if args.config != DEFAULT['config'] else 'moving-gif'
    res = args.res
    batch = args.batch if args.batch != DEFAULT['batch'] else 16
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
               'dtype': 'geometry networks fp32-accurate (TF32 + BF16 cross terms), appearance path 1xTF32 tensor-core convs (fp32 accumulate)',
               'data': 'synthetic (torch.rand frames, seeded default-init weights)',
               'config': {'workload': r['workload'], 'parallelism': 'replicas x%d' % world,
                          'l2': 'flushed between timed calls'},
               'e2e': dict(r['e2e'], value=frames / (float(t[1]) / 1e3)),
               'gpu_launches': r['kernels_per_call'] * args.steps, 'clocks': clocks, 'parity': r.get('parity')}
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
