"""Launch the hot kernels on representative shapes (for `ncu --set full` captures and quick CUDA-event timings).

    python tools/prof_kernels.py [grid|conv|convtc|all]
Shapes: grid_sample on the large vox-full pyramid levels (B=16: 3x256^2, 64x128^2, 128x64^2);
conv on shapes.yaml / taichi.yaml layers at batch 32.  Prints one line per kernel with ms and achieved GB/s or TFLOP/s.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from monkey_net_b200 import lib  # noqa: E402


def timeit(fn, flush, reps=5):
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for i in range(reps + 2):
        lib.call('mk_fill_zero', flush.data_ptr(), flush.numel(), st)
        lib.call('mk_l2_evict', flush.data_ptr(), flush.numel(), st)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(s.elapsed_time(e))
    return sum(ts) / len(ts)


def smooth_deformation(n, res, device, amp=0.3, coarse=8, seed=0):
    """identity grid + a smooth displacement field (coarse x coarse gaussian noise, bilinearly upsampled), amplitude
    `amp` in normalised units: the shape of the fields the dense-motion network produces (predicted at low resolution,
    sums of a few keypoint shifts), unlike per-pixel white noise which no layer of the model can emit."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, res), torch.linspace(-1, 1, res), indexing='ij')
    base = torch.stack([xs, ys], -1)[None]
    disp = torch.nn.functional.interpolate(torch.randn(n, 2, coarse, coarse, generator=g), size=(res, res),
                                           mode='bicubic', align_corners=True).permute(0, 2, 3, 1)
    return (base + amp * disp / disp.abs().max()).contiguous().to(device)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    dev = torch.device('cuda', 0)
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    if what in ('grid', 'all'):
        B, d = 16, 1
        for C, h in ((4, 256), (64, 128), (128, 64), (256, 32)):
            inp = torch.rand(B, h, h, C, device=dev)
            deform = smooth_deformation(B, 256, dev)
            out = torch.empty(B, h, h, C, device=dev)
            dinp, ddef = torch.zeros_like(inp), torch.zeros_like(deform)
            lc = 3 if C == 4 else C
            fb = 4 * (B * lc * h * h + 2 * B * h * h + B * lc * h * h)
            ms = timeit(lambda: lib.call('mk_grid_sample_fwd', inp.data_ptr(), B, h, h, C, C, deform.data_ptr(), d, 256,
                                         256, 0, out.data_ptr(), C, st), flush)
            print('grid_sample_fwd C=%d %dx%d: %.4f ms  %.0f GB/s (algorithmic)' % (lc, h, h, ms, fb / ms / 1e6))
            bb = fb + 4 * (B * lc * h * h + B * lc * h * h + 2 * B * h * h)
            ms = timeit(lambda: lib.call('mk_grid_sample_bwd', inp.data_ptr(), B, h, h, C, C, deform.data_ptr(), d, 256,
                                         256, 0, out.data_ptr(), C, dinp.data_ptr(), C, ddef.data_ptr(), st), flush)
            print('grid_sample_bwd C=%d %dx%d: %.4f ms  %.0f GB/s (algorithmic)' % (lc, h, h, ms, bb / ms / 1e6))
    if what in ('conv', 'all'):
        N = 32
        for cin, cout, h, pool in ((16, 32, 32, 0), (64, 128, 8, 0), (32, 64, 64, 0), (256, 512, 16, 0),
                                   (128, 256, 64, 0)):
            x = torch.randn(N, h, h, cin, device=dev)
            w = torch.randn(9 * cin * cout, device=dev) * 0.05
            y = torch.empty(N, h, h, cout, device=dev)
            dw = torch.empty(9 * cin * cout, device=dev)
            fl = 2.0 * N * h * h * cin * cout * 9
            ms = timeit(lambda: lib.call('mk_conv2d', x.data_ptr(), N, h, h, cin, cin, 0, w.data_ptr(), 3, 3, 1, None,
                                         None, None, 0, 0, 0.0, y.data_ptr(), cout, cout, pool, st), flush)
            print('conv3x3 fwd  %4d->%4d @%dx%d N=%d: %.4f ms  %.2f TFLOP/s' % (cin, cout, h, h, N, ms, fl / ms / 1e9))
            ms = timeit(lambda: lib.call('mk_conv2d_wgrad', x.data_ptr(), N, h, h, cin, cin, 0, y.data_ptr(), cout, cout,
                                         3, 3, 1, dw.data_ptr(), st), flush)
            print('conv3x3 wgrad %4d->%4d @%dx%d N=%d: %.4f ms  %.2f TFLOP/s' % (cin, cout, h, h, N, ms, fl / ms / 1e9))


def conv_tc_cases(dev, st, flush):
    """tensor-core conv forward + wgrad on layers of taichi@256 (B=8) and shapes@64 (B=32)."""
    for N, h, cin, cout, k, pad in ((8, 256, 48, 48, 3, 1), (8, 58, 256, 128, 4, 3), (16, 128, 64, 256, 3, 1),
                                    (32, 64, 24, 24, 3, 1)):
        x = torch.randn(N, h, h, cin, device=dev)
        w = torch.randn(cout, cin, 1, k, k, device=dev) * 0.05
        ho = h + 2 * pad - k + 1
        wt = torch.empty(k * k * cin * cout, device=dev)
        lib.call('mk_pack_weight', w.data_ptr(), cout, cin, k, k, 1, None, cin, cout, 2, wt.data_ptr(), None, None, st)
        y = torch.empty(N, ho, ho, cout, device=dev)
        dw = torch.empty(k * k * cin * cout, device=dev)
        fl = 2.0 * N * ho * ho * cin * cout * k * k
        ms = timeit(lambda: lib.call('mk_conv2d_tc', x.data_ptr(), N, h, h, cin, cin, 0, wt.data_ptr(), k, k, pad, None,
                                     None, None, 0, 0, 0.0, y.data_ptr(), cout, cout, st), flush)
        print('conv_tc  fwd   %4d->%4d k%d @%dx%d N=%d: %.4f ms  %.1f TFLOP/s' % (cin, cout, k, h, h, N, ms, fl / ms / 1e9))
        ms = timeit(lambda: lib.call('mk_conv2d_wgrad_tc', x.data_ptr(), N, h, h, cin, cin, y.data_ptr(), cout, cout, k, k,
                                     pad, dw.data_ptr(), st), flush)
        print('conv_tc  wgrad %4d->%4d k%d @%dx%d N=%d: %.4f ms  %.1f TFLOP/s' % (cin, cout, k, h, h, N, ms, fl / ms / 1e9))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'convtc':
        _dev = torch.device('cuda', 0)
        conv_tc_cases(_dev, torch.cuda.current_stream().cuda_stream, torch.zeros(256 << 20, dtype=torch.uint8, device=_dev))
        sys.exit(0)
    main()
