"""Launch the hot kernels on representative shapes (for `ncu --set full` captures and quick CUDA-event timings).

    python tools/prof_kernels.py [grid|conv|all]
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
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(s.elapsed_time(e))
    return sum(ts) / len(ts)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    dev = torch.device('cuda', 0)
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if what in ('grid', 'all'):
        B, d = 16, 1
        for C, h in ((4, 256), (64, 128), (128, 64), (256, 32)):
            inp = torch.rand(B, h, h, C, device=dev)
            ys, xs = torch.meshgrid(torch.linspace(-1, 1, 256, device=dev), torch.linspace(-1, 1, 256, device=dev),
                                    indexing='ij')
            deform = (torch.stack([xs, ys], -1)[None] + 0.05 * torch.randn(B, 256, 256, 2, device=dev)).contiguous()
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


if __name__ == '__main__':
    main()
