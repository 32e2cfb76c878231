"""Micro-benchmark of the forward / dgrad tensor-core convolution kernels on the layers that dominate taichi@256 and
shapes@64: per-tap kernel (mk_conv2d_tc) vs halo-window persistent kernel (mk_conv2d_tc_halo), 1xTF32 and 3xTF32,
CUDA events, L2 flushed between launches.  Prints us, algorithmic TFLOP/s and the HBM floor of each layer."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import torch  # noqa: E402
from monkey_net_b200 import lib  # noqa: E402
from prof_kernels import timeit  # noqa: E402

CASES = [  # N, H, cin, cout, k, pad, resid
    (8, 256, 48, 48, 3, 1, 1), (8, 256, 48, 48, 3, 1, 0), (8, 256, 128, 32, 3, 1, 0), (8, 256, 32, 128, 3, 1, 0),
    (8, 256, 140, 32, 3, 1, 0), (16, 256, 4, 64, 3, 1, 0), (16, 256, 12, 36, 3, 1, 0), (16, 256, 36, 12, 3, 1, 0),
    (8, 256, 16, 64, 4, 0, 0), (8, 253, 64, 16, 4, 3, 0), (8, 256, 44, 44, 1, 0, 0), (8, 256, 76, 16, 3, 1, 0),
    (8, 128, 64, 128, 3, 1, 0), (8, 126, 64, 128, 4, 0, 0), (16, 128, 64, 256, 3, 1, 0), (16, 64, 128, 512, 3, 1, 0),
    (32, 64, 24, 24, 3, 1, 1), (32, 64, 16, 32, 3, 1, 0), (32, 64, 4, 16, 3, 1, 0), (32, 61, 16, 32, 4, 0, 0),
]


def main():
    dev = torch.device('cuda', 0)
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    sel = sys.argv[1:] and [int(a) for a in sys.argv[1:]]
    print('%-34s %10s %10s %10s %10s %9s' % ('layer', 'tc us', 'halo us', 'tc_x3 us', 'halo_x3 us', 'HBM us'))
    for i, (N, h, cin, cout, k, pad, resid) in enumerate(CASES):
        if sel and i not in sel:
            continue
        x = torch.randn(N, h, h, cin, device=dev)
        w = torch.randn(cout, cin, 1, k, k, device=dev) * 0.05
        ho = h + 2 * pad - k + 1
        y = torch.empty(N, ho, ho, cout, device=dev)
        r = torch.randn(N, ho, ho, cout, device=dev) if resid else None
        fl = 2.0 * N * ho * ho * cin * cout * k * k
        hbm_us = 4.0 * (x.numel() + y.numel() * (2 if resid else 1)) / 6.568e12 * 1e6
        res = []
        for x3 in (0, 1):
            wt = torch.empty(k * k * cout * (cin + (((cin + 7) & ~7) if x3 else 0)), device=dev)
            lib.call('mk_pack_weight', w.data_ptr(), cout, cin, k, k, 1, None, cin, cout, 2 | (8 if x3 else 0),
                     wt.data_ptr(), None, None, st)
            rp, ldr = (r.data_ptr(), cout) if resid else (None, 0)
            tc = 'mk_conv2d_tc_x3' if x3 else 'mk_conv2d_tc'
            ms = timeit(lambda: lib.call(tc, x.data_ptr(), N, h, h, cin, cin, 0, wt.data_ptr(), k, k, pad, None, None, rp,
                                         ldr, 0, 0.0, y.data_ptr(), cout, cout, st), flush)
            y0 = y.clone()
            res.append(ms * 1e3)
            hl = 'mk_conv2d_tc_halo_x3' if x3 else 'mk_conv2d_tc_halo'
            try:
                ms = timeit(lambda: lib.call(hl, x.data_ptr(), N, h, h, cin, cin, wt.data_ptr(), k, k, pad, None, None, rp,
                                             ldr, 0, 0.0, y.data_ptr(), cout, cout, st), flush)
                err = float((y - y0).abs().max()) / float(y0.abs().max())
                res.append(ms * 1e3 if err < 2e-3 else -err)
            except RuntimeError:
                res.append(float('nan'))
        name = 'N%d %dx%d %d->%d k%d p%d%s' % (N, h, h, cin, cout, k, pad, ' +res' if resid else '')
        print('%-34s %10.1f %10.1f %10.1f %10.1f %9.1f   TF/s: %s' % (
            name, res[0], res[1], res[2], res[3], hbm_us, ' '.join('%.0f' % (fl / (u * 1e-6) / 1e12) for u in res)), flush=True)


if __name__ == '__main__':
    main()
