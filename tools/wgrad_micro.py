"""Micro-benchmark of the weight-gradient kernels: mk_conv2d_wgrad_tc vs mk_conv2d_wgrad_halo, 1xTF32 and 3xTF32."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import torch  # noqa: E402
from monkey_net_b200 import lib  # noqa: E402
from prof_kernels import timeit  # noqa: E402

CASES = [  # N, H, cin, cout, k, pad
    (8, 256, 48, 48, 3, 1), (8, 256, 128, 32, 3, 1), (8, 256, 32, 128, 3, 1), (8, 256, 140, 32, 3, 1),
    (16, 256, 4, 64, 3, 1), (16, 256, 36, 12, 3, 1), (16, 256, 12, 36, 3, 1), (8, 256, 16, 64, 4, 0),
    (8, 256, 76, 16, 3, 1), (8, 128, 64, 128, 3, 1), (8, 126, 64, 128, 4, 0), (16, 128, 256, 64, 3, 1),
    (16, 64, 512, 128, 3, 1), (32, 64, 24, 24, 3, 1), (32, 64, 16, 32, 3, 1), (32, 61, 16, 32, 4, 0),
]


def main():
    dev = torch.device('cuda', 0)
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    sel = sys.argv[1:] and [int(a) for a in sys.argv[1:]]
    print('%-30s %10s %10s %10s %10s %9s' % ('layer', 'tc us', 'halo us', 'tc_x3 us', 'halo_x3 us', 'HBM us'))
    for i, (N, h, cin, cout, k, pad) in enumerate(CASES):
        if sel and i not in sel:
            continue
        x = torch.randn(N, h, h, cin, device=dev)
        ho = h + 2 * pad - k + 1
        dy = torch.randn(N, ho, ho, cout, device=dev)
        d0 = torch.empty(k * k * cin * cout, device=dev)
        d1 = torch.empty(k * k * cin * cout, device=dev)
        fl = 2.0 * N * ho * ho * cin * cout * k * k
        hbm_us = 4.0 * (x.numel() + dy.numel()) / 6.568e12 * 1e6
        res = []
        for x3 in ('', '_x3'):
            ms = timeit(lambda: lib.call('mk_conv2d_wgrad_tc' + x3, x.data_ptr(), N, h, h, cin, cin, dy.data_ptr(), cout, cout,
                                         k, k, pad, d0.data_ptr(), st), flush)
            res.append(ms * 1e3)
            try:
                ms = timeit(lambda: lib.call('mk_conv2d_wgrad_halo' + x3, x.data_ptr(), N, h, h, cin, cin, dy.data_ptr(), cout,
                                             cout, k, k, pad, d1.data_ptr(), st), flush)
                err = float((d1 - d0).abs().max()) / float(d0.abs().max())
                res.append(ms * 1e3 if err < 3e-3 else -err)
            except RuntimeError:
                res.append(float('nan'))
        name = 'N%d %dx%d %d->%d k%d p%d' % (N, h, h, cin, cout, k, pad)
        print('%-30s %10.1f %10.1f %10.1f %10.1f %9.1f   TF/s: %s' % (
            name, res[0], res[1], res[2], res[3], hbm_us, ' '.join('%.0f' % (fl / (u * 1e-6) / 1e12) for u in res)), flush=True)


if __name__ == '__main__':
    main()
