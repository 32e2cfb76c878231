"""Per-tap probe of the halo-window tensor-core conv (csrc/conv_tc_halo.cu): weights non-zero for ONE filter tap at a
time, both descriptor base-offset conventions, against the exact fp32 kernel.  Prints one line per (convention, tap)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from monkey_net_b200 import lib  # noqa: E402


def run(cin, cout, k, pad, H, W, N, taps, baseoff):
    os.environ['MONKEY_B200_HALO_BASEOFF'] = '1' if baseoff else '0'
    dev = torch.device('cuda')
    st = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(1)
    x = torch.randn(N, H, W, cin, device=dev)
    w = torch.zeros(cout, cin, 1, k, k, device=dev)
    full = torch.randn(cout, cin, 1, k, k, device=dev) / (cin * k * k) ** 0.5
    for (r, s) in taps:
        w[:, :, 0, r, s] = full[:, :, 0, r, s]
    Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    wp, wt = torch.empty(k * k * cin * cout, device=dev), torch.empty(k * k * cin * cout, device=dev)
    lib.call('mk_pack_weight', w.data_ptr(), cout, cin, k, k, 1, None, cin, cout, 0, wp.data_ptr(), None, None, st)
    lib.call('mk_pack_weight', w.data_ptr(), cout, cin, k, k, 1, None, cin, cout, 2, wt.data_ptr(), None, None, st)
    y0 = torch.empty(N, Ho, Wo, cout, device=dev)
    y1 = torch.full((N, Ho, Wo, cout), float('nan'), device=dev)
    lib.call('mk_conv2d', x.data_ptr(), N, H, W, cin, cin, 0, wp.data_ptr(), k, k, pad, None, None, None, 0, 0, 0.0,
             y0.data_ptr(), cout, cout, 0, st)
    lib.call('mk_conv2d_tc_halo', x.data_ptr(), N, H, W, cin, cin, wt.data_ptr(), k, k, pad, None, None, None, 0, 0, 0.0,
             y1.data_ptr(), cout, cout, st)
    torch.cuda.synchronize()
    nan = int(torch.isnan(y1).sum())
    err = float((y0 - torch.nan_to_num(y1)).abs().max()) / (float(y0.abs().max()) + 1e-12)
    return err, nan


if __name__ == '__main__':
    for baseoff in (0, 1):
        for k, pad in ((3, 1), (4, 0)):
            for r in range(k):
                for s in range(k):
                    err, nan = run(32, 32, k, pad, 32, 32, 2, [(r, s)], baseoff)
                    print('baseoff=%d k=%d tap=(%d,%d) shift=%d rel_err=%.3e nan=%d' % (baseoff, k, r, s, 16 * r + s, err, nan),
                          flush=True)
            alltaps = [(r, s) for r in range(k) for s in range(k)]
            for (cin, cout, H, W, N) in ((32, 32, 32, 32, 2), (48, 48, 64, 64, 2), (160, 32, 16, 16, 2), (24, 144, 17, 21, 2)):
                err, nan = run(cin, cout, k, pad, H, W, N, alltaps, baseoff)
                print('baseoff=%d k=%d ALL taps cin=%d cout=%d %dx%d rel_err=%.3e nan=%d' % (baseoff, k, cin, cout, H, W, err, nan),
                      flush=True)
