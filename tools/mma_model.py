"""Per-layer lower bounds from the operand-row model of an SS-mode UTCHMMA measured on B200 this round,

    t_MMA = 18 clk + 0.35 clk x (M + N)     (M, N = operand rows of 32 bytes read from shared memory per instruction)

next to the CUDA-event timings of tools/conv_micro.py (default: profiles/r2_conv_micro_final.txt).  For every layer:
the MMA phase of the plain schedule (one MMA per tap and K step, M = 128 pixel rows, N = Cout tile), of the
column-taps-on-N schedule where the planner's criterion picks it (one MMA per tap ROW and K step, N = S * Cout), the
HBM floor, and what the measured time is as a multiple of the larger of the two bounds.  1.965 GHz, 148 SMs,
6568 GB/s (MEASURED_PEAKS.json).

    python tools/mma_model.py [profiles/r2_conv_micro_final.txt]
"""
import re
import sys

CLK_GHZ, SMS, HBM_GBS = 1.965, 148, 6568.0


def mma_clk(m, n):
    return 18.0 + 0.35 * (m + n)


def wants_ct(k, cin, cout, x3):
    """conv_halo.cu:halo_wants_ct"""
    if k <= 1 or cout > 128 or cout % 16 or k * cout > 256:
        return False
    ks, groups, passes = (cin + 7) // 8, (cout + 31) // 32, 2.0 if x3 else 1.0
    plain = passes * k * k * ks * mma_clk(128, cout)
    ct = passes * k * ks * mma_clk(128, k * cout)
    ep, ec = 1500.0 * groups, (1500.0 + 300.0 * (k - 1)) * groups
    return max(ct, ec) < 0.95 * max(plain, ep)


def layer_bounds(N, h, cin, cout, k, pad, x3):
    ho = h + 2 * pad - k + 1
    twv = 17 - k
    tiles = N * -(-ho // 8) * -(-ho // twv)                     # 128-row GEMM tiles (8 rows x 16 columns, twv valid)
    cout_tiles = -(-cout // 128)
    n_mma = min(cout, 128)
    n_mma = (n_mma + 15) // 16 * 16
    ks = (cin + 7) // 8
    passes = 2 if x3 else 1
    ct = wants_ct(k, cin, cout, x3)
    per_tile = passes * (k * ks * mma_clk(128, k * cout) if ct else k * k * ks * mma_clk(128, n_mma))
    waves = -(-(tiles * cout_tiles) // SMS)
    t_mma = waves * per_tile / (CLK_GHZ * 1e3)                  # us
    return ct, t_mma


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else 'profiles/r2_conv_micro_final.txt'
    pat = re.compile(r'N(\d+) (\d+)x\d+ (\d+)->(\d+) k(\d) p(\d)( \+res)?\s+([\d.]+)\s+([\d.]+|nan)\s+([\d.]+)\s+([\d.]+|nan)\s+([\d.]+)')
    print('%-34s | %22s | %30s' % ('layer', '1xTF32 halo kernel', 'reference precision (tf32x3)'))
    print('%-34s | %7s %7s %6s | %3s %7s %7s %7s %6s' % ('', 'us', 'MMA us', 'x', 'CT', 'us', 'MMA us', 'HBM us', 'x'))
    for line in open(path):
        m = pat.match(line.strip())
        if not m:
            continue
        N, h, cin, cout, k, pad = (int(m.group(i)) for i in range(1, 7))
        res = bool(m.group(7))
        halo, halo3, hbm = m.group(9), m.group(11), float(m.group(12))
        _, m1 = layer_bounds(N, h, cin, cout, k, pad, False)
        ct3, m3 = layer_bounds(N, h, cin, cout, k, pad, True)
        name = 'N%d %dx%d %d->%d k%d%s' % (N, h, h, cin, cout, k, ' +res' if res else '')
        f1 = float(halo) / max(m1, hbm) if halo != 'nan' else float('nan')
        f3 = float(halo3) / max(m3, hbm) if halo3 != 'nan' else float('nan')
        print('%-34s | %7s %7.1f %6.2f | %3s %7s %7.1f %7.1f %6.2f' % (name, halo, m1, f1, 'yes' if ct3 else 'no', halo3, m3, hbm, f3))
    print('\nx = measured / max(MMA-phase bound, HBM floor).  Values near 1: the kernel runs at the bound the model names; larger:'
          '\nsomething else limits (weight ring over L2->SM for the column-taps variant of 48->48 in reference precision, the epilogue,'
          '\nper-tile fixed costs of tiny-K layers).')


if __name__ == '__main__':
    main()
