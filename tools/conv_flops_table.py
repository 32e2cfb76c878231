"""Regenerates SURVEY 8(a)'s per-sample conv-GFLOP table programmatically (2*MACs of every conv of the oracle, hooked;
train step = 3*KP2 + 3*G + 12*D) - the algorithmic FLOPs `bench.py`'s roofline uses.  CPU only.

    python tools/conv_flops_table.py            # markdown table for the configurations of BASELINE.json / SURVEY 8(a)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ROWS = [('shapes', 64), ('taichi', 64), ('taichi', 256), ('moving-gif', 128), ('moving-gif', 256), ('vox-full', 256)]


def main():
    print('| config @res | KP (D=2) | G | D (1 pass) | inference / frame (KP1+G) | train step / sample |')
    print('|---|---:|---:|---:|---:|---:|')
    for name, res in ROWS:
        f = bench.conv_flops_per_step(bench.load_config(name), res, 1)
        print('| %s @%d | %.3f | %.3f | %.3f | %.2f | %.1f |' % (name, res, f['kp2'] / 1e9, f['g'] / 1e9, f['d'] / 1e9,
                                                              (f['kp2'] / 2 + f['g']) / 1e9,
                                                              f['train_step_per_sample'] / 1e9))


if __name__ == '__main__':
    main()
