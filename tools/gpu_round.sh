#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list of the same command, ncu full captures.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python tools/prof_kernels.py all > gpurun_out/kernels.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 2500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-bench > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_grid_sample_fwd -s 4 -c 2 -f -o gpurun_out/prof_grid \
    python tools/prof_kernels.py grid > gpurun_out/ncu_grid.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_ffma -s 10 -c 2 -f -o gpurun_out/prof_conv \
    python tools/prof_kernels.py conv > gpurun_out/ncu_conv.log 2>&1
tail -2 gpurun_out/tests.log; cat gpurun_out/bench_n1.json | cut -c1-600
