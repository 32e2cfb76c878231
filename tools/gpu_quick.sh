#!/bin/bash
# quick visit: tests + bench in both conv modes + kernel timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/tests.log
MONKEY_B200_CONV=tf32 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
timeout 300 python tools/prof_kernels.py grid > gpurun_out/kernels_grid.txt 2>&1
MONKEY_B200_CONV=tf32 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 2500 --csv --log-file gpurun_out/launches_tf32.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-bench > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/tests.log; cut -c1-400 gpurun_out/bench_tf32.json
