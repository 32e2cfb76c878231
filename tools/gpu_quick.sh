#!/bin/bash
# quick visit: tests + bench (tf32 + CUDA graph, and fp32) + launch list of the eager tf32 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -s 2>&1 | tail -40 > gpurun_out/tests.log
MONKEY_B200_CONV=tf32 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_tf32_graph.json 2> gpurun_out/bench_tf32_graph.err
MONKEY_B200_CONV=fp32 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-bench > gpurun_out/bench_fp32_graph.json 2> gpurun_out/bench_fp32_graph.err
MONKEY_B200_CONV=tf32 timeout 600 python bench.py --steps 10 --warmup 3 --graph off --no-cpu-baseline --no-kernel-bench > gpurun_out/bench_tf32_eager.json 2> gpurun_out/bench_tf32_eager.err
MONKEY_B200_CONV=tf32 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1800 --csv --log-file gpurun_out/launches_tf32.csv \
    python bench.py --steps 1 --warmup 1 --graph off --no-cpu-baseline --no-kernel-bench > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/tests.log; cut -c1-300 gpurun_out/bench_tf32_graph.json; tail -3 gpurun_out/bench_tf32_graph.err
