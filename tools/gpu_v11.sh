#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_7_dist.py tests/test_gpu_8_res256.py -q --tb=short --timeout 280 -p no:cacheprovider -s 2>&1 | tail -30 > gpurun_out/tests_78.log
tail -12 gpurun_out/tests_78.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_default.err; cut -c1-1500 gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_ref.json
