#!/bin/bash
# visit 22: wave-aware pixel splits of the weight-gradient kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_3_tc.py -q --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/tests_tc.log; tail -3 gpurun_out/tests_tc.log
timeout 200 python tools/wgrad_micro.py > gpurun_out/wgrad_micro.txt 2>&1; cat gpurun_out/wgrad_micro.txt | cut -c1-120
timeout 400 python bench.py --no-extras --no-cpu-baseline --no-kernel-bench > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_quick.json
timeout 300 python tools/step_profile.py --config taichi --res 256 --batch 8 --top 90 --out gpurun_out/step_taichi256_auto_v5.md > /dev/null 2> gpurun_out/step.err; head -14 gpurun_out/step_taichi256_auto_v5.md
