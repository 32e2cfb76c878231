#!/bin/bash
# visit 21: epilogue as passes over the 32 channels (one uniform branch per pass)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_3_tc.py -q --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/tests_tc.log; tail -3 gpurun_out/tests_tc.log
timeout 200 python tools/conv_micro.py 2>&1 | cut -c1-110 > gpurun_out/conv_micro.txt; cat gpurun_out/conv_micro.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; grep -i "smoke" gpurun_out/smoke.log | tail -8
timeout 400 python bench.py --no-extras --no-cpu-baseline --no-kernel-bench > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_quick.json
MONKEY_B200_CONV=tf32 timeout 400 python bench.py --no-extras --no-cpu-baseline --no-kernel-bench > gpurun_out/bench_quick_tf32.json 2> gpurun_out/bench_quick_tf32.err; cut -c1-300 gpurun_out/bench_quick_tf32.json
