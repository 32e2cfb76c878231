#!/bin/bash
# visit 26: conv weight gradients accumulated by the unpack kernel straight into the optimiser's buffers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_4_graph.py tests/test_gpu_2_modules.py tests/test_gpu_8_res256.py tests/test_gpu_6_reference_drivers.py -q --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/tests_tc.log; tail -3 gpurun_out/tests_tc.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; grep -i "smoke" gpurun_out/smoke.log | tail -4
timeout 400 python bench.py --no-extras --no-cpu-baseline --no-kernel-bench > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_quick.json
timeout 300 python tools/step_profile.py --config taichi --res 256 --batch 8 --top 12 --out gpurun_out/step_direct.md > /dev/null 2> gpurun_out/step.err; sed -n 1,24p gpurun_out/step_direct.md
