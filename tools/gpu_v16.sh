#!/bin/bash
# visit 16: TF32 + BF16-cross reference-precision scheme (2 MMAs per K step) in conv_halo / conv_tc / wgrad_halo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_3_tc.py -q --tb=short --timeout 300 -p no:cacheprovider -x 2>&1 | tail -60 > gpurun_out/tests_tc.log; tail -5 gpurun_out/tests_tc.log
timeout 200 python tools/conv_micro.py > gpurun_out/conv_micro.txt 2>&1; head -8 gpurun_out/conv_micro.txt
timeout 200 python tools/wgrad_micro.py > gpurun_out/wgrad_micro.txt 2>&1; head -8 gpurun_out/wgrad_micro.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; grep -i "smoke\|err\|loss" gpurun_out/smoke.log | tail -8
timeout 400 python bench.py --no-extras --no-cpu-baseline --no-kernel-bench > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_quick.json
