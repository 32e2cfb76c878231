#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_3_tc.py -k "conv_halo" -q --tb=short --timeout 120 -p no:cacheprovider -x 2>&1 | tail -30 > gpurun_out/halo_tests.log
tail -3 gpurun_out/halo_tests.log
if grep -q "passed" gpurun_out/halo_tests.log && ! grep -q "failed" gpurun_out/halo_tests.log; then
  timeout 400 python tools/conv_micro.py > gpurun_out/conv_micro.txt 2>&1
  cat gpurun_out/conv_micro.txt
fi
timeout 300 python -m pytest tests/test_gpu_3_tc.py -k "wgrad_halo" -q --tb=line --timeout 120 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/whalo_tests.log
tail -30 gpurun_out/whalo_tests.log
timeout 400 python tools/wgrad_micro.py > gpurun_out/wgrad_micro.txt 2>&1
cat gpurun_out/wgrad_micro.txt
