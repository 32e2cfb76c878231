#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_3_tc.py -k "halo" -q --tb=short --timeout 120 -p no:cacheprovider -x 2>&1 | tail -30 > gpurun_out/halo_tests.log
tail -3 gpurun_out/halo_tests.log
if grep -q "passed" gpurun_out/halo_tests.log && ! grep -q "failed" gpurun_out/halo_tests.log; then
  timeout 400 python tools/conv_micro.py > gpurun_out/conv_micro.txt 2>&1
  cat gpurun_out/conv_micro.txt
  timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -s 2>&1 | tail -60 > gpurun_out/tests.log
  tail -5 gpurun_out/tests.log
  timeout 300 python tools/step_profile.py --config taichi --res 256 --batch 8 --top 60 --out gpurun_out/step_taichi256_v3.md > /dev/null 2> gpurun_out/step_taichi256.err
  head -8 gpurun_out/step_taichi256_v3.md
fi
