#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 2 -c 1 -f -o gpurun_out/prof_halo48 python tools/conv_micro.py 1 > gpurun_out/ncu_halo.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 2 -c 1 -f -o gpurun_out/prof_halo128 python tools/conv_micro.py 3 >> gpurun_out/ncu_halo.log 2>&1
MONKEY_B200_CONV_HALO=0 timeout 300 python -m pytest tests/test_gpu_2_modules.py -q --tb=short --timeout 300 -p no:cacheprovider -k "golden" 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/tests.log
tail -8 gpurun_out/tests.log
