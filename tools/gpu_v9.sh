#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_3_tc.py -k "halo" -q --tb=short --timeout 120 -p no:cacheprovider -x 2>&1 | tail -5
timeout 400 python tools/conv_micro.py > gpurun_out/conv_micro.txt 2>&1
cat gpurun_out/conv_micro.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 2 -c 1 -f -o gpurun_out/prof_halo48r python tools/conv_micro.py 0 > gpurun_out/ncu_halo.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 2 -c 1 -f -o gpurun_out/prof_halo48 python tools/conv_micro.py 1 >> gpurun_out/ncu_halo.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wgrad_halo -s 2 -c 1 -f -o gpurun_out/prof_whalo48 python tools/wgrad_micro.py 0 >> gpurun_out/ncu_halo.log 2>&1
