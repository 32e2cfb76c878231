#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_3_tc.py -k "halo" -q --tb=short --timeout 120 -p no:cacheprovider -x 2>&1 | tail -4
echo "== default planner"; timeout 300 python tools/conv_micro.py 0 1 2 3 5 12 > gpurun_out/conv_micro_v6.txt 2>&1; cat gpurun_out/conv_micro_v6.txt
echo "== RB=2 forced"; MONKEY_B200_HALO_RB=2 timeout 300 python tools/conv_micro.py 0 1 2 3 5 12 2>&1 | tail -6
echo "== RB=4 forced"; MONKEY_B200_HALO_RB=4 timeout 300 python tools/conv_micro.py 0 1 2 5 2>&1 | tail -4
echo "== RB=1 forced"; MONKEY_B200_HALO_RB=1 timeout 300 python tools/conv_micro.py 0 1 2 5 2>&1 | tail -4
timeout 300 python tools/wgrad_micro.py 0 1 2 4 9 > gpurun_out/wgrad_micro_v3.txt 2>&1; cat gpurun_out/wgrad_micro_v3.txt
