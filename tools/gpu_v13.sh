#!/bin/bash
mkdir -p gpurun_out
echo "== cp.async.ca (default)"; timeout 200 python tools/prof_kernels.py grid 2>&1 | grep fwd
echo "== cp.async.cg"; MONKEY_B200_GS_CA=0 timeout 200 python tools/prof_kernels.py grid 2>&1 | grep fwd
timeout 300 python -m pytest tests/test_gpu_1_ops.py -q -k "grid_sample" --tb=short -p no:cacheprovider 2>&1 | tail -3
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 9 -c 1 -f -o gpurun_out/prof_halo48_x3 python tools/conv_micro.py 1 > gpurun_out/ncu_x3.log 2>&1; tail -2 gpurun_out/ncu_x3.log
