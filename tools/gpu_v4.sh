#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 2 -c 1 -f -o gpurun_out/prof_halo python tools/conv_micro.py 1 > gpurun_out/ncu_halo.log 2>&1
tail -3 gpurun_out/ncu_halo.log
ls -la gpurun_out/*.ncu-rep
