#!/bin/bash
# GPU visit: parity tests, kernel micro-benchmarks, ncu captures of the grid_sample and conv kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -s "$@" 2>&1 | tail -40 > gpurun_out/tests.log
timeout 300 python tools/prof_kernels.py grid > gpurun_out/grid_kernels.txt 2>&1
timeout 300 python tools/prof_kernels.py convtc > gpurun_out/convtc_kernels.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_grid_sample_fwd -s 14 -c 2 -f -o gpurun_out/prof_grid python tools/prof_kernels.py grid > gpurun_out/ncu_grid.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 3 -c 1 -f -o gpurun_out/prof_convtc python tools/prof_kernels.py convtc > gpurun_out/ncu_convtc.log 2>&1
tail -4 gpurun_out/tests.log; cat gpurun_out/grid_kernels.txt gpurun_out/convtc_kernels.txt; tail -2 gpurun_out/ncu_grid.log
