#!/bin/bash
# Profile visit: everything the summaries under profiles/ are made from.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/tests.log
timeout 300 python tools/step_profile.py --config shapes --res 64 --batch 32 --out gpurun_out/step_shapes64.md > gpurun_out/step_shapes64.log 2>&1
timeout 300 python tools/prof_kernels.py grid > gpurun_out/grid_kernels.txt 2>&1
# launch list of the bench command itself (eager so that every kernel is a separate launch record)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1100 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --graph off --no-cpu-baseline --no-kernel-bench --no-transfer > gpurun_out/ncu_bench.log 2>&1
# full captures: grid_sample forward on the 64ch x 128^2 level (8th..9th forward launch of prof_kernels grid), the conv kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_grid_sample_fwd -s 7 -c 2 -f -o gpurun_out/prof_grid python tools/prof_kernels.py grid > gpurun_out/ncu_grid.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 0 -c 4 -f -o gpurun_out/prof_convtc python tools/prof_kernels.py convtc > gpurun_out/ncu_convtc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wgrad_tc -s 0 -c 4 -f -o gpurun_out/prof_wgradtc python tools/prof_kernels.py convtc > gpurun_out/ncu_wgradtc.log 2>&1
tail -2 gpurun_out/tests.log; cat gpurun_out/grid_kernels.txt; tail -2 gpurun_out/ncu_convtc.log
