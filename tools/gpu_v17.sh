#!/bin/bash
# visit 17: full GPU suite with the TF32 + BF16-cross scheme, RB sweep of the halo planner, ncu of the dominant kernel, full bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/tests.log; tail -4 gpurun_out/tests.log
for RB in 1 2 4; do echo "== RB=$RB"; MONKEY_B200_HALO_RB=$RB timeout 120 python tools/conv_micro.py 0 1 2 3 4 5 6 13 14 15 2>&1 | cut -c1-100; done > gpurun_out/rb_sweep.txt 2>&1; grep -c "256x256" gpurun_out/rb_sweep.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 9 -c 1 -f -o gpurun_out/prof_halo48_x3b python tools/conv_micro.py 1 > gpurun_out/ncu_x3b.log 2>&1; tail -2 gpurun_out/ncu_x3b.log
timeout 300 python tools/step_profile.py --config taichi --res 256 --batch 8 --top 70 --out gpurun_out/step_taichi256_auto_v3.md > /dev/null 2> gpurun_out/step.err; head -12 gpurun_out/step_taichi256_auto_v3.md
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_default.json
