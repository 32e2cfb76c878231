#!/bin/bash
# visit 18: column taps stacked on N (CT) in k_conv_halo - parity, A/B against CT off; then the full suite, ncu, step anatomy, full bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_3_tc.py -q --tb=short --timeout 300 -p no:cacheprovider -x 2>&1 | tail -40 > gpurun_out/tests_tc.log; tail -5 gpurun_out/tests_tc.log
if ! grep -q "passed" gpurun_out/tests_tc.log || grep -q "failed\|error" gpurun_out/tests_tc.log; then echo "kernel tests failed: CT off for the rest"; export MONKEY_B200_HALO_CT=0; fi
echo "== CT on" > gpurun_out/ct_ab.txt; MONKEY_B200_HALO_CT=1 timeout 200 python tools/conv_micro.py 2>&1 | cut -c1-110 >> gpurun_out/ct_ab.txt
echo "== CT off" >> gpurun_out/ct_ab.txt; MONKEY_B200_HALO_CT=0 timeout 200 python tools/conv_micro.py 0 1 2 4 5 7 8 9 11 16 17 18 19 2>&1 | cut -c1-110 >> gpurun_out/ct_ab.txt
head -24 gpurun_out/ct_ab.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider --deselect tests/test_gpu_3_tc.py 2>&1 | tail -40 > gpurun_out/tests.log; tail -4 gpurun_out/tests.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 9 -c 1 -f -o gpurun_out/prof_halo48_x3b python tools/conv_micro.py 1 > gpurun_out/ncu_x3b.log 2>&1; tail -2 gpurun_out/ncu_x3b.log
timeout 300 python tools/step_profile.py --config taichi --res 256 --batch 8 --top 70 --out gpurun_out/step_taichi256_auto_v3.md > /dev/null 2> gpurun_out/step.err; head -12 gpurun_out/step_taichi256_auto_v3.md
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_default.json
