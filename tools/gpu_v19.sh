#!/bin/bash
# visit 19: CT criterion + fallback, separate weight-ring producer, 3 staging buffers, chunked keypoint head
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_3_tc.py tests/test_gpu_1_ops.py -q --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/tests_tc.log; tail -5 gpurun_out/tests_tc.log
echo "== default" > gpurun_out/ct_ab.txt; timeout 200 python tools/conv_micro.py 2>&1 | cut -c1-110 >> gpurun_out/ct_ab.txt
echo "== NSTG=2" >> gpurun_out/ct_ab.txt; MONKEY_B200_HALO_NSTG=2 timeout 200 python tools/conv_micro.py 0 1 2 3 9 12 13 14 15 2>&1 | cut -c1-110 >> gpurun_out/ct_ab.txt
cat gpurun_out/ct_ab.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 2 -c 1 -f -o gpurun_out/prof_halo_32_128 python tools/conv_micro.py 3 > gpurun_out/ncu_32_128.log 2>&1; tail -2 gpurun_out/ncu_32_128.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; grep -i "smoke" gpurun_out/smoke.log | tail -8
timeout 400 python bench.py --no-extras --no-cpu-baseline --no-kernel-bench > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_quick.json
MONKEY_B200_CONV=tf32 timeout 400 python bench.py --no-extras --no-cpu-baseline --no-kernel-bench > gpurun_out/bench_quick_tf32.json 2> gpurun_out/bench_quick_tf32.err; cut -c1-300 gpurun_out/bench_quick_tf32.json
timeout 300 python tools/step_profile.py --config taichi --res 256 --batch 8 --top 70 --out gpurun_out/step_taichi256_auto_v4.md > /dev/null 2> gpurun_out/step.err; head -30 gpurun_out/step_taichi256_auto_v4.md
