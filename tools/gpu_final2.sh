#!/bin/bash
# final visit of round 2: what the driver runs (tests, smoke, bench both arms) + the profile artefacts behind profiles/
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; grep smoke gpurun_out/smoke.log | tail -8
timeout 1200 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/tests.log; tail -4 gpurun_out/tests.log
timeout 200 python tools/conv_micro.py 2>&1 | cut -c1-110 > gpurun_out/conv_micro.txt
timeout 200 python tools/wgrad_micro.py 2>&1 | cut -c1-110 > gpurun_out/wgrad_micro.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 9 -c 1 -f -o gpurun_out/prof_halo48_ct_x3 python tools/conv_micro.py 1 > gpurun_out/ncu_ct_x3.log 2>&1; tail -2 gpurun_out/ncu_ct_x3.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --graph off --no-extras --no-cpu-baseline --no-kernel-bench > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200; wc -l gpurun_out/launches_bench.csv
timeout 300 python tools/step_profile.py --config taichi --res 256 --batch 8 --top 90 --out gpurun_out/step_taichi256_auto_final.md > /dev/null 2> gpurun_out/step.err; head -12 gpurun_out/step_taichi256_auto_final.md
timeout 300 python tools/step_profile.py --config shapes --res 64 --batch 32 --top 40 --out gpurun_out/step_shapes64_auto_final.md > /dev/null 2>> gpurun_out/step.err; head -6 gpurun_out/step_shapes64_auto_final.md
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_default.json
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_reference.json
