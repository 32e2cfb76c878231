#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; grep smoke gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/tests.log; tail -4 gpurun_out/tests.log
timeout 300 python tools/step_profile.py --config taichi --res 256 --batch 8 --top 70 --out gpurun_out/step_taichi256_auto_v2.md > /dev/null 2> gpurun_out/step.err; head -16 gpurun_out/step_taichi256_auto_v2.md
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_default.json
