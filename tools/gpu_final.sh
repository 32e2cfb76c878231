#!/bin/bash
# what the driver runs at round end, in one visit: smoke, GPU tests, the default bench line
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/tests.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?" >> gpurun_out/bench_default.err
grep smoke gpurun_out/smoke.log; tail -3 gpurun_out/tests.log; cut -c1-250 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
