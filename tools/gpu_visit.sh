#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -s "$@" 2>&1 | tail -40 > gpurun_out/tests.log
timeout 300 python tools/prof_kernels.py grid > gpurun_out/grid_kernels.txt 2>&1
timeout 300 python tools/prof_kernels.py convtc > gpurun_out/convtc_kernels.txt 2>&1
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -4 gpurun_out/tests.log; cat gpurun_out/grid_kernels.txt gpurun_out/convtc_kernels.txt; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
