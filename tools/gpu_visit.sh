#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -s 2>&1 | tail -30 > gpurun_out/tests.log
timeout 200 python bench.py --steps 20 --warmup 5 --adam flat --no-cpu-baseline --no-kernel-bench --no-transfer > gpurun_out/bench_flat.json 2> gpurun_out/bench_flat.err
timeout 200 python bench.py --steps 20 --warmup 5 --adam torch --no-cpu-baseline --no-kernel-bench --no-transfer > gpurun_out/bench_torch.json 2> gpurun_out/bench_torch.err
tail -3 gpurun_out/tests.log; cut -c1-220 gpurun_out/bench_flat.json; cut -c1-220 gpurun_out/bench_torch.json; tail -2 gpurun_out/bench_flat.err
