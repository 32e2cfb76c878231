#!/bin/bash
# 2-GPU visit (charged x2: keep it short, every command under its own tight timeout)
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 150 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-bench --no-transfer > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-bench --no-transfer > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/smoke.log; cut -c1-200 gpurun_out/bench_n1.json; cut -c1-200 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err | cut -c1-300
