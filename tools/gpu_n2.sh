#!/bin/bash
# 2-GPU visit (charged x2: keep it short, every command under its own tight timeout)
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
cut -c1-200 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err | cut -c1-300
