#!/bin/bash
# 8-GPU visit at the final commit: pre-flight invariant + headline step at 8 ranks
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $R --master-port 29500 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "n8 rc=$?"; tail -2 gpurun_out/bench_n8.err; cut -c1-400 gpurun_out/bench_n8.json
