#!/bin/bash
# 8-GPU visit: pre-flight invariant + headline step at 8 ranks (peer-memory BN statistics), then the launch-bound 64x64 step with the
# peer-memory kernel and with NCCL statistics
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $R --master-port 29500 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "n8 rc=$?"; tail -3 gpurun_out/bench_n8.err; cut -c1-600 gpurun_out/bench_n8.json
timeout 120 $R --master-port 29501 bench.py --gpus 8 --steps 20 --warmup 5 --config shapes --res 64 --batch 32 > gpurun_out/bench_n8_shapes_p2p.json 2> gpurun_out/bench_n8_shapes_p2p.err; echo "n8 shapes p2p rc=$?"; cut -c1-300 gpurun_out/bench_n8_shapes_p2p.json
MONKEY_B200_BN_P2P=0 timeout 120 $R --master-port 29502 bench.py --gpus 8 --steps 20 --warmup 5 --config shapes --res 64 --batch 32 > gpurun_out/bench_n8_shapes_nccl.json 2> gpurun_out/bench_n8_shapes_nccl.err; echo "n8 shapes nccl rc=$?"; cut -c1-300 gpurun_out/bench_n8_shapes_nccl.json
