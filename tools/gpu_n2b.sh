#!/bin/bash
# 2-GPU visit at the final commit: on-GPU N-rank == 1-rank invariant (test + bench pre-flight), weak scaling of the headline step
mkdir -p gpurun_out
timeout 280 python -m pytest tests/test_gpu_7_dist.py -q -s --tb=short -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/tests_dist.log; tail -4 gpurun_out/tests_dist.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"; tail -2 gpurun_out/bench_n2.err; cut -c1-300 gpurun_out/bench_n2.json
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-kernel-bench > gpurun_out/bench_n1_same_box.json 2> gpurun_out/bench_n1_same_box.err; cut -c1-300 gpurun_out/bench_n1_same_box.json
