#!/bin/bash
# round-2 visit 1: halo-window probe, GPU suite state, step anatomy at 256x256
mkdir -p gpurun_out
timeout 180 python tools/halo_probe.py > gpurun_out/halo_probe.txt 2>&1; echo "rc=$?" >> gpurun_out/halo_probe.txt
timeout 700 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/tests.log
timeout 300 python tools/step_profile.py --config taichi --res 256 --batch 8 --top 60 --out gpurun_out/step_taichi256.md > /dev/null 2> gpurun_out/step_taichi256.err
timeout 300 python tools/step_profile.py --config shapes --res 64 --batch 32 --top 60 --out gpurun_out/step_shapes64.md > /dev/null 2> gpurun_out/step_shapes64.err
tail -30 gpurun_out/halo_probe.txt; tail -5 gpurun_out/tests.log; head -12 gpurun_out/step_taichi256.md
