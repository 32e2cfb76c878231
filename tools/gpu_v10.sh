#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -x 2>&1 | tail -25 > gpurun_out/tests.log
tail -6 gpurun_out/tests.log
timeout 300 python tools/conv_micro.py 0 1 2 5 > gpurun_out/conv_micro_v4.txt 2>&1; cat gpurun_out/conv_micro_v4.txt
timeout 300 python tools/wgrad_micro.py 0 1 2 > gpurun_out/wgrad_micro_v2.txt 2>&1; cat gpurun_out/wgrad_micro_v2.txt
timeout 300 python tools/step_profile.py --config taichi --res 256 --batch 8 --top 70 --out gpurun_out/step_taichi256_auto.md > /dev/null 2> gpurun_out/step_taichi256.err
head -50 gpurun_out/step_taichi256_auto.md
MONKEY_B200_CONV=tf32 timeout 300 python tools/step_profile.py --config taichi --res 256 --batch 8 --top 70 --out gpurun_out/step_taichi256_tf32.md > /dev/null 2>> gpurun_out/step_taichi256.err
head -12 gpurun_out/step_taichi256_tf32.md
