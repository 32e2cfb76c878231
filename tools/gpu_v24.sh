#!/bin/bash
# visit 24: pack kernels A/B (8 x 32 tiles vs elementwise)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_3_tc.py tests/test_gpu_2_modules.py -q --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/tests_tc.log; tail -3 gpurun_out/tests_tc.log
timeout 300 python tools/step_profile.py --config taichi --res 256 --batch 8 --top 3 --out gpurun_out/step_pack_tiled.md > /dev/null 2> gpurun_out/step.err; grep -i "step (CUDA\|pack" gpurun_out/step_pack_tiled.md
MONKEY_B200_PACK_TILED=0 timeout 300 python tools/step_profile.py --config taichi --res 256 --batch 8 --top 3 --out gpurun_out/step_pack_elem.md > /dev/null 2> gpurun_out/step.err; grep -i "step (CUDA\|pack" gpurun_out/step_pack_elem.md
