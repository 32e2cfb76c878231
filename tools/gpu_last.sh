#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_2_modules.py tests/test_gpu_4_graph.py -m gpu -x -q --timeout 90 -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/last_tests.log
MONKEY_B200_CONV_HALO=1 timeout 60 python -m pytest tests/test_gpu_3_tc.py -m gpu -q -k halo --timeout 50 -p no:cacheprovider --tb=line 2>&1 | tail -12 > gpurun_out/halo_tests.log
tail -2 gpurun_out/last_tests.log; cat gpurun_out/halo_tests.log
