#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_3_tc.py -k "halo" -q --tb=short --timeout 120 -p no:cacheprovider -x 2>&1 | tail -5
timeout 300 python tools/conv_micro.py 0 1 2 3 5 12 14 > gpurun_out/conv_micro_v5.txt 2>&1; cat gpurun_out/conv_micro_v5.txt
timeout 500 python tools/diag_train256.py > gpurun_out/diag_train256.txt 2>&1; tail -6 gpurun_out/diag_train256.txt
