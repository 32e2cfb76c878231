#!/bin/bash
# visit 20: epilogue affine vectors from shared memory; A/B of the residual case; full suite; full bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_3_tc.py -q --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/tests_tc.log; tail -3 gpurun_out/tests_tc.log
echo "== default" > gpurun_out/ct_ab.txt; timeout 200 python tools/conv_micro.py 2>&1 | cut -c1-110 >> gpurun_out/ct_ab.txt
echo "== CT=2 (forced)" >> gpurun_out/ct_ab.txt; MONKEY_B200_HALO_CT=2 timeout 200 python tools/conv_micro.py 0 1 3 5 8 12 16 17 2>&1 | cut -c1-110 >> gpurun_out/ct_ab.txt
echo "== CT=0" >> gpurun_out/ct_ab.txt; MONKEY_B200_HALO_CT=0 timeout 200 python tools/conv_micro.py 0 1 2 4 9 11 2>&1 | cut -c1-110 >> gpurun_out/ct_ab.txt
echo "== NSTG=1" >> gpurun_out/ct_ab.txt; MONKEY_B200_HALO_NSTG=1 timeout 200 python tools/conv_micro.py 0 1 3 2>&1 | cut -c1-110 >> gpurun_out/ct_ab.txt
cat gpurun_out/ct_ab.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider --deselect tests/test_gpu_3_tc.py 2>&1 | tail -30 > gpurun_out/tests.log; tail -4 gpurun_out/tests.log
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_default.json
