#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -s 2>&1 | tail -150 > gpurun_out/tests.log
grep smoke gpurun_out/smoke.log; tail -5 gpurun_out/tests.log
