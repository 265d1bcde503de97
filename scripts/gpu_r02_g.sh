#!/bin/bash
# full GPU test suite + smoke + default bench + profile pass
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r02g_tests.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02g_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02g_smoke.log
tail -3 gpurun_out/r02g_tests.log; tail -2 gpurun_out/r02g_smoke.log
bash scripts/profile_r02.sh r02b bench
