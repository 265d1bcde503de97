#!/bin/bash
# final check A: full GPU test suite + smoke + default bench
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests -m gpu -x -q > gpurun_out/r02fa_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r02fa_tests.log
timeout -k 5 90 python __graft_entry__.py smoke > gpurun_out/r02fa_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02fa_smoke.log
tail -3 gpurun_out/r02fa_tests.log; tail -2 gpurun_out/r02fa_smoke.log
timeout -k 5 240 python bench.py > gpurun_out/r02fa_bench.json 2> gpurun_out/r02fa_bench.err; echo "bench exit $?"
cat gpurun_out/r02fa_bench.json
