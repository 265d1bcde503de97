#!/bin/bash
# round 2, first GPU pass: tests, smoke, default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r02a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r02a_tests.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r02a_smoke.log
timeout 600 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench exit $?" >> gpurun_out/r02a_bench.err
tail -3 gpurun_out/r02a_tests.log; tail -2 gpurun_out/r02a_smoke.log; tail -3 gpurun_out/r02a_bench.err
