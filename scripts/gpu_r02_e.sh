#!/bin/bash
# extraction tests + timing after a kernel change
mkdir -p gpurun_out
: > gpurun_out/r02e.log
timeout -k 5 60 python scripts/time_extract.py 32 20 >> gpurun_out/r02e.log 2>&1
timeout -k 5 60 python scripts/time_extract.py 3 5 240 320 >> gpurun_out/r02e.log 2>&1
timeout -k 5 200 python -m pytest tests/test_gpu_extract.py tests/test_gpu_bench_shape.py -x -q 2>&1 | tail -5 >> gpurun_out/r02e.log
cat gpurun_out/r02e.log
