#!/bin/bash
# k_harris_stream (variant 3): parity and timing, each under its own timeout
mkdir -p gpurun_out
: > gpurun_out/r02c.log
for v in ${VARIANTS:-3}; do
  export SFM_HARRIS_VARIANT=$v
  echo "=== variant $v" >> gpurun_out/r02c.log
  timeout 120 python scripts/time_extract.py 32 20 >> gpurun_out/r02c.log 2>&1
  echo "time rc=$?" >> gpurun_out/r02c.log
  timeout 120 python scripts/time_extract.py 3 5 240 320 >> gpurun_out/r02c.log 2>&1
  timeout 600 python -m pytest tests/test_gpu_extract.py tests/test_gpu_bench_shape.py -x -q 2>&1 | tail -5 >> gpurun_out/r02c.log
done
cat gpurun_out/r02c.log
