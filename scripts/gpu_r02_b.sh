#!/bin/bash
# k_harris variants: parity (R bit-exact tests) and timing
mkdir -p gpurun_out
for v in 0 1 2; do
  export SFM_HARRIS_VARIANT=$v
  echo "=== variant $v" >> gpurun_out/r02b.log
  timeout 600 python -m pytest tests/test_gpu_extract.py -x -q 2>&1 | tail -3 >> gpurun_out/r02b.log
  timeout 300 python scripts/time_extract.py 32 20 >> gpurun_out/r02b.log 2>&1
done
cat gpurun_out/r02b.log
