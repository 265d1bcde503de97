#!/bin/bash
# run every diagnostic stage under its own timeout; outputs to gpurun_out/diag_<stage>.log
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv | tee gpurun_out/diag_gpu.log
for st in "$@"; do
  echo "=== stage $st"
  timeout 300 python scripts/gpu_diag.py $st > gpurun_out/diag_$st.log 2>&1
  echo "exit $?"; tail -40 gpurun_out/diag_$st.log
done
