#!/bin/bash
# ncu of the k_harris variants (set SFM_HARRIS_VARIANT) on one 32 x 1080p extraction call
mkdir -p gpurun_out
export SFM_HARRIS_VARIANT=${1:-3}
tag=${2:-hs}
timeout 120 python scripts/prof_extract.py 32 2 > gpurun_out/${tag}_plain.log 2>&1 || { cat gpurun_out/${tag}_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_harris' -s 4 -c 4 -o gpurun_out/${tag}_full -f python scripts/prof_extract.py 32 2 > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${tag}_ncu.log
