#!/bin/bash
# ncu --set full of kernels matching $1 in one 32 x 1080p extraction call -> gpurun_out/$2_full.ncu-rep
mkdir -p gpurun_out
k=${1:-k_nms}; tag=${2:-nms}; cnt=${3:-4}
timeout 120 python scripts/prof_extract.py 32 2 > gpurun_out/${tag}_plain.log 2>&1 || { cat gpurun_out/${tag}_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -s $cnt -c $cnt -o gpurun_out/${tag}_full -f python scripts/prof_extract.py 32 2 > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${tag}_ncu.log
