#!/bin/bash
# one ncu --set full capture of k_describe on the 32 x 1080p call (run only after the same command exited 0 without ncu)
mkdir -p gpurun_out
timeout 120 python scripts/time_extract.py 32 2 > gpurun_out/ncu_desc_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_describe -s 1 -c 1 -f -o gpurun_out/r02_describe \
    python scripts/time_extract.py 32 2 > gpurun_out/ncu_desc.log 2>&1
tail -3 gpurun_out/ncu_desc.log
