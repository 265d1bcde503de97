#!/bin/bash
# Refresh of the extraction half of scripts/profile_r02.sh (the matcher captures stand when the matcher did not change):
# ncu launch list of a 2-step bench + ncu --set full of the 9 launches of one 32 x 1080p extraction call.
tag=${1:-r02b}
out=gpurun_out/prof_$tag
mkdir -p $out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu --no-4k --no-geometry --no-all-pairs"
timeout -k 5 120 $BENCH > $out/plain_bench.log 2>&1 &&
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/launches_bench_steps2.csv $BENCH > $out/ncu_launches.log 2>&1
echo "launch list rc=$?"
EXT="python scripts/prof_extract.py 32 2"
timeout -k 5 60 $EXT > $out/plain_extract.log 2>&1 &&
timeout -k 5 200 ncu --set full --clock-control none --import-source on -k 'regex:^k_' -s 9 -c 9 -o $out/extract_full -f $EXT > $out/ncu_extract.log 2>&1
echo "extract full rc=$?"
ls -la $out | head -8
