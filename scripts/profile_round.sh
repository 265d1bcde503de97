#!/bin/bash
# Round profile under gpurun (one B200): plain bench, ncu launch list of the same command, and
# ncu --set full captures of the hot kernels.  Outputs under gpurun_out/prof_<tag>/.
#   gpurun --timeout 1500 -- 'bash scripts/profile_round.sh r01b'
tag=${1:-prof}
out=gpurun_out/prof_$tag
mkdir -p $out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu --no-4k --no-geometry"
MATCH="python scripts/time_match.py --leg big --steps 2"
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err || { echo "bench failed"; tail -5 $out/bench_n1.err; exit 1; }
$BENCH > $out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/launches_bench_steps2.csv $BENCH > $out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$BENCH > $out/plain_bench2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_harris|k_nms|k_describe|k_median_compact' -s 39 -c 13 \
    -o $out/extract_full -f $BENCH > $out/ncu_extract.log 2>&1
echo "extract full rc=$?"
$MATCH > $out/plain_match.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_match_tc|k_match_recheck|k_match_rescan|k_match_sort' -s 12 -c 4 \
    -o $out/match_full -f $MATCH > $out/ncu_match.log 2>&1
echo "match full rc=$?"
ls -la $out
