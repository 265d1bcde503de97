#!/bin/bash
# Round-2 profile under gpurun (one B200).  Outputs under gpurun_out/prof_<tag>/.
#   gpurun --timeout 1500 -- 'bash scripts/profile_r02.sh r02a [bench]'
# 1. (optional, "bench") python bench.py with default flags -> bench_n1.json
# 2. ncu launch list (gpu__time_duration) of a 2-step bench without the CPU / all-pairs legs
# 3. ncu --set full of every kernel (9 launches) of ONE 32 x 1080p extraction call (the second call of scripts/prof_extract.py)
# 4. ncu --set full of the matcher kernels of the 66-pair 8192 x 8192 leg and of the bench-shaped 31-pair leg
tag=${1:-r02}
out=gpurun_out/prof_$tag
mkdir -p $out
if [ "$2" == "bench" ]; then
  python bench.py > $out/bench_n1.json 2> $out/bench_n1.err || { echo "bench failed"; tail -5 $out/bench_n1.err; exit 1; }
fi
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu --no-4k --no-geometry --no-all-pairs"
$BENCH > $out/plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/launches_bench_steps2.csv $BENCH > $out/ncu_launches.log 2>&1
echo "launch list rc=$?"
EXT="python scripts/prof_extract.py 32 2"
$EXT > $out/plain_extract.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_' -s 9 -c 9 -o $out/extract_full -f $EXT > $out/ncu_extract.log 2>&1
echo "extract full rc=$?"
MATCH="python scripts/time_match.py --leg big --steps 2"
$MATCH > $out/plain_match.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_match_tc|k_match_recheck|k_match_rescan|k_match_sort|k_match_prep' -s 15 -c 5 \
    -o $out/match_full -f $MATCH > $out/ncu_match.log 2>&1
echo "match full rc=$?"
MATCHR="python scripts/time_match.py --leg real --steps 2"
$MATCHR > $out/plain_match_real.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_match' -s 24 -c 8 \
    -o $out/match_real_full -f $MATCHR > $out/ncu_match_real.log 2>&1
echo "match real full rc=$?"
ls -la $out
