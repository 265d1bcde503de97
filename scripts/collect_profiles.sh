#!/bin/bash
# Copy a profile pass (gpurun_out/prof_<tag>, written by scripts/profile_r02.sh) into the tracked profiles/ records.
tag=${1:-r02b}
P=gpurun_out/prof_$tag
rev=$(git rev-parse --short HEAD)
cp $P/bench_n1.json profiles/r02_bench_n1.json
cp $P/launches_bench_steps2.csv profiles/r02_ncu_launches_bench_steps2.csv
python scripts/launch_summary.py $P/launches_bench_steps2.csv "ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 2 --warmup 3 --no-cpu --no-4k --no-geometry --no-all-pairs (round 2, build of commit $rev)" > profiles/r02_ncu_launches_summary.txt
python scripts/ncu_summary.py profiles/r02_ncu_full_summary.csv "ncu --set full --clock-control none, round 2 (build of commit $rev): the 9 launches of one 32 x 1080p extraction call (scripts/prof_extract.py 32 2, second call); matcher kernels of the 66-pair 8192 x 8192 leg; matcher kernels of the bench-shaped 31-pair leg (scripts/time_match.py)" $P/extract_full.ncu-rep $P/match_full.ncu-rep $P/match_real_full.ncu-rep
python - <<'PY'
import csv, json
rd = list(csv.reader(open('profiles/r02_ncu_full_summary.csv')))[1:]
h = rd[0]
hs = nms = 0.0
tc = None
seen = False
for r in rd[2:]:
    d = dict(zip(h, r))
    n = d['Kernel Name']
    by = (float(d['dram__bytes_read.sum']) + float(d['dram__bytes_write.sum'])) * 1e6
    if not seen and 'k_harris' in n: hs += by
    if not seen and 'k_nms' in n: nms += by
    if 'k_match_tc' in n and tc is None: tc = by
    if 'k_describe' in n: seen = True
out = {"k_harris_dram_bytes_per_image": hs / 32, "k_nms_dram_bytes_per_image": nms / 32, "k_match_tc_dram_bytes_per_launch": tc,
       "source": "profiles/r02_ncu_full_summary.csv: dram__bytes_read.sum + dram__bytes_write.sum; k_harris / k_nms: every launch of one 32 x 1080p extraction call divided by 32; k_match_tc: the launch of the 66-pair 8192 x 8192 leg",
       "algorithmic_bytes_per_image_k_harris": (8 * 2753325 + 4 * (2753325 - 2073600)), "algorithmic_bytes_per_image_k_nms": 4 * 2753325}
json.dump(out, open('profiles/r02_ncu_traffic.json', 'w'), indent=1)
print(out)
PY
