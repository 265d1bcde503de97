#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bench_shape.py -x -q 2>&1 | tail -4
for f in "" "--serial-step"; do
  timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu --no-4k --no-geometry --no-all-pairs $f > gpurun_out/r02h_bench$f.json 2> gpurun_out/r02h_bench$f.err; echo "rc $?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02h_bench$f.json').read().strip().splitlines()[-1])
print('flag [$f] value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'sustained',round(d['sustained']['ms_per_step'],3),'profiled',round(d['ms_per_step_profiled'],3), 'launches', d['gpu_launches'], 'mpp', d['config']['matches_per_pair'])
PY
done
