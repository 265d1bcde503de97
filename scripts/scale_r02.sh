#!/bin/bash
# Multi-GPU record under `gpurun --gpus N`: the bench exactly as the driver launches it (default steps / warm-up).
#   gpurun --gpus 4 --timeout 900 -- 'bash scripts/scale_r02.sh 4'
n=${1:-2}
out=gpurun_out/scale_r02
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node $n --master-port $((29520 + n)) bench.py --gpus $n --steps 100 --warmup 5 > $out/bench_n$n.json 2> $out/bench_n$n.err
echo "bench n=$n rc=$?"; tail -3 $out/bench_n$n.err
python - <<PY
import json
d=json.loads(open('$out/bench_n$n.json').read().strip().splitlines()[-1])
ap=d.get('all_pairs') or {}
print('n',d['n_gpus'],'value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'exchange_ms',d.get('exchange_ms'),'sustained',round(d['sustained']['value']))
print('all_pairs s',ap.get('seconds'),'ag_ms',ap.get('all_gather_ms'),'frac_sust',(ap.get('whole_path') or {}).get('frac_of_sustained_peak'),'clocks',ap.get('clocks'))
PY
