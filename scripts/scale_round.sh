#!/bin/bash
# Multi-GPU record under `gpurun --gpus 8`: configs[4] in full and the bench at N = 8, 4, 2.
out=gpurun_out/scale_${1:-r01}
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29511 scripts/run_config5.py > $out/config5_full_n8.json 2> $out/config5.err
echo "config5 rc=$?"; tail -1 $out/config5_full_n8.json
for n in 8 4 2; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29520 + n)) bench.py --gpus $n --steps 100 --warmup 5 --no-cpu --no-4k --no-geometry > $out/bench_n$n.json 2> $out/bench_n$n.err
  echo "bench n=$n rc=$?"; python -c "
import json,sys
d=json.loads(open('$out/bench_n$n.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['match']['value'])"
done
