#!/bin/bash
# stream_resident against the serial step at N GPUs
n=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for f in "" "--serial-step"; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29540 + n)) bench.py --gpus $n --steps 100 --warmup 5 --no-4k --no-geometry --no-all-pairs $f > gpurun_out/r02i_n${n}$f.json 2> gpurun_out/r02i_n${n}$f.err; echo "rc $?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02i_n${n}$f.json').read().strip().splitlines()[-1])
print('n $n flag [$f] value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'sustained',round(d['sustained']['ms_per_step'],3),'e2e',round(d['e2e']['value']),'exchange_ms',d['exchange_ms'])
PY
done
