#!/bin/bash
# A/B the k_harris variants (development aid)
for v in "$@"; do
  SFM_HARRIS_VARIANT=$v python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/var$v.json 2> gpurun_out/var$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/var$v.json"))
k=d["kernels"]
print("variant $v: step %.3f ms" % d["ms_per_step"], {a:round(b["ms_per_step"],4) for a,b in k.items() if not a.startswith("k_match")})
PY
done
