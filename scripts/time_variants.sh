#!/bin/bash
# A/B the k_harris variants (development aid)
for v in 0 1 2 3; do
  SFM_HARRIS_VARIANT=$v python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/var$v.json 2> gpurun_out/var$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/var$v.json"))
k=d["kernels"]
print("variant $v: step %.3f ms  harris %.4f  nms %.4f  compact %.4f finish %.4f describe %.4f" % (d["ms_per_step"],k["k_harris"]["ms_per_step"],k["k_nms"]["ms_per_step"],k["k_median_compact"]["ms_per_step"],k["k_median_finish"]["ms_per_step"],k["k_describe"]["ms_per_step"]))
PY
done
