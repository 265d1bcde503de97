"""Times the RANSAC stage on the GPU: host sampler, per-kernel CUDA events, whole call."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from sfmfromscratch_b200 import _native as N, geometry as geo  # noqa: E402
from sfmfromscratch_b200.synth import two_view_correspondences  # noqa: E402

it = geo.calculate_num_ransac_iterations(0.98, 8, 0.4)
out = {}
for n, outl, pose in [(2500, 0.45, False), (600, 0.3, False), (2500, 0.0, True), (600, 0.02, True)]:
    p1, p2, K = two_view_correspondences(n, 41, outl)
    geo._samples_host.cache_clear()
    t = time.perf_counter(); geo.sample_indices(n, it); t_s = time.perf_counter() - t
    a, b = torch.from_numpy(p1.astype(np.float64)).cuda(), torch.from_numpy(p2.astype(np.float64)).cuda()
    kw = dict(pose=(K, K, np.eye(3), np.zeros(3))) if pose else {}
    for _ in range(3):
        geo.ransac_device(a, b, it, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        idx, res, best = geo.ransac_device(a, b, it, **kw)
    e1.record(); torch.cuda.synchronize()
    N.profile_enable(True)
    for _ in range(5):
        geo.ransac_device(a, b, it, **kw)
    torch.cuda.synchronize()
    prof = {k: v[1] / v[0] for k, v in N.profile_collect().items()}
    N.profile_enable(False)
    out[f"n{n}_{'pose' if pose else 'find'}"] = dict(sampler_ms=t_s * 1e3, call_ms=e0.elapsed_time(e1) / 20, kernels_ms=prof,
                                                     result=res.cpu().tolist())
print(json.dumps(out, indent=1))
