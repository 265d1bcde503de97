"""Extraction-only timing (development aid): B x 1080p through sfm_extract_batch, per-kernel CUDA events from the library."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfmfromscratch_b200 import _native as N
from sfmfromscratch_b200.extractor import extract_batch_device, make_params
from sfmfromscratch_b200.synth import frame_sequence

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
H, W = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1080, 1920)
N.load_library()
N.get_ctx(0)
imgs = torch.from_numpy(frame_sequence(H, W, 0, 2 * B, 2 * B, threads=8)).cuda()
p, keep = make_params({}, pyramid=True)
halves = [imgs[:B], imgs[B:]]
for i in range(4):
    out = extract_batch_device(halves[i & 1], p, want_aux=False, check=False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(steps):
    out = extract_batch_device(halves[i & 1], p, want_aux=False, check=False)
b.record()
torch.cuda.synchronize()
total = a.elapsed_time(b) / steps
N.profile_enable(True, 0)
for i in range(steps):
    out = extract_batch_device(halves[i & 1], p, want_aux=False, check=False)
torch.cuda.synchronize()
st = N.profile_collect(0)
N.profile_enable(False, 0)
print(f"--- {B} x {H}x{W}: {total:.4f} ms per call ({B * H * W / total / 1e6:.2f} Gpixel/s), variant {os.environ.get('SFM_HARRIS_VARIANT', '0')}")
for k, v in sorted(st.items()):
    print(f"    {k:18s} {v[1] / steps:.4f} ms  ({v[0] / steps:.0f} launches)")
print("    counts", out['count'].cpu().tolist()[:4], "checksum", int(out['x'].long().sum().item()), float(out['desc'].double().sum().item()))
