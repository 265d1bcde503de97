"""Small extraction-only workload for ncu (development aid): B images 1080p, n calls."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sfmfromscratch_b200.extractor import extract_batch_device, make_params
from sfmfromscratch_b200.synth import synth_image
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
imgs = np.stack([synth_image(1080, 1920, s) for s in range(min(B, 4))] * ((B + 3) // 4))[:B]
dev = torch.from_numpy(imgs).cuda()
p, keep = make_params({}, pyramid=True)
for _ in range(n):
    out = extract_batch_device(dev, p, want_aux=False)
torch.cuda.synchronize()
print("counts", out['count'].cpu().tolist()[:4])
