"""Matcher-only timing (development aid): the bench's "match" leg (all pairs of 12 sets of 8192
descriptors) and a 31-pair batch of 1700-row sets, per-kernel CUDA events from the library.
Also checks one pair of each against the exact mode."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfmfromscratch_b200 import _native as N
from sfmfromscratch_b200 import pipeline as PL
from sfmfromscratch_b200.matcher import match_batch_device
from sfmfromscratch_b200.synth import synth_descriptor_base, synth_descriptors


def leg(n_sets, n, pairs, steps=20, real=None):
    dev = torch.device("cuda:0")
    if real is None:
        base = synth_descriptor_base(n)
        sets = np.stack([synth_descriptors(n, i, base=base) for i in range(n_sets)])
        d_sets = torch.from_numpy(sets).to(dev)
        d_cnt = torch.full((n_sets,), n, dtype=torch.int32, device=dev)
    else:
        d_sets, d_cnt = real
    d_pairs = torch.from_numpy(pairs).to(dev)
    for _ in range(3):
        match_batch_device(d_sets, d_cnt, d_pairs, 0.8, cap=n)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        match_batch_device(d_sets, d_cnt, d_pairs, 0.8, cap=n)
    b.record()
    torch.cuda.synchronize()
    total = a.elapsed_time(b) / steps
    N.profile_enable(True, 0)
    for _ in range(steps):
        match_batch_device(d_sets, d_cnt, d_pairs, 0.8, cap=n)
    torch.cuda.synchronize()
    st = N.profile_collect(0)
    N.profile_enable(False, 0)
    mm, mc, mcnt, mst = match_batch_device(d_sets, d_cnt, d_pairs, 0.8, cap=n, want_stats=True)
    np_ = len(pairs)
    print(f"--- {np_} pairs of {n} x {n}: {total:.4f} ms per step; flagged {mst[:, 0].sum().item() / (np_ * n):.5f}, "
          f"groups per row {mst[:, 1].sum().item() / (np_ * n):.3f}, matches per pair {mcnt.float().mean().item():.1f}")
    for k, v in sorted(st.items()):
        print(f"    {k:18s} {v[1] / steps:.4f} ms")
    flop = 256.0 * np_ * n * n
    print(f"    whole path {flop / (total * 1e-3) / 1e12:.1f} TFLOP/s; k_match_tc {flop / (st['k_match_tc'][1] / steps * 1e-3) / 1e12:.1f} TFLOP/s")
    # exact mode must give the same matches
    sub = d_pairs[:2].contiguous()
    r1 = match_batch_device(d_sets, d_cnt, sub, 0.8, cap=n)
    r2 = match_batch_device(d_sets, d_cnt, sub, 0.8, cap=n, mode=N.SFM_MATCH_EXACT)
    same = torch.equal(r1[2], r2[2])
    for q in range(sub.shape[0]):
        k = int(r1[2][q])
        same = same and torch.equal(r1[0][q, :k], r2[0][q, :k]) and torch.equal(r1[1][q, :k], r2[1][q, :k])
    print("    auto == exact on 2 pairs:", same)
    assert same
    per = mst.cpu().numpy()
    print("    per pair flagged rows (first 8):", per[:8, 0].tolist(), " groups:", per[:8, 1].tolist(), " counts:", d_cnt[:8].tolist())


def real_sets(n_img=32, distinct=8):
    """The bench step's descriptors: ScaleRotInvSIFT defaults on synthetic 1080p images."""
    from sfmfromscratch_b200.extractor import extract_batch_device, make_params
    from sfmfromscratch_b200.synth import synth_image
    imgs = np.stack([synth_image(1080, 1920, s) for s in range(distinct)])
    imgs = torch.from_numpy(np.concatenate([imgs] * (n_img // distinct))).cuda()
    params, keep = make_params({}, pyramid=True)
    out = extract_batch_device(imgs, params, want_aux=False)
    return out['desc'].contiguous(), out['count'].contiguous()


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--leg", default="all", choices=["all", "big", "real"])
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    if a.leg in ("all", "real"):
        d, c = real_sets()
        leg(32, d.shape[1], np.stack([np.arange(31), np.arange(1, 32)], 1).astype(np.int32), steps=a.steps, real=(d, c))
    if a.leg in ("all", "big"):
        leg(12, 8192, PL.all_pairs(12), steps=a.steps)
    if a.leg == "all":
        leg(32, 1700, np.stack([np.arange(31), np.arange(1, 32)], 1).astype(np.int32), steps=a.steps)
