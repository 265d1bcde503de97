"""configs[4] data on ONE GPU: every pair matched with and without the ratio prune / rejection (or, with
--exact-all, by the tensor-core path and by the exact float32 scan); per pair the match COUNT, every matched
(row, index) and every confidence (bitwise) are compared on the device -- both paths emit in the canonical
(confidence, row) order, so the comparison is element-wise.  Pairs that differ are re-run in SFM_MATCH_EXACT
mode to say which setting is right.  python scripts/check_config5.py [--images 512] [--ranks 8] [--exact-all]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sfmfromscratch_b200 import _native as N  # noqa: E402
from sfmfromscratch_b200 import pipeline as PL  # noqa: E402
from sfmfromscratch_b200.matcher import match_batch_device, match_workspace  # noqa: E402
from scripts.run_config5 import synth_block  # noqa: E402


def differing_pairs(r0, r1):
    """Pairs whose results differ in count, in any matched (row, index) or in any confidence bit (device-side)."""
    m0, c0, n0 = r0[0], r0[1], r0[2]
    m1, c1, n1 = r1[0], r1[1], r1[2]
    live = torch.arange(m0.shape[1], device=m0.device)[None, :] < torch.minimum(n0, n1)[:, None]
    idx_bad = ((m0 != m1).any(dim=2) & live).any(dim=1)
    conf_bad = ((c0.view(torch.int32) != c1.view(torch.int32)) & live).any(dim=1)
    return torch.nonzero((n0 != n1) | idx_bad | conf_bad).flatten().tolist()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=512)
    ap.add_argument("--ranks", type=int, default=8, help="the multi-GPU run's world size (its per-rank seeds are reproduced)")
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--chunk", type=int, default=512)
    ap.add_argument("--exact-all", action="store_true", help="compare the tensor-core path with the exact scan on every pair")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    per = (a.images + a.ranks - 1) // a.ranks
    desc = torch.zeros((a.ranks * per, a.n, 128), device=dev)
    counts = torch.zeros((a.ranks * per,), dtype=torch.int32, device=dev)
    for r in range(a.ranks):
        s0, s1 = PL.shard_images(a.images, r, a.ranks)
        desc[r * per: r * per + (s1 - s0)] = synth_block(s1 - s0, a.n, 77 + r, dev)
        counts[r * per: r * per + (s1 - s0)] = a.n
    pairs = torch.from_numpy(PL.all_pairs(a.images)).to(dev)
    ws = match_workspace(desc.shape[0], a.n, a.chunk, dev)
    if a.exact_all:
        tot = [0, 0]
        nbad = 0
        for c0 in range(0, len(pairs), a.chunk):
            pc = pairs[c0:c0 + a.chunk]
            r0 = match_batch_device(desc, counts, pc, 0.8, cap=a.n, ws=ws, prepared=c0 > 0)
            r0 = [x.clone() for x in r0]
            r1 = match_batch_device(desc, counts, pc, 0.8, cap=a.n, ws=ws, prepared=True, mode=N.SFM_MATCH_EXACT)
            tot[0] += int(r0[2].sum()); tot[1] += int(r1[2].sum())
            for q in differing_pairs(r0, r1):
                nbad += 1
                k0, k1 = int(r0[2][q]), int(r1[2][q])
                s0 = {(int(x), int(y)): float(z) for (x, y), z in zip(r0[0][q, :k0].cpu().numpy(), r0[1][q, :k0].cpu().numpy())}
                s1 = {(int(x), int(y)): float(z) for (x, y), z in zip(r1[0][q, :k1].cpu().numpy(), r1[1][q, :k1].cpu().numpy())}
                print("pair", pc[q].tolist(), "auto", k0, "exact", k1, "only auto", [(k, s0[k]) for k in set(s0) - set(s1)],
                      "only exact", [(k, s1[k]) for k in set(s1) - set(s0)], flush=True)
            if (c0 // a.chunk) % 32 == 0:
                print("chunk", c0, "totals", tot, "differing pairs", nbad, flush=True)
        print("matches auto", tot[0], "exact", tot[1], "pairs whose counts, indices or confidences differ", nbad)
        return
    tot = [0, 0]
    bad = []
    for c0 in range(0, len(pairs), a.chunk):
        pc = pairs[c0:c0 + a.chunk]
        res = []
        for k, knob in enumerate((0, N.SFM_MATCH_NO_PRUNE)):
            m, c, cnt = match_batch_device(desc, counts, pc, 0.8, mode=N.SFM_MATCH_AUTO | knob, cap=a.n, ws=ws,
                                           prepared=(c0 > 0 or k > 0))
            res.append((m.clone(), c.clone(), cnt.clone()))
            tot[k] += int(cnt.sum())
        for q in differing_pairs(res[0], res[1]):
            bad.append((c0 + q, int(res[0][2][q]), int(res[1][2][q])))
    print("matches with prune", tot[0], "without", tot[1], "pairs whose counts, indices or confidences differ", len(bad))
    for (pi, n_p, n_np) in bad[:40]:
        pc = pairs[pi:pi + 1]
        m, c, cnt = match_batch_device(desc, counts, pc, 0.8, cap=a.n, mode=N.SFM_MATCH_EXACT)
        print("pair", pairs[pi].tolist(), "prune", n_p, "no-prune", n_np, "exact", int(cnt[0]))


if __name__ == "__main__":
    main()
