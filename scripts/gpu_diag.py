"""Staged GPU diagnostics (development aid): python scripts/gpu_diag.py <stage>."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from oracle import oracle as O  # noqa: E402
from sfmfromscratch_b200 import _native as N  # noqa: E402
from sfmfromscratch_b200 import extractor as X  # noqa: E402
from sfmfromscratch_b200 import matcher as M  # noqa: E402
from sfmfromscratch_b200.synth import second_view, synth_descriptors, synth_image  # noqa: E402


def stage_harris():
    for (h, w, params) in [(96, 128, {}), (101, 135, {}), (480, 640, {}), (70, 50, {'gaussian_size': 5, 'sigma': 2.0}),
                           (65, 200, {'gaussian_size': 3, 'sigma': 1.0}), (130, 131, {'gaussian_size': 9, 'sigma': 3.0})]:
        img = synth_image(h, w, 1)
        R = X.harris_response(img, params)
        Ro = O.harris_response(img, params.get('gaussian_size', 7), params.get('sigma', 5))
        eq = np.array_equal(R.view(np.uint32), Ro.view(np.uint32))
        print(f"harris {h}x{w} {params}: bit-exact={eq} maxdiff={np.abs(R - Ro).max():.3e} nbad={(R != Ro).sum()}")
        if not eq:
            bad = np.argwhere(R != Ro)
            print("  first bad", bad[:5], R[tuple(bad[0])], Ro[tuple(bad[0])])


def cmp_extract(img, params, tag):
    t = time.time()
    g = X.ScaleRotInvSIFT(img, params)
    torch.cuda.synchronize()
    tg = time.time() - t
    o = O.ScaleRotInvSIFT(img, params)
    gx, gy = g.detect_keypoints()
    ox, oy = o.detect_keypoints()
    print(f"{tag}: gpu n={len(gx)} oracle n={len(ox)} gpu {tg*1e3:.1f} ms")
    for l in range(o._pyramid_level):
        print(f"   level {l}: gpu {(g.levels == l).sum()} oracle {(o.levels == l).sum()}")
    if len(gx) == len(ox):
        same = (gx == ox) & (gy == oy)
        print(f"   keypoints equal: {same.all()} ({(~same).sum()} differ)")
        if len(gx):
            d = np.abs(g.extract_descriptors().astype(np.float64) - o.extract_descriptors()).max(axis=1)
            print(f"   desc maxdiff {d.max():.3e}; >3e-7: {(d > 3e-7).sum()}; >1e-5: {(d > 1e-5).sum()}; exact rows {(d == 0).sum()}")
            if (~same).any():
                i = np.nonzero(~same)[0][0]
                print("   first diff", i, gx[i], gy[i], ox[i], oy[i], g.confidences[i], o.confidences[i])
    else:
        sg = set(zip(g.levels.tolist(), g.level_x.tolist(), g.level_y.tolist()))
        so = set(zip(o.levels.tolist(), o.level_x.tolist(), o.level_y.tolist()))
        print("   only gpu", sorted(sg - so)[:8], "only oracle", sorted(so - sg)[:8])
    return g, o


def stage_extract():
    cmp_extract(synth_image(96, 128, 0), {'num_interest_points': 600}, "96x128")
    cmp_extract(synth_image(240, 320, 0), {}, "240x320")
    cmp_extract(synth_image(101, 135, 11), {'num_interest_points': 400}, "odd 101x135")
    cmp_extract(synth_image(120, 160, 7), {'num_interest_points': 900, 'ksize': 3, 'sigma': 6, 'feature_width': 18,
                                          'pyramid_level': 3, 'pyramid_scale_factor': 1.1}, "mainpy 120x160")
    cmp_extract(synth_image(480, 640, 0), {}, "480x640")
    img = synth_image(96, 128, 3)
    g = X.NaiveSIFT(img, {'num_interest_points': 300}); gx, gy = g.detect_keypoints()
    o = O.NaiveSIFT(img, {'num_interest_points': 300}); ox, oy = o.detect_keypoints()
    print("naive: n", len(gx), len(ox), "eq", np.array_equal(gx, ox) and np.array_equal(gy, oy),
          "desc maxdiff", np.abs(g.extract_descriptors() - o.extract_descriptors()).max() if len(gx) == len(ox) else None)
    z = np.zeros((64, 80), np.float32)
    g = X.NaiveSIFT(z, {'num_interest_points': 100}); gx, gy = g.detect_keypoints()
    print("flat image: n", len(gx), "(plateau / overflow retry path)")


def cmp_match(f1, f2, thr, mode, tag):
    mo, co = O.NNRatioFeatureMatcher(thr).match_features_ratio_test(f1, f2)
    d1 = torch.from_numpy(f1).cuda(); d2 = torch.from_numpy(f2).cuda()
    torch.cuda.synchronize()
    t = time.time()
    m, c, cnt = M.match_device(d1, d2, thr, mode)
    torch.cuda.synchronize()
    dt = time.time() - t
    k = int(cnt.cpu()[0])
    m = m[:k].cpu().numpy().astype(np.int64); c = c[:k].cpu().numpy()
    ok = (len(m) == len(mo)) and np.array_equal(m, mo) and np.array_equal(c, co)
    print(f"{tag} mode={mode}: gpu {k} matches, oracle {len(mo)}, identical={ok}, {dt*1e3:.2f} ms")
    if not ok and len(mo):
        sm = {tuple(r) for r in m.tolist()}; so = {tuple(r) for r in mo.tolist()}
        print("   only gpu", sorted(sm - so)[:6], "only oracle", sorted(so - sm)[:6])
        if len(m) == len(mo):
            bad = np.nonzero((m != mo).any(axis=1) | (c != co))[0]
            print("   first bad rows", bad[:5], m[bad[:3]], mo[bad[:3]], c[bad[:3]], co[bad[:3]])
    return ok


def stage_match_exact():
    g = np.load(os.path.join(ROOT, "tests/golden/matcher_220x260.npz"))
    cmp_match(g["f1"], g["f2"], 0.8, N.SFM_MATCH_EXACT, "golden 220x260")
    cmp_match(synth_descriptors(1000, 0), synth_descriptors(1100, 1), 0.8, N.SFM_MATCH_EXACT, "1000x1100")
    cmp_match(synth_descriptors(3, 0), synth_descriptors(2, 1), 0.9, N.SFM_MATCH_EXACT, "3x2")


def stage_match_tc():
    g = np.load(os.path.join(ROOT, "tests/golden/matcher_220x260.npz"))
    cmp_match(g["f1"], g["f2"], 0.8, N.SFM_MATCH_AUTO, "golden 220x260")
    cmp_match(synth_descriptors(1000, 0), synth_descriptors(1100, 1), 0.8, N.SFM_MATCH_AUTO, "1000x1100")
    cmp_match(synth_descriptors(3, 0), synth_descriptors(2, 1), 0.9, N.SFM_MATCH_AUTO, "3x2")
    f1, f2 = synth_descriptors(2048, 2), synth_descriptors(2300, 3)
    cmp_match(f1, f2, 0.8, N.SFM_MATCH_AUTO, "2048x2300")
    # stats through the batch API
    desc = torch.zeros((2, 2300, 128), device='cuda'); desc[0, :2048] = torch.from_numpy(f1).cuda(); desc[1] = torch.from_numpy(f2).cuda()
    counts = torch.tensor([2048, 2300], dtype=torch.int32, device='cuda')
    pairs = torch.tensor([[0, 1], [1, 0]], dtype=torch.int32, device='cuda')
    m, c, cnt, st = M.match_batch_device(desc, counts, pairs, 0.8, want_stats=True)
    print("batch counts", cnt.cpu().tolist(), "stats [flagged rows, groups visited]", st.cpu().tolist())


def stage_match_big():
    n = 8192
    f1, f2 = synth_descriptors(n, 10), synth_descriptors(n, 11)
    d1 = torch.from_numpy(f1).cuda(); d2 = torch.from_numpy(f2).cuda()
    res = {}
    for mode in (N.SFM_MATCH_AUTO, N.SFM_MATCH_EXACT):
        for it in range(3):
            torch.cuda.synchronize(); t = time.time()
            m, c, cnt = M.match_device(d1, d2, 0.8, mode)
            torch.cuda.synchronize(); dt = time.time() - t
        k = int(cnt.cpu()[0])
        res[mode] = (m[:k].cpu().numpy(), c[:k].cpu().numpy())
        print(f"8192x8192 mode {mode}: {k} matches, {dt*1e3:.2f} ms -> {n*n/dt/1e9:.2f} G pairs/s, {256*n*n/dt/1e12:.1f} TFLOP/s")
    print("auto == exact:", np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]))
    desc = torch.stack([d1, d2]); counts = torch.tensor([n, n], dtype=torch.int32, device='cuda')
    pairs = torch.tensor([[0, 1]], dtype=torch.int32, device='cuda')
    m, c, cnt, st = M.match_batch_device(desc, counts, pairs, 0.8, want_stats=True)
    print("stats [flagged rows, groups visited]", st.cpu().tolist())


def stage_time_extract():
    imgs = np.stack([synth_image(1080, 1920, s) for s in range(4)])
    p, keep = X.make_params({}, pyramid=True)
    dev = torch.from_numpy(imgs).cuda()
    for B in (1, 4):
        for it in range(3):
            torch.cuda.synchronize(); t = time.time()
            out = X.extract_batch_device(dev[:B], p, want_aux=False)
            torch.cuda.synchronize(); dt = time.time() - t
        print(f"extract B={B} 1080p: {dt*1e3:.2f} ms  -> {B*1080*1920/dt/1e6:.0f} Mpix/s; counts {out['count'].cpu().tolist()}")


if __name__ == "__main__":
    st = sys.argv[1]
    t0 = time.time()
    globals()["stage_" + st]()
    torch.cuda.synchronize()
    print(f"[stage {st} done in {time.time()-t0:.1f}s]")
