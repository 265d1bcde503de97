"""Parity of the CUDA matcher (through the C ABI) with the oracle / reference
goldens.  Needs a B200: pytest -m gpu."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from parity import assert_matches_identical  # noqa: E402

AUTO, EXACT = 0, 1


def _mods():
    from oracle import oracle as O
    import sfmfromscratch_b200 as S
    return O, S


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("mode", [AUTO, EXACT])
def test_matcher_golden(golden_dir, mode):
    """Reference outputs incl. duplicate train rows, an exact hit (d0 == 0) and a zero descriptor."""
    _, S = _mods()
    g = load(golden_dir, "matcher_220x260.npz")
    for thr, mk, ck in ((0.8, "matches", "conf"), (0.95, "matches95", "conf95")):
        m, c = S.NNRatioFeatureMatcher(thr, mode=mode).match_features_ratio_test(g["f1"], g["f2"])
        assert m.dtype == np.int64 and c.dtype == np.float32 and m.shape[1] == 2
        assert_matches_identical(m, c, g[mk], g[ck])


@pytest.mark.parametrize("mode", [AUTO, EXACT])
def test_two_view_golden_descriptors(golden_dir, mode):
    """The reference's own descriptors of the two-view pair -> the reference's matches, bit for bit."""
    _, S = _mods()
    for name in ("two_view_96x128.npz", "two_view_240x320.npz"):
        g = load(golden_dir, name)
        m, c = S.NNRatioFeatureMatcher(0.8, mode=mode).match_features_ratio_test(g["D1"], g["D2"])
        assert_matches_identical(m, c, g["matches"], g["conf"])


@pytest.mark.parametrize("n1,n2,thr", [(1, 2, 0.9), (3, 2, 0.9), (5, 300, 0.8), (300, 5, 0.8), (129, 257, 0.7),
                                        (1000, 1100, 0.8), (2048, 2300, 0.85), (4097, 1025, 0.8), (700, 9000, 0.8)])
@pytest.mark.parametrize("mode", [AUTO, EXACT])
def test_matcher_vs_oracle(n1, n2, thr, mode):
    O, S = _mods()
    from sfmfromscratch_b200.synth import synth_descriptors
    base_n = max(n1, n2)
    from sfmfromscratch_b200.synth import synth_descriptor_base
    base = synth_descriptor_base(base_n)
    f1 = synth_descriptors(n1, n1, base=base)
    f2 = synth_descriptors(n2, n2 + 1, base=base)
    m, c = S.NNRatioFeatureMatcher(thr, mode=mode).match_features_ratio_test(f1, f2)
    mo, co = O.NNRatioFeatureMatcher(thr).match_features_ratio_test(f1, f2)
    assert_matches_identical(m, c, mo, co)


@pytest.mark.parametrize("mode", [AUTO, EXACT])
def test_sparse_real_descriptors(mode):
    """Descriptors as the extractor makes them (coarse levels have 1-10 non-zeros, exact duplicates occur)."""
    O, S = _mods()
    from sfmfromscratch_b200.synth import second_view, synth_image
    a = synth_image(480, 640, 0)
    b = second_view(a, 1)
    da = S.ScaleRotInvSIFT(a, {}).extract_descriptors()
    db = S.ScaleRotInvSIFT(b, {}).extract_descriptors()
    for thr in (0.8, 0.85):
        m, c = S.NNRatioFeatureMatcher(thr, mode=mode).match_features_ratio_test(da, db)
        mo, co = O.NNRatioFeatureMatcher(thr).match_features_ratio_test(da, db)
        assert len(mo) > 100
        assert_matches_identical(m, c, mo, co)


def test_error_behaviour():
    _, S = _mods()
    z = np.zeros((4, 128), np.float32)
    with pytest.raises(IndexError):
        S.NNRatioFeatureMatcher().match_features_ratio_test(z, z[:1])
    m, c = S.NNRatioFeatureMatcher().match_features_ratio_test(z, z)     # all distances 0: d1 > 0 fails
    assert m.shape == (0,) and c.shape == (0,)
    with pytest.raises(ValueError):
        S.NNRatioFeatureMatcher().match_features_ratio_test(np.zeros((4, 64), np.float32), np.zeros((4, 64), np.float32))


def test_batch_api_equals_single():
    import torch
    _, S = _mods()
    from sfmfromscratch_b200.synth import synth_descriptor_base, synth_descriptors
    base = synth_descriptor_base(900)
    sets = [synth_descriptors(n, i, base=base) for i, n in enumerate((900, 640, 777, 2))]
    nmax = 900
    desc = torch.zeros((4, nmax, 128), dtype=torch.float32, device='cuda')
    for i, s in enumerate(sets):
        desc[i, :len(s)] = torch.from_numpy(s).cuda()
    counts = torch.tensor([len(s) for s in sets], dtype=torch.int32, device='cuda')
    pl = [(0, 1), (1, 0), (0, 2), (2, 1), (1, 3), (2, 2)]
    pairs = torch.tensor(pl, dtype=torch.int32, device='cuda')
    m, c, cnt, st = S.match_batch_device(desc, counts, pairs, 0.8, want_stats=True)
    cnt = cnt.cpu().numpy()
    for k, (i, j) in enumerate(pl):
        ms, cs = S.NNRatioFeatureMatcher(0.8).match_features_ratio_test(sets[i], sets[j])
        assert cnt[k] == len(ms)
        if len(ms):
            assert np.array_equal(m[k, :cnt[k]].cpu().numpy().astype(np.int64), ms)
            assert np.array_equal(c[k, :cnt[k]].cpu().numpy(), cs)
    # a set matched with itself: every row's nearest neighbour is itself at distance 0
    k = pl.index((2, 2))
    mm = m[k, :cnt[k]].cpu().numpy()
    assert np.array_equal(mm[:, 0], mm[:, 1]) and np.all(c[k, :cnt[k]].cpu().numpy() == 0)


def test_full_size_8192_auto_equals_exact_and_oracle_sample():
    """configs[4] pair shape (8192 x 8192): the tensor-core path equals the exact
    float32 scan (two independent CUDA paths), and a 256-row sample equals the oracle."""
    import torch
    O, S = _mods()
    from sfmfromscratch_b200.synth import synth_descriptor_base, synth_descriptors
    base = synth_descriptor_base(8192)
    f1 = synth_descriptors(8192, 0, base=base)
    f2 = synth_descriptors(8192, 1, base=base)
    d1, d2 = torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda()
    out = []
    for mode in (AUTO, EXACT):
        m, c, cnt = S.match_device(d1, d2, 0.8, mode)
        k = int(cnt.cpu()[0])
        out.append((m[:k].cpu().numpy().astype(np.int64), c[:k].cpu().numpy()))
    assert len(out[0][0]) > 1000
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    rows = np.arange(0, 8192, 32)
    mo, co = O.NNRatioFeatureMatcher(0.8).match_features_ratio_test(f1[rows], f2)
    sel = np.isin(out[0][0][:, 0], rows)
    mg = out[0][0][sel].copy()
    mg[:, 0] = np.searchsorted(rows, mg[:, 0])
    assert_matches_identical(mg, out[0][1][sel], mo, co)


def test_4k_pair_20k_descriptors_auto_equals_exact():
    """configs[2] matching half: ~20 k x 20 k descriptors (three column splits per row block).
    Both CUDA paths must agree bit for bit; a 128-row sample is checked against the oracle."""
    import torch
    O, S = _mods()
    from sfmfromscratch_b200.synth import synth_descriptor_base, synth_descriptors
    n1, n2 = 19531, 20007
    base = synth_descriptor_base(n2)
    f1 = synth_descriptors(n1, 7, base=base)
    f2 = synth_descriptors(n2, 8, base=base)
    d1, d2 = torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda()
    res = []
    for mode in (AUTO, EXACT):
        m, c, cnt = S.match_device(d1, d2, 0.8, mode)
        k = int(cnt.cpu()[0])
        res.append((m[:k].cpu().numpy().astype(np.int64), c[:k].cpu().numpy()))
    assert len(res[0][0]) > 3000
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    rows = np.arange(0, n1, 153)
    mo, co = O.NNRatioFeatureMatcher(0.8).match_features_ratio_test(f1[rows], f2)
    sel = np.isin(res[0][0][:, 0], rows)
    mg = res[0][0][sel].copy()
    mg[:, 0] = np.searchsorted(rows, mg[:, 0])
    assert_matches_identical(mg, res[0][1][sel], mo, co)


def test_prepared_workspace_reused_across_pair_chunks():
    """SFM_MATCH_PREPARED: the per-set preparation of the first chunk serves later chunks with
    other pair lists (and other pair counts) in the same workspace; results equal one-shot calls."""
    import torch
    _, S = _mods()
    from sfmfromscratch_b200.matcher import match_workspace
    from sfmfromscratch_b200.synth import synth_descriptor_base, synth_descriptors
    base = synth_descriptor_base(1200)
    sizes = (1200, 1100, 900, 1000, 1200)
    desc = torch.zeros((len(sizes), 1200, 128), dtype=torch.float32, device='cuda')
    for i, n in enumerate(sizes):
        desc[i, :n] = torch.from_numpy(synth_descriptors(n, 40 + i, base=base)).cuda()
    counts = torch.tensor(sizes, dtype=torch.int32, device='cuda')
    chunks = [[(0, 1), (1, 2), (2, 3)], [(3, 4), (4, 0)], [(2, 0), (1, 4), (3, 1), (0, 4)]]
    ws = match_workspace(len(sizes), 1200, 4, 'cuda')
    for k, ch in enumerate(chunks):
        pairs = torch.tensor(ch, dtype=torch.int32, device='cuda')
        got = S.match_batch_device(desc, counts, pairs, 0.8, ws=ws, prepared=k > 0)
        ref = S.match_batch_device(desc, counts, pairs, 0.8)
        assert torch.equal(got[2], ref[2])
        for q in range(len(ch)):
            n = int(ref[2][q])
            assert n > 50
            assert torch.equal(got[0][q, :n], ref[0][q, :n]) and torch.equal(got[1][q, :n], ref[1][q, :n])
    with pytest.raises(ValueError):
        S.match_batch_device(desc, counts, pairs, 0.8, prepared=True)


@pytest.mark.parametrize("thr", [0.0, 0.3, 0.64, 0.8, 0.95, 1.0, 1.25, -0.5])
def test_ratio_prune_is_exact_across_thresholds(thr):
    """The re-check proves rows rejected from the approximate keys alone (ratio prune) -- the emitted
    set must still be the exact scan's for every threshold, including rows whose ratio sits on it.
    Query rows are train rows plus noise of continuously growing strength, so ratios cover (0, 1];
    exact duplicates and zero rows are mixed in (d0 = d1 = 0, d1 > 0 with d0 = 0)."""
    import torch
    _, S = _mods()
    rng = np.random.default_rng(7)
    n1, n2 = 3000, 2777
    f2 = rng.gamma(0.5, size=(n2, 128)).astype(np.float32)
    f2 /= np.linalg.norm(f2, axis=1, keepdims=True)
    f2 = np.sqrt(f2)
    f2[100:110] = f2[100]                         # ten identical train rows
    f2[200] = 0.0
    src = rng.integers(0, n2, size=n1)
    strength = np.linspace(0.0, 1.5, n1, dtype=np.float32)[:, None]
    f1 = f2[src] + strength * rng.gamma(0.5, size=(n1, 128)).astype(np.float32) * 0.2
    f1[5] = f2[100]                               # nearest and second-nearest both at distance 0
    f1[6] = 0.0                                   # exact hit on the zero row, d1 > 0
    d1, d2 = torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda()
    res = []
    for mode in (AUTO, EXACT):
        m, c, cnt = S.match_device(d1, d2, thr, mode)
        k = int(cnt.cpu()[0])
        res.append((m[:k].cpu().numpy(), c[:k].cpu().numpy()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    if thr >= 0.3:
        assert len(res[0][0]) > 100
    if thr < 0.0:
        assert len(res[0][0]) == 0


def test_ragged_random_batches_auto_equals_exact():
    """Edge sizes of the re-check / rescan kernels: tiny and ragged sets (2 .. 700 rows; candidate groups
    and MMA tiles mostly padding), duplicated and zero rows, several thresholds -- the tensor-core path
    must equal the exact scan pair by pair, and the oracle on the smallest pairs."""
    import torch
    O, S = _mods()
    rng = np.random.default_rng(11)
    for trial in range(6):
        n_sets = int(rng.integers(3, 7))
        nmax = int(rng.choice([2, 3, 9, 64, 257, 700]))
        sizes = [int(rng.integers(2, nmax + 1)) for _ in range(n_sets)]
        sizes[0] = nmax
        desc = np.zeros((n_sets, nmax, 128), np.float32)
        pool = np.sqrt(rng.gamma(0.4, size=(nmax, 128)).astype(np.float32))
        for i, n in enumerate(sizes):
            d = pool[rng.permutation(nmax)[:n]] + 0.05 * i * np.sqrt(rng.gamma(0.4, size=(n, 128)).astype(np.float32))
            if n > 4:
                d[1] = d[0]                       # duplicate rows inside a set
                d[2] = 0.0
            desc[i, :n] = d
        pl = [(a, b) for a in range(n_sets) for b in range(n_sets) if a != b][:10] + [(0, 0)]
        dd = torch.from_numpy(desc).cuda()
        cc = torch.tensor(sizes, dtype=torch.int32, device='cuda')
        pp = torch.tensor(pl, dtype=torch.int32, device='cuda')
        thr = float(rng.choice([0.5, 0.8, 0.97, 1.0]))
        ra = S.match_batch_device(dd, cc, pp, thr, mode=AUTO)
        re = S.match_batch_device(dd, cc, pp, thr, mode=EXACT)
        assert torch.equal(ra[2], re[2]), (trial, nmax, sizes, thr, ra[2].tolist(), re[2].tolist())
        for q in range(len(pl)):
            k = int(ra[2][q])
            assert torch.equal(ra[0][q, :k], re[0][q, :k]) and torch.equal(ra[1][q, :k], re[1][q, :k]), (trial, q, pl[q])
        if nmax <= 64 and thr < 1.0:              # ties at thr >= 1 are ordered by numpy's unstable argsort in the oracle
            for q, (a, b) in enumerate(pl[:4]):
                mo, co = O.NNRatioFeatureMatcher(thr).match_features_ratio_test(desc[a, :sizes[a]], desc[b, :sizes[b]])
                k = int(ra[2][q])
                assert k == len(mo)
                if k:
                    assert_matches_identical(ra[0][q, :k].cpu().numpy().astype(np.int64), ra[1][q, :k].cpu().numpy(), mo, co)
