"""Nearest-point association after the two-view stage:
Runner.py:241-247 and Runner.py:361-385.  Index work: bit-exact."""
import os

import numpy as np
import pytest

from oracle import geometry as G


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "geometry_assoc.npz"))


@pytest.mark.parametrize("k", range(3))
def test_oracle_association_matches_golden(gold, k):
    rows, near = G.associate_prev_frame(gold[f"assoc{k}_prev"], gold[f"assoc{k}_query"], float(gold[f"assoc{k}_thr"]))
    assert np.array_equal(rows, gold[f"assoc{k}_rows"]) and np.array_equal(near, gold[f"assoc{k}_near"])


@pytest.mark.parametrize("k", range(2))
def test_oracle_dedup_matches_golden(gold, k):
    ia, store = G.dedup_points(gold[f"dedup{k}_a"])
    ib, store = G.dedup_points(gold[f"dedup{k}_b"], store)
    assert np.array_equal(ia, gold[f"dedup{k}_idx_a"]) and np.array_equal(ib, gold[f"dedup{k}_idx_b"])
    assert np.array_equal(store, gold[f"dedup{k}_store"])


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(3))
def test_gpu_association_matches_golden(gold, k):
    from sfmfromscratch_b200 import association as A
    rows, near = A.associate_prev_frame(gold[f"assoc{k}_prev"], gold[f"assoc{k}_query"], float(gold[f"assoc{k}_thr"]))
    assert rows.dtype == np.int64 and np.array_equal(rows, gold[f"assoc{k}_rows"]) and np.array_equal(near, gold[f"assoc{k}_near"])


@pytest.mark.gpu
@pytest.mark.parametrize("m,q,thr,seed,integer", [(2500, 2500, 5.0, 0, True), (700, 33, 1.5, 1, False), (31, 1000, 50.0, 2, True),
                                                   (1, 1, 5.0, 3, True), (1000, 1000, 0.0, 4, True)])
def test_gpu_association_matches_oracle(m, q, thr, seed, integer):
    from sfmfromscratch_b200 import association as A
    rng = np.random.default_rng(seed)
    if integer:                       # keypoint coordinates: many exact distance ties, first index must win
        prev = rng.integers(0, 200, (m, 2)).astype(np.int64)
        qry = rng.integers(0, 200, (q, 2)).astype(np.int64)
    else:
        prev, qry = rng.uniform(0, 100, (m, 2)), rng.uniform(0, 100, (q, 2))
    rows, near = A.associate_prev_frame(prev, qry, thr)
    ro, no = G.associate_prev_frame(prev, qry, thr)
    assert np.array_equal(rows, ro) and np.array_equal(near, no)
    e1, e2 = A.associate_prev_frame(prev, np.zeros((0, 2)), thr)
    assert e1.shape == (0,) and e2.shape == (0,)


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(2))
def test_gpu_dedup_matches_golden(gold, k):
    from sfmfromscratch_b200 import association as A
    s = A.PointStore()
    rng = np.random.default_rng(0)
    a, b = gold[f"dedup{k}_a"], gold[f"dedup{k}_b"]
    s.add_points(a, rng.integers(0, 900, (len(a), 2)), 0)
    first = len(s.point_indices)
    s.add_points(b, rng.integers(0, 900, (len(b), 2)), 1)
    assert np.array_equal(s.point_indices[:first], gold[f"dedup{k}_idx_a"])
    assert np.array_equal(s.point_indices[first:], gold[f"dedup{k}_idx_b"])
    assert np.array_equal(np.array(s.global_points_3D), gold[f"dedup{k}_store"])
    assert s.frame_indices == [0] * len(a) + [1] * len(b) and len(s.global_points_2D) == len(a) + len(b)


@pytest.mark.gpu
def test_gpu_dedup_chains_and_overflow():
    """Near-duplicate chains (a, a+0.6e-6, a+1.2e-6: the third is new again because only the first
    is stored), a batch that is one point repeated (every pair is near: the pair list overflows its
    default capacity and the call is repeated), and a large store."""
    import torch
    from sfmfromscratch_b200 import association as A
    rng = np.random.default_rng(5)
    base = rng.normal(size=(300, 3))
    pts = base.copy()
    pts[100] = pts[7]
    pts[101] = pts[7] + np.array([6e-7, 0, 0])
    pts[102] = pts[7] + np.array([1.2e-6, 0, 0])
    pts[103] = pts[102]
    idx, store = G.dedup_points(pts)
    gi, gn, k = A.dedup_device(torch.from_numpy(pts).cuda(), None)
    assert k == len(store) and np.array_equal(gi.cpu().numpy(), idx)
    same = np.tile(base[:1], (200, 1))
    idx, store = G.dedup_points(same)
    gi, gn, k = A.dedup_device(torch.from_numpy(same).cuda(), None, pair_cap=10)
    assert k == 1 and np.array_equal(gi.cpu().numpy(), idx)
    big = rng.normal(size=(20000, 3))
    batch = np.concatenate([big[rng.integers(0, 20000, 500)], rng.normal(size=(500, 3))])
    rng.shuffle(batch)
    idx, store = G.dedup_points(batch, big)
    gi, gn, k = A.dedup_device(torch.from_numpy(batch).cuda(), torch.from_numpy(big).cuda())
    assert np.array_equal(gi.cpu().numpy(), idx) and 20000 + k == len(store)
