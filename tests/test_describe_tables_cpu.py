"""The float32 thresholds the descriptor kernel bins orientations with decide exactly as the reference's float64
tests do (ScaleRotInvSIFT.py:66-87: np.histogram of float32 arctan2 values against np.linspace(-pi, pi, 37), then of
the float64-shifted values against np.linspace(-pi, pi, 9)).  Host-only: the tables come from the library's
sfm_describe_tables, the float64 side is numpy."""
import numpy as np
import pytest

from sfmfromscratch_b200 import _native as N


@pytest.fixture(scope="module")
def tables():
    return N.describe_tables()


def _probe(thr):
    """float32 values around every threshold (8 neighbours each side) and a random spread of the range."""
    rng = np.random.default_rng(7)
    near = []
    for t in np.asarray(thr, np.float32).ravel():
        if not np.isfinite(t):
            continue
        v = np.float32(t)
        lo = hi = v
        near.append(v)
        for _ in range(8):
            lo = np.nextafter(lo, np.float32(-np.inf)); hi = np.nextafter(hi, np.float32(np.inf))
            near += [lo, hi]
    spread = rng.uniform(-np.pi, np.pi, 200_000).astype(np.float32)
    tiny = (rng.standard_normal(2000) * 1e-20).astype(np.float32)
    ends = np.array([np.float32(np.pi), -np.float32(np.pi), np.nextafter(np.float32(np.pi), np.float32(0)), 0.0, -0.0], np.float32)
    return np.concatenate([np.array(near, np.float32), spread, tiny, ends])


def test_36_bin_thresholds_decide_as_float64_edges(tables):
    ef37, _ = tables
    e37 = np.linspace(-np.pi, np.pi, 37)
    o = _probe(ef37)
    v = o.astype(np.float64)
    for i in range(37):
        assert np.array_equal(o >= ef37[i], v >= e37[i]), i
    assert np.array_equal(o <= ef37[37], v <= e37[36])
    # and the bins they give are numpy's
    inside = (v >= e37[0]) & (v <= e37[36])
    want = np.clip(np.searchsorted(e37, v[inside], side="right") - 1, 0, 35)
    got = np.clip((o[inside, None] >= ef37[None, :37]).sum(1) - 1, 0, 35)
    assert np.array_equal(got, want)
    h_np, _ = np.histogram(v, bins=e37)
    assert np.array_equal(np.bincount(got, minlength=36), h_np)


def test_8_bin_thresholds_decide_as_shifted_float64_edges(tables):
    _, slot = tables
    e9 = np.linspace(-np.pi, np.pi, 9)
    e37 = np.linspace(-np.pi, np.pi, 37)
    assert np.all(np.isposinf(slot[:, 9]))
    for b in range(37):
        dom = (e37[b] + e37[b + 1]) / 2.0 if b < 36 else 0.0
        o = _probe(slot[b, :9])
        rel = o.astype(np.float64) - dom
        for k in range(8):
            assert np.array_equal(o >= slot[b, k], rel >= e9[k]), (b, k)
        assert np.array_equal(o <= slot[b, 8], rel <= e9[8]), b
        assert np.all(np.diff(slot[b, :8]) > 0)
        # the histogram the thresholds give is numpy's on the shifted values
        c = (o[:, None] >= slot[b, None, :8]).sum(1)                     # edges at or below: 0 = under the range
        keep = (c >= 1) & ((c < 8) | (o <= slot[b, 8]))
        h_np, _ = np.histogram(rel, bins=e9)
        assert np.array_equal(np.bincount(c[keep] - 1, minlength=8), h_np), b
