"""The oracle restatement against the committed outputs of the real reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from parity import assert_descriptors_close, assert_keypoints_equal, assert_matches_identical


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_harris_response_bit_exact(golden_dir):
    g = load(golden_dir, "two_view_96x128.npz")
    R = O.harris_response(g["img1"])
    assert np.array_equal(R.view(np.uint32), g["R1"].view(np.uint32))
    assert np.array_equal(O.gaussian_kernel(7, 5).astype(np.float32), g["gauss"])


@pytest.mark.parametrize("name,levels,factor", [("two_view_96x128.npz", 4, 2), ("srs_odd_101x135.npz", 4, 2),
                                                 ("srs_mainpy_120x160.npz", 3, 1.1)])
def test_pyramid_bit_exact(golden_dir, name, levels, factor):
    g = load(golden_dir, name)
    img = g["img1"] if "img1" in g else g["img"]
    pyr = O.build_pyramid(img, levels, factor)
    for l in range(1, levels):
        ref = g[f"pyr{l}"]
        assert pyr[l].shape == ref.shape
        assert np.array_equal(pyr[l].view(np.uint32), ref.view(np.uint32)), f"level {l}"


@pytest.mark.parametrize("name", ["two_view_96x128.npz", "two_view_240x320.npz"])
def test_two_view(golden_dir, name):
    g = load(golden_dir, name)
    params = {'num_interest_points': 600} if "96x128" in name else {}
    feats = []
    for i in (1, 2):
        e = O.ScaleRotInvSIFT(g[f"img{i}"], params)
        X, Y = e.detect_keypoints()
        assert_keypoints_equal(X, Y, g[f"X{i}"], g[f"Y{i}"])
        assert_descriptors_close(e.extract_descriptors(), g[f"D{i}"], atol=1.3e-7)
        feats.append(e.extract_descriptors())
    # matcher on the reference's own descriptors: bit-identical
    m, c = O.NNRatioFeatureMatcher(0.8).match_features_ratio_test(g["D1"], g["D2"])
    assert_matches_identical(m, c, g["matches"], g["conf"])


def test_mainpy_params(golden_dir):
    g = load(golden_dir, "srs_mainpy_120x160.npz")
    e = O.ScaleRotInvSIFT(g["img"], json.loads(str(g["params"])))
    X, Y = e.detect_keypoints()
    assert_keypoints_equal(X, Y, g["X"], g["Y"])
    assert_descriptors_close(e.extract_descriptors(), g["D"], atol=1.3e-7)


def test_odd_size(golden_dir):
    g = load(golden_dir, "srs_odd_101x135.npz")
    e = O.ScaleRotInvSIFT(g["img"], {'num_interest_points': 400})
    X, Y = e.detect_keypoints()
    assert_keypoints_equal(X, Y, g["X"], g["Y"])
    assert_descriptors_close(e.extract_descriptors(), g["D"], atol=1.3e-7)


def test_naive_sift(golden_dir):
    g = load(golden_dir, "naive_96x128.npz")
    e = O.NaiveSIFT(g["img"], {'num_interest_points': 300})
    with pytest.raises(RuntimeError):
        e.extract_descriptors()
    X, Y = e.detect_keypoints()
    assert_keypoints_equal(X, Y, g["X"], g["Y"])
    assert np.array_equal(e.confidences, g["conf"])
    assert_descriptors_close(e.extract_descriptors(), g["D"], atol=1.3e-7)


def test_matcher_edge_cases(golden_dir):
    g = load(golden_dir, "matcher_220x260.npz")
    for thr, mk, ck in ((0.8, "matches", "conf"), (0.95, "matches95", "conf95")):
        m, c = O.NNRatioFeatureMatcher(thr).match_features_ratio_test(g["f1"], g["f2"])
        assert_matches_identical(m, c, g[mk], g[ck])
    assert m.dtype == np.int64 and c.dtype == np.float32
    # duplicate train rows -> ratio 1 -> no match; exact hit -> confidence 0, first in order
    assert 3 in set(g["matches"][:, 0]) and g["conf"][0] == 0.0
    with pytest.raises(IndexError):
        O.NNRatioFeatureMatcher().match_features_ratio_test(g["f1"], g["f2"][:1])
    e1, e2 = O.NNRatioFeatureMatcher(0.0).match_features_ratio_test(g["f1"][10:12], g["f2"])
    assert e1.shape == (0,) and e2.shape == (0,)


def test_numpy_sum_order():
    """orc_dist restates np.sum(axis=2) over 128 float32 (8 accumulators + tree)."""
    rng = np.random.default_rng(5)
    a = rng.random((40, 128), dtype=np.float32)
    b = rng.random((50, 128), dtype=np.float32)
    ref = np.sqrt(np.sum((a[:, None] - b[None, :]) ** 2, axis=2))
    assert np.array_equal(O.dist_matrix(a, b).view(np.uint32), ref.view(np.uint32))
