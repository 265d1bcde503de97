"""The parity rules themselves (tests/parity.py), on the CPU: the descriptor check has no quota -- a descriptor over
the tolerance must be licensed by a window sample within 4 ulp of a histogram edge (or a 36-bin top-2 tie) -- and the
license prover neither licenses generic keypoints nor misses the constructed edge case."""
import numpy as np
import pytest

from parity import (DESC_ATOL, DescriptorExplainer, assert_descriptors_close, assert_matches_equivalent)


def test_no_quota_without_a_license():
    D = np.random.default_rng(0).random((500, 128)).astype(np.float32)
    E = D.copy()
    assert assert_descriptors_close(E, D) == 0
    E[123, 7] += 5 * DESC_ATOL                            # ONE descriptor of 500 off: the old 0.4 % quota let it pass
    with pytest.raises(AssertionError):
        assert_descriptors_close(E, D)
    with pytest.raises(AssertionError):
        assert_descriptors_close(E, D, explain=lambda i: None)
    assert assert_descriptors_close(E, D, explain=lambda i: "licensed" if i == 123 else None) == 1


def test_explainer_licenses_bin_edge_samples_only():
    from oracle import oracle as O
    from sfmfromscratch_b200.synth import synth_image
    # generic image: (almost) no keypoint has a sample within 4 ulp of an edge
    img = synth_image(240, 320, 3)
    o = O.ScaleRotInvSIFT(img, {})
    ex = DescriptorExplainer(img, {}, o.levels, o.level_x, o.level_y)
    licensed = [i for i in range(len(o.levels)) if ex(i) is not None]
    assert len(licensed) <= max(2, len(o.levels) // 100)
    # axis-aligned step edge: gradients at exact multiples of pi/4 sit ON the 8-bin edges
    step = np.zeros((64, 80), np.float32)
    step[20:44, 30:60] = 0.5
    n = O.NaiveSIFT(step, {'num_interest_points': 4000})
    X, Y = n.detect_keypoints()
    exn = DescriptorExplainer(step, {'num_interest_points': 4000}, None, X, Y, pyramid=False)
    near = [i for i in range(len(X)) if 22 <= Y[i] <= 42 and 32 <= X[i] <= 58]
    assert near and all(exn(i) is not None for i in near[:20])


def test_matches_equivalent_rule():
    m = np.array([[0, 5], [3, 9], [7, 1]])
    c = np.array([0.1, 0.5, 0.79999995], np.float32)
    assert assert_matches_equivalent(m, c, m, c, 0.8) == 0
    assert assert_matches_equivalent(m, c, m[:2], c[:2], 0.8) == 1          # the dropped row sits on the threshold
    with pytest.raises(AssertionError):
        assert_matches_equivalent(m, c, m[[0, 2]], c[[0, 2]], 0.8)          # a row far from the threshold may not vanish
    bad = m.copy()
    bad[1, 1] = 8
    with pytest.raises(AssertionError):
        assert_matches_equivalent(bad, c, m, c, 0.8)
