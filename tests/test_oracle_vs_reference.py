"""The oracle against the LIVE reference imported from /root/reference (build
container only; skipped where the tree is absent).  Wider than the committed
fixtures: more seeds, sizes and parameter sets."""
import numpy as np
import pytest

from oracle import oracle as O
from parity import assert_descriptors_close, assert_keypoints_equal, assert_matches_identical
from sfmfromscratch_b200.synth import second_view, synth_descriptors, synth_image

CASES = [
    (72, 96, 2, {}),
    (120, 160, 4, {'num_interest_points': 800}),
    (101, 77, 5, {'num_interest_points': 300, 'ksize': 5, 'gaussian_size': 5, 'sigma': 2.5}),
    (96, 128, 6, {'ksize': 3, 'sigma': 6, 'feature_width': 18, 'pyramid_level': 3, 'pyramid_scale_factor': 1.1}),
    (90, 120, 8, {'pyramid_level': 2, 'pyramid_scale_factor': 1.5, 'feature_width': 12, 'alpha': 0.06}),
]


@pytest.mark.parametrize("h,w,seed,params", CASES)
def test_scale_rot_inv_sift(reference_modules, h, w, seed, params):
    fe, _ = reference_modules
    img = synth_image(h, w, seed)
    r = fe.ScaleRotInvSIFT(img, params)
    o = O.ScaleRotInvSIFT(img, params)
    assert_keypoints_equal(*o.detect_keypoints(), *r.detect_keypoints())
    assert_descriptors_close(o.extract_descriptors(), r.extract_descriptors(), atol=1.3e-7)


def test_naive_sift(reference_modules):
    fe, _ = reference_modules
    img = synth_image(80, 100, 9)
    r, o = fe.NaiveSIFT(img, {'num_interest_points': 200}), O.NaiveSIFT(img, {'num_interest_points': 200})
    assert_keypoints_equal(*o.detect_keypoints(), *r.detect_keypoints())
    assert np.array_equal(o.confidences, r.confidences)
    assert_descriptors_close(o.extract_descriptors(), r.extract_descriptors(), atol=1.3e-7)


def test_harris_plateau_image(reference_modules):
    """Constant regions: R == 0 exactly on a plateau, the median gate's `R == 0` branch."""
    fe, _ = reference_modules
    img = np.zeros((48, 64), np.float32)
    img[10:30, 20:50] = synth_image(20, 30, 1)
    r = fe.NaiveSIFT(img, {'num_interest_points': 5000})
    o = O.NaiveSIFT(img, {'num_interest_points': 5000})
    rx, ry = r.detect_keypoints()
    ox, oy = o.detect_keypoints()
    # with ties (many R == 0) only the multiset of (x, y, conf) is defined
    assert sorted(zip(rx.tolist(), ry.tolist())) == sorted(zip(ox.tolist(), oy.tolist()))


@pytest.mark.parametrize("thr", [0.6, 0.8, 0.85, 1.0])
def test_matcher(reference_modules, thr):
    _, fm = reference_modules
    f1 = synth_descriptors(180, 3)
    f2 = synth_descriptors(200, 4)
    f2[7] = f2[8]
    f1[0] = f2[100]
    m, c = fm.NNRatioFeatureMatcher(thr).match_features_ratio_test(f1, f2)
    mo, co = O.NNRatioFeatureMatcher(thr).match_features_ratio_test(f1, f2)
    if thr < 1.0:
        assert_matches_identical(mo, co, m, c)
    else:
        # ratio == 1 rows (tied nearest neighbours) match; which of the tied columns is reported is arbitrary
        assert np.array_equal(np.sort(mo[:, 0]), np.sort(m[:, 0]))


def test_two_view_end_to_end(reference_modules):
    fe, fm = reference_modules
    a = synth_image(120, 160, 21)
    b = second_view(a, 22)
    ra, rb = fe.ScaleRotInvSIFT(a, {}), fe.ScaleRotInvSIFT(b, {})
    m, c = fm.NNRatioFeatureMatcher(0.8).match_features_ratio_test(ra.extract_descriptors(), rb.extract_descriptors())
    mo, co = O.NNRatioFeatureMatcher(0.8).match_features_ratio_test(ra.extract_descriptors(), rb.extract_descriptors())
    assert_matches_identical(mo, co, m, c)
