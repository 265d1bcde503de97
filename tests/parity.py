"""Parity rules shared by the tests (stated here as code).

Keypoints   X, Y identical element for element.  The reference orders a level
            by an unstable argsort of the float32 response, so inside a run of
            bit-equal responses any permutation is the same answer; the seeded
            synthetic images have no such ties and the tests assert equality.
Descriptors same keypoint => |delta| <= DESC_ATOL for every element.  The
            reference itself is only reproducible to 1 ulp here: its
            np.linalg.norm goes through BLAS sdot, whose summation order depends
            on buffer alignment.  The CUDA path evaluates arctan2 in double and
            rounds once, numpy's float32 arctan2 is a few-ulp SVML routine, so a
            sample lying within a few ulp of a histogram edge can land in the
            neighbouring bin: at most DESC_FLIP_FRAC of the keypoints (and never
            fewer than 1 allowed) may exceed the tolerance.
Matches     for identical descriptor inputs the (row, index, confidence)
            triples are bit-identical; order is compared after sorting runs of
            equal confidence by row (the reference's tie order is arbitrary).
            When descriptors differ by the tolerance above (end-to-end runs),
            rows whose ratio is within RATIO_EDGE of the threshold or whose two
            nearest distances are within RATIO_EDGE relative are exempt.
"""
import numpy as np

DESC_ATOL = 3.0e-7        # ~2.5 ulp at 1.0
DESC_FLIP_FRAC = 0.004
RATIO_EDGE = 1.0e-5


def assert_keypoints_equal(X, Y, Xr, Yr):
    X, Y, Xr, Yr = map(np.asarray, (X, Y, Xr, Yr))
    assert X.shape == Xr.shape and Y.shape == Yr.shape, (X.shape, Xr.shape)
    assert X.dtype == np.int64 or X.size == 0
    bad = np.nonzero((X != Xr) | (Y != Yr))[0]
    assert bad.size == 0, f"{bad.size} keypoints differ, first at {bad[:5]}"


def assert_descriptors_close(D, Dr, atol=DESC_ATOL, flip_frac=DESC_FLIP_FRAC):
    D, Dr = np.asarray(D), np.asarray(Dr)
    assert D.shape == Dr.shape, (D.shape, Dr.shape)
    if D.size == 0:
        return 0
    assert D.dtype == np.float32
    err = np.abs(D.astype(np.float64) - Dr.astype(np.float64)).max(axis=1)
    over = int((err > atol).sum())
    allowed = max(1, int(np.ceil(flip_frac * len(err))))
    assert over <= allowed, f"{over} of {len(err)} descriptors differ by more than {atol} (allowed {allowed}); worst {err.max()}"
    return over


def canonical_matches(matches, conf):
    matches, conf = np.asarray(matches), np.asarray(conf)
    if matches.size == 0:
        return np.zeros((0, 2), np.int64), np.zeros((0,), np.float32)
    order = np.lexsort((matches[:, 0], conf))
    return matches[order].astype(np.int64), conf[order].astype(np.float32)


def assert_matches_identical(matches, conf, matches_ref, conf_ref):
    m, c = canonical_matches(matches, conf)
    mr, cr = canonical_matches(matches_ref, conf_ref)
    assert m.shape == mr.shape, (m.shape, mr.shape)
    assert np.array_equal(m, mr), "match indices differ"
    assert np.array_equal(c.view(np.uint32), cr.view(np.uint32)), "confidences differ bitwise"
    # the emitted order must already be ascending in confidence
    conf = np.asarray(conf)
    assert np.all(np.diff(conf) >= 0)
