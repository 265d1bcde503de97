"""The drop-in proved through the reference's OWN caller.  Runner.FeatureRunner (Runner.py:22-73) -- file in,
PIL resize, gray mix, `feature_extractor_class(image, params)` twice, `NNRatioFeatureMatcher(...)` -- is run from
the staged, unmodified Runner.py (oracle/_ref, see oracle/stage_reference.py) with sfmfromscratch_b200/dropin
in front of it on sys.path, so `from FeatureExtractor import FeatureExtractor` / `from FeatureMatcher import
NNRatioFeatureMatcher` (Runner.py:9-10) bind to the B200 classes.  The result is compared with the golden
fixture the same class produced from the unshadowed reference (tests/golden/make_golden_runner.py).  Also:
the reference calls its extractor from an 8-thread pool (Runner.py:186-191) -- one context shared by 8 threads
must return what serial calls return."""
import contextlib
import io
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from parity import (DescriptorExplainer, assert_descriptors_close, assert_keypoints_equal, assert_matches_equivalent,
                    assert_matches_identical)
from refstub import reference_path

DROPIN = os.path.join(ROOT, "sfmfromscratch_b200", "dropin")
STAGED = os.path.join(ROOT, "oracle", "_ref")


def _staged_or_skip():
    sys.path.insert(0, ROOT)
    from oracle import stage_reference
    d = stage_reference.stage() or stage_reference.staged()
    if not d or not os.path.exists(os.path.join(d, "Runner.py")):
        pytest.skip("no staged reference (oracle/_ref): run python oracle/stage_reference.py in the build container")
    return d


def test_runner_binds_to_dropin_packages():
    """Import resolution only (no GPU): with dropin/ first, the reference's Runner module sees the B200 classes."""
    ref = _staged_or_skip()
    with reference_path(DROPIN, ref):
        import Runner
        assert os.path.abspath(Runner.__file__).startswith(os.path.abspath(ref))
        assert os.path.abspath(sys.modules["FeatureExtractor"].__file__).startswith(DROPIN)
        assert os.path.abspath(sys.modules["FeatureMatcher"].__file__).startswith(DROPIN)
        import sfmfromscratch_b200 as S
        assert Runner.NNRatioFeatureMatcher is S.NNRatioFeatureMatcher
        assert Runner.FeatureExtractor is S.FeatureExtractor


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["mainpy", "defaults"])
def test_feature_runner_with_dropin_equals_reference_golden(case):
    ref = _staged_or_skip()
    g = np.load(os.path.join(GOLDEN, "runner_two_view.npz"))
    params, thr = json.loads(str(g["cases"]))[case]
    p1, p2 = os.path.join(GOLDEN, "runner_im1.png"), os.path.join(GOLDEN, "runner_im2.png")
    with reference_path(DROPIN, ref):
        import Runner
        from FeatureExtractor import ScaleRotInvSIFT
        assert os.path.abspath(sys.modules["FeatureExtractor"].__file__).startswith(DROPIN)
        with contextlib.redirect_stdout(io.StringIO()):
            fr = Runner.FeatureRunner(p1, p2, scale_factor=0.5, feature_extractor_class=ScaleRotInvSIFT,
                                      extractor_params=dict(params), match_threshold=thr)
    # the reference's own ingest ran (Runner.py:33-46): same gray images as in the golden run
    assert np.array_equal(fr._image1_bw, g[f"{case}_bw1"]) and np.array_equal(fr._image2_bw, g[f"{case}_bw2"])
    for i, ex in ((1, fr.extractor1), (2, fr.extractor2)):
        X, Y, D = getattr(fr, f"X{i}"), getattr(fr, f"Y{i}"), getattr(fr, f"descriptors{i}")
        assert_keypoints_equal(X, Y, g[f"{case}_X{i}"], g[f"{case}_Y{i}"])
        assert D.dtype == np.float32 and D.shape == g[f"{case}_D{i}"].shape
        prove = DescriptorExplainer(g[f"{case}_bw{i}"], params, ex.levels, ex.level_x, ex.level_y)
        assert assert_descriptors_close(D, g[f"{case}_D{i}"], prove) == 0
    # matcher through the reference's caller: bit-identical to the oracle on the descriptors it was given ...
    from oracle import oracle as O
    mo, co = O.NNRatioFeatureMatcher(thr).match_features_ratio_test(fr.descriptors1, fr.descriptors2)
    assert fr.matches.dtype == np.int64 and fr.confidences.dtype == np.float32
    assert_matches_identical(fr.matches, fr.confidences, mo, co)
    # ... and equal to the reference's end-to-end result up to rows on the ratio threshold
    edge = assert_matches_equivalent(fr.matches, fr.confidences, g[f"{case}_matches"], g[f"{case}_conf"], thr)
    assert edge <= 1
    assert len(fr.matches) > 20


@pytest.mark.gpu
def test_eight_threads_share_one_context():
    """Runner.py:186-191: FeatureRunner objects are built from an 8-thread pool.  Eight threads extracting and
    matching at once through the one per-device context return exactly what serial calls return."""
    from concurrent.futures import ThreadPoolExecutor
    import sfmfromscratch_b200 as S
    from sfmfromscratch_b200.synth import second_view, synth_image
    imgs = [synth_image(150 + 8 * k, 200 + 4 * k, 50 + k) for k in range(8)]
    views = [second_view(im, 70 + k) for k, im in enumerate(imgs)]
    params = {'num_interest_points': 800}

    def work(k):
        a, b = S.ScaleRotInvSIFT(imgs[k], params), S.ScaleRotInvSIFT(views[k], params)
        m, c = S.NNRatioFeatureMatcher(0.8).match_features_ratio_test(a.extract_descriptors(), b.extract_descriptors())
        return a.detect_keypoints(), a.extract_descriptors(), b.detect_keypoints(), b.extract_descriptors(), m, c
    serial = [work(k) for k in range(8)]
    for _ in range(3):
        with ThreadPoolExecutor(8) as ex:
            par = list(ex.map(work, range(8)))
        for s, p in zip(serial, par):
            assert np.array_equal(s[0][0], p[0][0]) and np.array_equal(s[0][1], p[0][1])
            assert np.array_equal(s[1], p[1]) and np.array_equal(s[3], p[3])
            assert np.array_equal(s[2][0], p[2][0]) and np.array_equal(s[4], p[4]) and np.array_equal(s[5], p[5])
    assert sum(len(s[4]) for s in serial) > 100
