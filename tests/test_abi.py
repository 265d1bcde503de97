"""CPU checks of the boundary: the library loads, exports every symbol the
header declares, sizes workspaces, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from sfmfromscratch_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(N.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return N.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "sfmb200.h")).read()
    declared = sorted(set(re.findall(r"SFM_EXPORT[^;(]*?\b(sfm_\w+)\s*\(", hdr)))
    assert declared == sorted(N.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sfm_version() == 1


def test_default_params_and_sizes(lib):
    p = N.SfmExtractParams()
    lib.sfm_extract_default_params(C.byref(p))
    assert (p.num_interest_points, p.ksize, p.gaussian_size, p.feature_width, p.pyramid_level) == (2500, 7, 7, 16, 4)
    assert (p.sigma, p.alpha, p.pyramid_scale_factor) == (5.0, 0.05, 2.0)
    assert lib.sfm_extract_max_keypoints(C.byref(p)) == 2500
    one = lib.sfm_extract_workspace_bytes(1, 1080, 1920, C.byref(p))
    many = lib.sfm_extract_workspace_bytes(8, 1080, 1920, C.byref(p))
    assert 0 < one < many <= 8 * one
    p.gaussian_size = 8            # even window: unsupported
    assert lib.sfm_extract_workspace_bytes(1, 64, 64, C.byref(p)) == 0
    assert lib.sfm_match_workspace_bytes(2, 1000, 1) > 0
    assert lib.sfm_match_workspace_bytes(0, 1000, 1) == 0


def test_struct_layout_matches_header():
    # int32 x3, double x2, int32 x2, double, int32 x3, pointer -- natural C alignment
    assert C.sizeof(N.SfmExtractParams) == 72
    assert N.SfmExtractParams.sigma.offset == 16 and N.SfmExtractParams.gauss_weights.offset == 64


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.sfm_ctx_create(0, C.byref(h)) == N.SFM_ERR_CUDA
    assert b"no CPU fallback" in lib.sfm_last_error(None)
    from sfmfromscratch_b200 import NNRatioFeatureMatcher, ScaleRotInvSIFT
    with pytest.raises(Exception):
        ScaleRotInvSIFT(np.zeros((32, 32), np.float32), {})
    with pytest.raises(Exception):
        NNRatioFeatureMatcher().match_features_ratio_test(np.zeros((4, 128), np.float32), np.zeros((4, 128), np.float32))


def test_host_side_error_behaviour():
    from sfmfromscratch_b200 import NaiveSIFT, NNRatioFeatureMatcher
    with pytest.raises(RuntimeError, match="Keypoints not detected"):
        NaiveSIFT(np.zeros((32, 32), np.float32), {}).extract_descriptors()          # NaiveSIFT.py:49-50
    with pytest.raises(IndexError):                                                 # NNRatioFeatureMatcher.py:44
        NNRatioFeatureMatcher().match_features_ratio_test(np.zeros((4, 128), np.float32), np.zeros((1, 128), np.float32))
    m, c = NNRatioFeatureMatcher().match_features_ratio_test(np.zeros((0, 128), np.float32), np.zeros((5, 128), np.float32))
    assert m.shape == (0,) and c.shape == (0,)


def test_dropin_package_names():
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "from FeatureExtractor import FeatureExtractor, NaiveSIFT, ScaleRotInvSIFT;"
            "from FeatureExtractor.SIFT.ScaleRotInvSIFT import ScaleRotInvSIFT as S2;"
            "from FeatureMatcher import NNRatioFeatureMatcher;"
            "import sfmfromscratch_b200 as s; assert S2 is s.ScaleRotInvSIFT and NNRatioFeatureMatcher is s.NNRatioFeatureMatcher;"
            "print('ok')") % (ROOT, os.path.join(ROOT, "sfmfromscratch_b200", "dropin"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr
