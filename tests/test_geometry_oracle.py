"""Two-view RANSAC, CPU side: the oracle against the reference's
golden vectors and the live reference, the library's host sampler against numpy, and the kernels'
float64 arithmetic (compiled for the host from the very same header) against the oracle."""
import ctypes as C
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from oracle import geometry as G
from sfmfromscratch_b200 import _native as N
from sfmfromscratch_b200.synth import two_view_correspondences

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_FIND, N_POSE = 5, 3


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "geometry_ransac.npz"))


@pytest.mark.parametrize("k", range(N_FIND))
def test_oracle_find_inliers_matches_golden(gold, k):
    d = {}
    a, b = G.find_inliers(gold[f"find{k}_p1"], gold[f"find{k}_p2"], max_iterations=int(gold[f"find{k}_it"]), detail=d)
    assert a.dtype == gold[f"find{k}_in1"].dtype and np.array_equal(a, gold[f"find{k}_in1"])
    assert np.array_equal(b, gold[f"find{k}_in2"])
    assert np.array_equal(d["counts"], gold[f"find{k}_counts"])
    assert np.array_equal(d["samples"], gold[f"find{k}_samples"])


@pytest.mark.parametrize("k", range(N_POSE))
def test_oracle_camera_motion_matches_golden(gold, k):
    K = gold[f"pose{k}_K"]
    R, T, a, b = G.ransac_camera_motion(gold[f"pose{k}_p1"], gold[f"pose{k}_p2"], K, K, np.eye(3), np.zeros(3),
                                        max_iterations=int(gold[f"pose{k}_it"]))
    assert np.array_equal(a, gold[f"pose{k}_in1"]) and np.array_equal(b, gold[f"pose{k}_in2"])
    if gold[f"pose{k}_R"].size:
        assert np.array_equal(R, gold[f"pose{k}_R"]) and np.array_equal(T, gold[f"pose{k}_T"])
    else:
        assert R is None and T is None


def test_oracle_small_inputs_and_conversion():
    p = np.arange(14).reshape(7, 2)
    assert G.find_inliers(p, p) == (None, None, None, None)
    assert G.ransac_camera_motion(p, p, np.eye(3), np.eye(3), np.eye(3), np.zeros(3)) == (None, None, None, None)
    X1, Y1, X2, Y2 = (np.arange(10) * s for s in (1, 2, 3, 4))
    m = np.array([[3, 1], [0, 9], [5, 5]])
    a, b = G.convert_matches_to_coords(m, X1, Y1, X2, Y2, 2)
    assert np.array_equal(a, [[3, 6], [0, 0]]) and np.array_equal(b, [[3, 4], [27, 36]]) and a.dtype == np.int64
    e1, e2 = G.convert_matches_to_coords(np.array([]), X1, Y1, X2, Y2)
    assert e1.shape == (0,) and e2.shape == (0,)
    assert G.num_ransac_iterations(0.98, 8, 0.4) == 5967          # Runner.py:170


def test_oracle_matches_live_reference():
    ref_root = os.environ.get("SFM_REFERENCE", "/root/reference")
    if not os.path.exists(os.path.join(ref_root, "SFM.py")):
        pytest.skip("reference tree not present (GPU box)")
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_SFM", os.path.join(ref_root, "SFM.py"))
    S = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(S)
    for n, seed, outl, it in [(300, 11, 0.35, 250), (20, 12, 0.2, 120)]:
        p1, p2, K = two_view_correspondences(n, seed, outl)
        a, b = S.CameraPose.find_inliers(p1, p2, max_iterations=it)
        a2, b2 = G.find_inliers(p1, p2, max_iterations=it)
        assert np.array_equal(a, a2) and np.array_equal(b, b2)
    p1, p2, K = two_view_correspondences(30, 13, 0.0)
    r = S.CameraPose(p1, p2, K, K).ransac_camera_motion(np.eye(3), np.zeros(3), max_iterations=60)
    r2 = G.ransac_camera_motion(p1, p2, K, K, np.eye(3), np.zeros(3), max_iterations=60)
    assert all(np.array_equal(x, y) for x, y in zip(r, r2))
    assert S.CameraPose.calculate_num_ransac_iterations(0.98, 8, 0.4) == G.num_ransac_iterations(0.98, 8, 0.4)


def test_reference_candidate_order_is_rounding_noise():
    """Why (R, T) parity is stated on the candidate SET: the order in which the reference tries the
    four decompositions of E flips under a 1e-13 relative perturbation of F (the third singular
    pair of E is numerically null, so LAPACK's sign for it is noise)."""
    rng = np.random.default_rng(0)
    K = np.array([[800.0, 0, 480], [0, 800.0, 270], [0, 0, 1]])
    flips = 0
    for _ in range(200):
        F = rng.normal(size=(3, 3))
        U, D, Vt = np.linalg.svd(F)
        D[2] = 0
        F = U @ np.diag(D) @ Vt
        a = G.pose_candidates(F, K, K)
        b = G.pose_candidates(F * (1 + 1e-13 * rng.normal(size=(3, 3))), K, K)
        flips += any(np.abs(x[0] - y[0]).max() > 1e-6 or np.abs(x[1] - y[1]).max() > 1e-6 for x, y in zip(a, b))
        # ... while the set of four is stable
        for Rc, Tc in a:
            assert min(max(np.abs(Rc - y[0]).max(), np.abs(Tc - y[1]).max()) for y in b) < 1e-6
    assert flips > 20


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(N.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return N.load_library()


@pytest.mark.parametrize("n,it", [(8, 40), (9, 100), (600, 200), (2500, 60), (1024, 50), (1025, 50), (65537, 3)])
def test_host_sampler_is_numpys_legacy_stream(lib, n, it):
    out = np.zeros((it, 8), np.int32)
    assert lib.sfm_ransac_sample_indices(5, n, it, out.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(out, G.sample_indices(n, it))
    # and it is the stream the reference consumes: np.random.seed(5) then np.random.choice
    np.random.seed(5)
    assert np.array_equal(out[0], np.random.choice(n, 8, replace=False))
    assert lib.sfm_ransac_sample_indices(5, 7, 1, out.ctypes.data_as(C.c_void_p)) == N.SFM_ERR_BAD_ARG


@pytest.fixture(scope="module")
def host_math():
    """tests/native/ransac_host_check.cu: the kernels' building blocks compiled for the host."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    src = os.path.join(ROOT, "tests", "native", "ransac_host_check.cu")
    out = os.path.join(ROOT, "tests", "native", "_build", "libransac_host_check.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    hdr = os.path.join(ROOT, "sfmfromscratch_b200", "csrc", "ransac_math.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC,-ffp-contract=off",
                        "-shared", "-o", out, src], check=True)
    L = C.CDLL(out)

    def run(p1, p2, idx, thr=1.0, pose=None):
        it = len(idx)
        a, b = np.ascontiguousarray(p1, np.float64), np.ascontiguousarray(p2, np.float64)
        s = np.ascontiguousarray(idx, np.int32)
        F, cnt = np.zeros((it, 9)), np.zeros(it, np.int32)
        cand, valid = np.zeros((it, 48)), np.zeros(it, np.uint32)
        dp = lambda x: x.ctypes.data_as(C.c_void_p)
        extra = [None] * 4
        if pose is not None:
            keep = [np.ascontiguousarray(x, np.float64) for x in pose]
            extra = [dp(x) for x in keep]
        L.ransac_host_eval(dp(a), dp(b), len(a), dp(s), it, C.c_double(thr), 0 if pose is None else 1, *extra,
                           dp(F), dp(cnt), dp(cand), dp(valid))
        return F.reshape(it, 3, 3), cnt, cand.reshape(it, 4, 12), valid
    return run


def f_rel_err(F, Fo):
    """max |F - (+-)Fo| / max |Fo| per hypothesis (the null vector's sign is free)."""
    sgn = np.sign((F * Fo).sum((1, 2)))[:, None, None]
    return np.abs(F * sgn - Fo).max((1, 2)) / np.abs(Fo).max((1, 2))


@pytest.mark.parametrize("n,seed,outl,it", [(600, 0, 0.3, 400), (50, 1, 0.5, 300), (9, 2, 0.0, 50), (2500, 3, 0.4, 100)])
def test_kernel_math_on_host_matches_oracle(host_math, n, seed, outl, it):
    p1, p2, K = two_view_correspondences(n, seed, outl)
    d = {}
    G.find_inliers(p1, p2, max_iterations=it, detail=d)
    F, cnt, _, _ = host_math(p1, p2, d["samples"])
    rel = f_rel_err(F, d["F"])
    assert np.median(rel) < 1e-13 and rel.max() < 1e-9
    assert np.array_equal(cnt, d["counts"])


@pytest.mark.parametrize("n,seed,outl,it", [(40, 5, 0.0, 100), (100, 6, 0.02, 100), (60, 8, 0.3, 80)])
def test_kernel_pose_math_on_host_matches_oracle(host_math, n, seed, outl, it):
    p1, p2, K = two_view_correspondences(n, seed, outl)
    d = {}
    G.ransac_camera_motion(p1, p2, K, K, np.eye(3), np.zeros(3), max_iterations=it, detail=d)
    F, cnt, cand, valid = host_math(p1, p2, d["samples"], pose=(K, K, np.eye(3), np.zeros(3)))
    for i in range(it):
        for c, (Rc, Tc) in enumerate(G.pose_candidates(d["F"][i], K, K)):
            dist = np.abs(cand[i] - np.concatenate([Rc.ravel(), Tc])).max(1)
            k = int(np.argmin(dist))
            assert dist[k] < 1e-9                                   # same candidate set
            assert bool((valid[i] >> k) & 1) == bool(d["valid"][i, c])   # same cheirality verdicts
