"""ctypes binding of libsfmb200.so (C ABI: include/sfmb200.h).

The library is the only compute path: if it is missing, or no B200 is
visible, every operation raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsfmb200.so")

SFM_OK = 0
SFM_ERR_BAD_ARG = -1
SFM_ERR_CUDA = -2
SFM_ERR_WORKSPACE = -3
SFM_ERR_CAPACITY = -4
SFM_ERR_UNSUPPORTED = -5
SFM_MATCH_AUTO = 0
SFM_MATCH_EXACT = 1
SFM_MATCH_PREPARED = 16
SFM_MATCH_NO_PRUNE = 32
DESC_DIM = 128

# every symbol include/sfmb200.h declares
EXPORTS = [
    "sfm_version", "sfm_ctx_create", "sfm_ctx_destroy", "sfm_last_error", "sfm_ctx_sm_count",
    "sfm_ctx_launch_count", "sfm_ctx_set_option", "sfm_profile_enable", "sfm_profile_collect",
    "sfm_extract_default_params", "sfm_extract_max_keypoints", "sfm_extract_workspace_bytes",
    "sfm_extract_batch", "sfm_extract_status", "sfm_harris_response", "sfm_describe_tables",
    "sfm_ingest_workspace_bytes", "sfm_ingest_rgb8",
    "sfm_match_workspace_bytes", "sfm_match_prepared_bytes", "sfm_match_ratio", "sfm_match_ratio_batch",
    "sfm_matches_to_coords", "sfm_ransac_sample_indices", "sfm_ransac_workspace_bytes", "sfm_find_inliers",
    "sfm_ransac_camera_motion", "sfm_ransac_debug_views",
    "sfm_associate_nearest", "sfm_dedup_workspace_bytes", "sfm_dedup_points",
]


class SfmExtractParams(C.Structure):
    _fields_ = [
        ("num_interest_points", C.c_int32),
        ("ksize", C.c_int32),
        ("gaussian_size", C.c_int32),
        ("sigma", C.c_double),
        ("alpha", C.c_double),
        ("feature_width", C.c_int32),
        ("pyramid_level", C.c_int32),
        ("pyramid_scale_factor", C.c_double),
        ("rotation_invariant", C.c_int32),
        ("split_k_by_level", C.c_int32),
        ("cand_full", C.c_int32),
        ("gauss_weights", C.POINTER(C.c_float)),
    ]


class SfmKernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_int32), ("total_ms", C.c_float)]


_lib = None
_lib_lock = threading.Lock()
_ctxs = {}


def load_library() -> C.CDLL:
    """dlopen libsfmb200.so and declare its prototypes.  Raises if absent."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(sfmfromscratch_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, i32p, fp = C.c_void_p, C.c_void_p, C.c_void_p       # device pointers travel as integers
        PP = C.POINTER(SfmExtractParams)
        L.sfm_version.restype = C.c_int
        L.sfm_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.sfm_ctx_destroy.argtypes = [C.c_void_p]
        L.sfm_ctx_destroy.restype = None
        L.sfm_last_error.argtypes = [C.c_void_p]
        L.sfm_last_error.restype = C.c_char_p
        L.sfm_ctx_sm_count.argtypes = [C.c_void_p]
        L.sfm_ctx_launch_count.argtypes = [C.c_void_p]
        L.sfm_ctx_launch_count.restype = C.c_ulonglong
        L.sfm_ctx_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.sfm_profile_enable.argtypes = [C.c_void_p, C.c_int]
        L.sfm_profile_collect.argtypes = [C.c_void_p, C.POINTER(SfmKernelStat), C.c_int]
        L.sfm_extract_default_params.argtypes = [PP]
        L.sfm_extract_default_params.restype = None
        L.sfm_extract_max_keypoints.argtypes = [PP]
        L.sfm_extract_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int, PP]
        L.sfm_extract_workspace_bytes.restype = C.c_size_t
        L.sfm_extract_batch.argtypes = [vp, vp, fp, C.c_int, C.c_int, C.c_int, PP, vp, C.c_size_t,
                                        i32p, i32p, i32p, i32p, i32p, fp, fp, i32p, C.c_int]
        L.sfm_extract_status.argtypes = [vp, vp, vp]
        L.sfm_harris_response.argtypes = [vp, vp, fp, C.c_int, C.c_int, PP, fp]
        L.sfm_ingest_workspace_bytes.argtypes = [C.c_int] * 5
        L.sfm_ingest_workspace_bytes.restype = C.c_size_t
        L.sfm_ingest_rgb8.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, fp]
        L.sfm_match_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
        L.sfm_match_workspace_bytes.restype = C.c_size_t
        L.sfm_match_prepared_bytes.argtypes = [C.c_int, C.c_int]
        L.sfm_match_prepared_bytes.restype = C.c_size_t
        L.sfm_match_ratio.argtypes = [vp, vp, fp, C.c_int, fp, C.c_int, C.c_int, C.c_float, C.c_int, vp,
                                      C.c_size_t, i32p, fp, i32p, C.c_int]
        L.sfm_match_ratio_batch.argtypes = [vp, vp, fp, i32p, C.c_int, C.c_int, i32p, C.c_int, C.c_float,
                                            C.c_int, vp, C.c_size_t, i32p, fp, i32p, i32p, C.c_int]
        L.sfm_matches_to_coords.argtypes = [vp, vp, i32p, i32p, i32p, i32p, i32p, i32p, C.c_int, vp, vp, i32p]
        L.sfm_ransac_sample_indices.argtypes = [C.c_uint32, C.c_int, C.c_int, vp]
        L.sfm_describe_tables.argtypes = [vp, vp]
        L.sfm_ransac_workspace_bytes.argtypes = [C.c_int]
        L.sfm_ransac_workspace_bytes.restype = C.c_size_t
        L.sfm_find_inliers.argtypes = [vp, vp, vp, vp, C.c_int, i32p, C.c_int, C.c_double, vp, C.c_size_t, i32p, i32p, vp]
        L.sfm_ransac_camera_motion.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, i32p, C.c_int, C.c_double, vp,
                                               C.c_size_t, i32p, i32p, vp]
        L.sfm_ransac_debug_views.argtypes = [vp, C.c_int] + [C.POINTER(C.c_void_p)] * 4
        L.sfm_associate_nearest.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int, C.c_double, i32p, vp, i32p, i32p, i32p]
        L.sfm_dedup_workspace_bytes.argtypes = [C.c_int, C.c_int]
        L.sfm_dedup_workspace_bytes.restype = C.c_size_t
        L.sfm_dedup_points.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int, C.c_double, C.c_int, vp, C.c_size_t, i32p, i32p, i32p]
        _lib = L
        return L


class SfmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsfmb200 error {code}: {msg}")
        self.code = code


def get_ctx(device: int = 0) -> int:
    """One context per (process, device).  Raises RuntimeError without a B200."""
    L = load_library()
    with _lib_lock:
        if device in _ctxs:
            return _ctxs[device]
        h = C.c_void_p()
        rc = L.sfm_ctx_create(device, C.byref(h))
        if rc != SFM_OK:
            msg = L.sfm_last_error(None).decode()
            raise SfmError(rc, msg)
        _ctxs[device] = h.value
        return h.value


def check(rc: int, ctx: int) -> None:
    if rc != SFM_OK:
        raise SfmError(rc, load_library().sfm_last_error(ctx).decode())


SFM_OPT_HARRIS_STREAM_MIN_BANDS = 1


def set_option(option: int, value: int, device: int = 0) -> None:
    """sfm_ctx_set_option on the context of `device` (tuning only: results never depend on it)."""
    check(load_library().sfm_ctx_set_option(get_ctx(device), option, value), get_ctx(device))


def describe_tables():
    """(ef37 [38], slot_thr [37, 10]) float32: the descriptor stage's decision thresholds (host function, no GPU)."""
    import numpy as np
    ef37 = np.zeros(38, np.float32)
    slot = np.zeros((37, 10), np.float32)
    rc = load_library().sfm_describe_tables(ef37.ctypes.data, slot.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"sfm_describe_tables failed with {rc}")
    return ef37, slot


def launch_count(device: int = 0) -> int:
    return int(load_library().sfm_ctx_launch_count(get_ctx(device)))


def profile_enable(on: bool, device: int = 0) -> None:
    check(load_library().sfm_profile_enable(get_ctx(device), 1 if on else 0), get_ctx(device))


def profile_collect(device: int = 0) -> dict:
    """{kernel name: (launches, total_ms)} of the launches recorded since the last call."""
    arr = (SfmKernelStat * 64)()
    n = load_library().sfm_profile_collect(get_ctx(device), arr, 64)
    if n < 0:
        check(n, get_ctx(device))
    return {arr[i].name.decode(): (int(arr[i].launches), float(arr[i].total_ms)) for i in range(n)}
