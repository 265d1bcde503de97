"""Host-side mirror of the two-view RANSAC that consumes the matcher's output
(SURVEY.md section 8f row 2): `_convert_matches_to_coords` (Runner.py:423-434),
`CameraPose.find_inliers` (SFM.py:126-160), `CameraPose.ransac_camera_motion`
(SFM.py:38-124) and `CameraPose.calculate_num_ransac_iterations` (SFM.py:185-187).

All hypothesis work runs on the GPU (csrc/ransac.cu).  The 8-subsets come from the
library's host-side replica of numpy's legacy generator, so a call returns the
inliers the reference returns for the same arrays.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from functools import lru_cache
from typing import Optional, Tuple

import numpy as np
import torch

from . import _native as N

RANSAC_SEED = 5     # SFM.py:45,133


@lru_cache(maxsize=64)
def _samples_host(n: int, iterations: int, seed: int) -> torch.Tensor:
    out = torch.empty((iterations, 8), dtype=torch.int32)
    if torch.cuda.is_available():
        out = out.pin_memory()
    rc = N.load_library().sfm_ransac_sample_indices(seed, n, iterations, out.data_ptr())
    if rc != N.SFM_OK:
        raise ValueError(f"sfm_ransac_sample_indices({seed}, {n}, {iterations}) failed with {rc}")
    return out


def sample_indices(n: int, iterations: int, seed: int = RANSAC_SEED) -> np.ndarray:
    """The `iterations` draws of np.random.choice(n, 8, replace=False) after np.random.seed(seed)."""
    return _samples_host(int(n), int(iterations), int(seed)).numpy().astype(np.int64)


def calculate_num_ransac_iterations(prob_success: float, sample_size: int, ind_prob_correct: float) -> int:
    """SFM.py:185-187 (host arithmetic, as in the reference)."""
    return int(np.log(1 - prob_success) / np.log(1 - (ind_prob_correct ** sample_size)))


def convert_matches_to_coords(sift_matches, X1, Y1, X2, Y2, num_matches: int = 2500):
    """Runner.py:423-434 on host arrays (an index gather; the device-resident version is
    `matches_to_coords_device`)."""
    sift_matches = np.asarray(sift_matches)
    if sift_matches.shape[0] == 0:
        return np.array([]), np.array([])
    m = sift_matches[:num_matches]
    return (np.column_stack((X1[m[:, 0]], Y1[m[:, 0]])), np.column_stack((X2[m[:, 1]], Y2[m[:, 1]])))


def matches_to_coords_device(matches: torch.Tensor, count: torch.Tensor, x1, y1, x2, y2, num_matches: int = 2500):
    """sfm_matches_to_coords: the matcher's device outputs (matches [*,2] int32, count [1] int32) and
    the extractor's int32 coordinate tensors -> (p1 [num_matches,2] f64, p2, n [1] int32), on the GPU."""
    L = N.load_library()
    dev = matches.device
    ctx = N.get_ctx(dev.index)
    num = int(min(num_matches, matches.shape[0]))
    with torch.cuda.device(dev):
        p1 = torch.zeros((num, 2), dtype=torch.float64, device=dev)
        p2 = torch.zeros((num, 2), dtype=torch.float64, device=dev)
        n = torch.zeros((1,), dtype=torch.int32, device=dev)
        N.check(L.sfm_matches_to_coords(ctx, torch.cuda.current_stream().cuda_stream, matches.data_ptr(), count.data_ptr(),
                                        x1.data_ptr(), y1.data_ptr(), x2.data_ptr(), y2.data_ptr(), num,
                                        p1.data_ptr(), p2.data_ptr(), n.data_ptr()), ctx)
    return p1, p2, n


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def ransac_device(p1: torch.Tensor, p2: torch.Tensor, iterations: int, threshold: float = 1.0,
                  pose: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]] = None,
                  samples: Optional[torch.Tensor] = None, keep_workspace: bool = False):
    """sfm_find_inliers / sfm_ransac_camera_motion on float64 CUDA tensors [n,2].

    Returns (inlier_idx [n] int32, result [4] int32, best [57] f64[, workspace]) as device tensors:
    result = (winner or -1, inlier count, valid-candidate bits, number of degenerate samples -- see
    include/sfmb200.h: 0 means the outcome is the reference's to the index); best = winner's F (9) followed,
    in pose mode, by its four candidates (R row-major, T)."""
    if not (p1.is_cuda and p2.is_cuda) or p1.dtype != torch.float64 or p2.dtype != torch.float64 \
            or p1.dim() != 2 or p1.shape[1] != 2 or p1.shape != p2.shape:
        raise ValueError("p1, p2 must be float64 CUDA tensors of identical shape [n, 2]")
    n = p1.shape[0]
    if n < 8:
        raise ValueError("RANSAC needs at least 8 correspondences")
    p1, p2 = p1.contiguous(), p2.contiguous()
    dev = p1.device
    L = N.load_library()
    ctx = N.get_ctx(dev.index)
    with torch.cuda.device(dev):
        if samples is None:
            samples = _samples_host(n, int(iterations), RANSAC_SEED).to(dev, non_blocking=True)
        nbytes = L.sfm_ransac_workspace_bytes(int(iterations))
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        idx = torch.empty((n,), dtype=torch.int32, device=dev)
        res = torch.zeros((4,), dtype=torch.int32, device=dev)
        best = torch.zeros((57,), dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        if pose is None:
            N.check(L.sfm_find_inliers(ctx, st, p1.data_ptr(), p2.data_ptr(), n, samples.data_ptr(), int(iterations),
                                       float(threshold), ws.data_ptr(), nbytes, idx.data_ptr(), res.data_ptr(),
                                       best.data_ptr()), ctx)
        else:
            K1, K2, Rb, Tb = (np.ascontiguousarray(a, dtype=np.float64) for a in pose)
            if K1.shape != (3, 3) or K2.shape != (3, 3) or Rb.shape != (3, 3) or Tb.size != 3:
                raise ValueError("K1, K2, R_base must be 3x3 and T_base a 3-vector")
            N.check(L.sfm_ransac_camera_motion(ctx, st, p1.data_ptr(), p2.data_ptr(), n, _dptr(K1), _dptr(K2), _dptr(Rb),
                                               _dptr(Tb), samples.data_ptr(), int(iterations), float(threshold),
                                               ws.data_ptr(), nbytes, idx.data_ptr(), res.data_ptr(), best.data_ptr()), ctx)
        ws.record_stream(torch.cuda.current_stream())
        samples.record_stream(torch.cuda.current_stream())
    return (idx, res, best, ws) if keep_workspace else (idx, res, best)


def debug_views(ws: torch.Tensor, iterations: int):
    """(F [it,3,3], counts [it], valid [it], candidates [it,4,12]) of the last call on `ws`, as numpy."""
    L = N.load_library()
    ptrs = [C.c_void_p() for _ in range(4)]
    rc = L.sfm_ransac_debug_views(ws.data_ptr(), iterations, *[C.byref(p) for p in ptrs])
    if rc != N.SFM_OK:
        raise ValueError("sfm_ransac_debug_views failed")
    torch.cuda.synchronize(ws.device)
    host = ws.cpu().numpy()
    base = ws.data_ptr()
    off = [p.value - base for p in ptrs]
    F = host[off[0]:off[0] + iterations * 72].view(np.float64).reshape(iterations, 3, 3)
    counts = host[off[1]:off[1] + iterations * 4].view(np.int32)
    valid = host[off[2]:off[2] + iterations * 4].view(np.uint32)
    cand = host[off[3]:off[3] + iterations * 384].view(np.float64).reshape(iterations, 4, 12)
    return F.copy(), counts.copy(), valid.copy(), cand.copy()


def _upload(p: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(p, dtype=np.float64)).to('cuda')


class CameraPose:
    """The RANSAC half of SFM.py's CameraPose (:22-160,185-187), same signatures and return values.

    When the winning hypothesis of `ransac_camera_motion` has more than one pose candidate in front
    of both cameras, the reference returns whichever comes first in an order fixed by the signs
    LAPACK gives the (numerically null) third singular pair of E -- rounding noise; this class
    returns the first valid one in the library's canonical order (include/sfmb200.h)."""

    def __init__(self, pts1, pts2, K1, K2):
        self.pts1, self.pts2, self.K1, self.K2 = pts1, pts2, K1, K2

    calculate_num_ransac_iterations = staticmethod(calculate_num_ransac_iterations)

    @staticmethod
    def find_inliers(p1, p2, threshold=1.0, max_iterations=1000):
        if len(p1) < 8:
            return None, None, None, None                      # SFM.py:130-131 (a 4-tuple in the reference too)
        p1, p2 = np.asarray(p1), np.asarray(p2)
        if max_iterations < 1:
            return np.array([]), np.array([])
        idx, res, _ = ransac_device(_upload(p1), _upload(p2), max_iterations, threshold)
        res = res.cpu().numpy()
        if res[0] < 0:
            return np.array([]), np.array([])                  # np.array([]) of the untouched best lists
        keep = idx[:int(res[1])].cpu().numpy()
        return p1[keep], p2[keep]

    def ransac_camera_motion(self, R_base, T_base, threshold=1.0, max_iterations=1000):
        if len(self.pts1) < 8:
            return None, None, None, None
        p1, p2 = np.asarray(self.pts1), np.asarray(self.pts2)
        if max_iterations < 1:
            return None, None, np.array([]), np.array([])
        idx, res, best = ransac_device(_upload(p1), _upload(p2), max_iterations, threshold,
                                       pose=(self.K1, self.K2, R_base, np.asarray(T_base, dtype=np.float64).reshape(3)))
        res = res.cpu().numpy()
        if res[0] < 0:
            return None, None, np.array([]), np.array([])
        best = best.cpu().numpy()
        c = int(res[2] & -res[2]).bit_length() - 1             # first valid candidate, canonical order
        cand = best[9 + 12 * c: 9 + 12 * (c + 1)]
        keep = idx[:int(res[1])].cpu().numpy()
        return cand[:9].reshape(3, 3).copy(), cand[9:].copy(), p1[keep], p2[keep]


def find_inliers_many(pairs, threshold: float = 1.0, max_iterations: int = 1000, threads: int = 8):
    """`CameraPose.find_inliers` over a list of (p1, p2) pairs -- what the reference's 8-thread pool
    does pair by pair (Runner.py:186-191,351).  The 8-subsets of every pair are drawn concurrently on
    `threads` host threads (the draw is the sequential MT19937 replay; ctypes releases the GIL), the
    hypothesis kernels of all pairs queue on the current stream, and results are read back once.
    Returns a list with the reference's per-pair return values."""
    from concurrent.futures import ThreadPoolExecutor
    pairs = [(np.asarray(a), np.asarray(b)) for a, b in pairs]
    todo = [k for k, (a, _) in enumerate(pairs) if len(a) >= 8 and max_iterations >= 1]
    out = [(None, None, None, None) if len(a) < 8 else (np.array([]), np.array([])) for a, _ in pairs]
    if not todo:
        return out
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        samples = list(ex.map(lambda k: _samples_host(len(pairs[k][0]), int(max_iterations), RANSAC_SEED), todo))
    res = []
    for k, s in zip(todo, samples):
        a, b = pairs[k]
        res.append(ransac_device(_upload(a), _upload(b), max_iterations, threshold, samples=s.to('cuda', non_blocking=True)))
    torch.cuda.synchronize()
    for k, (idx, r, _) in zip(todo, res):
        r = r.cpu().numpy()
        if r[0] >= 0:
            keep = idx[:int(r[1])].cpu().numpy()
            out[k] = (pairs[k][0][keep], pairs[k][1][keep])
    return out
