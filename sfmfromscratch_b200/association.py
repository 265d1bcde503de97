"""Host-side mirror of the nearest-point association loops after the two-view stage (SURVEY.md
section 8f row 3): Runner.py:241-247 (2-D association against the triangulated set) and
Runner.py:361-385 (`add_points` with `is_new_point` / `find_existing_point`).

The distance scans run on the GPU (csrc/assoc.cu) in float64 with numpy's evaluation order, so the
indices are the reference's.  No CPU fallback.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _native as N


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def associate_device(ref: torch.Tensor, query: torch.Tensor, dist_threshold: float):
    """sfm_associate_nearest on float64 CUDA tensors ref [m,2], query [q,2] ->
    (nearest [q] i32, dist [q] f64, kept [q] i32, count [1] i32), all on the device."""
    for t in (ref, query):
        if not t.is_cuda or t.dtype != torch.float64 or t.dim() != 2 or t.shape[1] != 2:
            raise ValueError("ref and query must be float64 CUDA tensors of shape [n, 2]")
    ref, query = ref.contiguous(), query.contiguous()
    m, q = ref.shape[0], query.shape[0]
    if m < 1 or q < 1:
        raise ValueError("empty point set")
    dev = ref.device
    L = N.load_library()
    ctx = N.get_ctx(dev.index)
    with torch.cuda.device(dev):
        nearest = torch.empty((q,), dtype=torch.int32, device=dev)
        dist = torch.empty((q,), dtype=torch.float64, device=dev)
        flag = torch.empty((q,), dtype=torch.int32, device=dev)
        kept = torch.empty((q,), dtype=torch.int32, device=dev)
        count = torch.zeros((1,), dtype=torch.int32, device=dev)
        N.check(L.sfm_associate_nearest(ctx, _stream(), ref.data_ptr(), m, query.data_ptr(), q, float(dist_threshold),
                                        nearest.data_ptr(), dist.data_ptr(), flag.data_ptr(), kept.data_ptr(),
                                        count.data_ptr()), ctx)
    return nearest, dist, kept, count


def associate_prev_frame(points_2d_prev, prev_frame_2d, dist_threshold: float = 5.0) -> Tuple[np.ndarray, np.ndarray]:
    """Runner.py:241-247: for every row q of `prev_frame_2d`, `mask = np.argmin(dist to points_2d_prev)`;
    the row is kept when that distance is below `dist_threshold`.  Returns (kept q rows, their
    `mask` values) as int64 -- the reference then gathers `p3d[mask]` and `next_frame_2d[q]`."""
    a = np.ascontiguousarray(points_2d_prev, dtype=np.float64)
    b = np.ascontiguousarray(prev_frame_2d, dtype=np.float64)
    if b.shape[0] == 0:
        return np.zeros((0,), np.int64), np.zeros((0,), np.int64)
    if a.shape[0] == 0:
        raise ValueError("attempt to get argmin of an empty sequence")      # np.argmin's error in the reference
    nearest, _, kept, count = associate_device(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), dist_threshold)
    k = int(count.cpu()[0])
    rows = kept[:k].cpu().numpy().astype(np.int64)
    return rows, nearest.cpu().numpy().astype(np.int64)[rows]


def dedup_device(points: torch.Tensor, store: Optional[torch.Tensor], threshold: float = 1e-6,
                 pair_cap: Optional[int] = None):
    """sfm_dedup_points on float64 CUDA tensors points [n,3], store [e,3] (or None) ->
    (index [n] i32, is_new [n] i32, n_new int).  Repeats the call with a larger pair list if the
    batch holds more near-duplicates than `pair_cap`."""
    if not points.is_cuda or points.dtype != torch.float64 or points.dim() != 2 or points.shape[1] != 3:
        raise ValueError("points must be a float64 CUDA tensor of shape [n, 3]")
    points = points.contiguous()
    n = points.shape[0]
    e = 0 if store is None else store.shape[0]
    if e:
        store = store.contiguous()
    dev = points.device
    L = N.load_library()
    ctx = N.get_ctx(dev.index)
    cap = n if pair_cap is None else int(pair_cap)
    with torch.cuda.device(dev):
        index = torch.empty((n,), dtype=torch.int32, device=dev)
        is_new = torch.empty((n,), dtype=torch.int32, device=dev)
        n_new = torch.zeros((1,), dtype=torch.int32, device=dev)
        while True:
            nbytes = L.sfm_dedup_workspace_bytes(n, cap)
            ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            N.check(L.sfm_dedup_points(ctx, _stream(), points.data_ptr(), n, store.data_ptr() if e else None, e,
                                       float(threshold), cap, ws.data_ptr(), nbytes, index.data_ptr(), is_new.data_ptr(),
                                       n_new.data_ptr()), ctx)
            k = int(n_new.cpu()[0])
            if k >= 0:
                return index, is_new, k
            cap = int(index[:1].cpu()[0])                      # the capacity the batch needs


class PointStore:
    """The bookkeeping of SFMRunner.add_points (Runner.py:361-371) with the same list attributes:
    `global_points_3D`, `global_points_2D`, `frame_indices`, `point_indices`."""

    def __init__(self):
        self.global_points_3D: List[np.ndarray] = []
        self.global_points_2D: List[np.ndarray] = []
        self.frame_indices: List[int] = []
        self.point_indices: List[int] = []
        self._dev: Optional[torch.Tensor] = None                # device copy of global_points_3D

    def add_points(self, points_3d, points_2d, frame_idx, threshold: float = 1e-6):
        pts = np.ascontiguousarray(points_3d, dtype=np.float64).reshape(-1, 3)
        n = min(len(pts), len(points_2d))                       # the reference zips the two
        if n == 0:
            return
        dev_pts = torch.from_numpy(pts[:n]).cuda()
        index, is_new, _ = dedup_device(dev_pts, self._dev, threshold)
        idx = index.cpu().numpy()
        new = is_new.cpu().numpy().astype(bool)
        for i in range(n):
            if new[i]:
                self.global_points_3D.append(points_3d[i])
            self.global_points_2D.append(points_2d[i])
            self.frame_indices.append(frame_idx)
            self.point_indices.append(int(idx[i]))
        if new.any():
            fresh = dev_pts[torch.from_numpy(new).cuda()]
            self._dev = fresh if self._dev is None else torch.cat([self._dev, fresh])
