"""Host-side mirror of FeatureMatcher/NNRatioFeatureMatcher.py:4-60."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _native as N


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def match_device(f1: torch.Tensor, f2: torch.Tensor, ratio_threshold: float, mode: int = N.SFM_MATCH_AUTO):
    """sfm_match_ratio on float32 CUDA tensors [n1,128], [n2,128] -> device
    tensors (matches [n1,2] int32, conf [n1] f32, count [1] int32)."""
    for f in (f1, f2):
        if not f.is_cuda or f.dtype != torch.float32 or f.dim() != 2 or f.shape[1] != N.DESC_DIM:
            raise ValueError("features must be float32 CUDA tensors of shape [n, 128]")
    f1, f2 = f1.contiguous(), f2.contiguous()
    L = N.load_library()
    ctx = N.get_ctx(f1.device.index)
    n1, n2 = f1.shape[0], f2.shape[0]
    with torch.cuda.device(f1.device):
        nbytes = L.sfm_match_workspace_bytes(2, max(n1, n2), 1)
        ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=f1.device)
        m = torch.empty((n1, 2), dtype=torch.int32, device=f1.device)
        c = torch.empty((n1,), dtype=torch.float32, device=f1.device)
        cnt = torch.zeros((1,), dtype=torch.int32, device=f1.device)
        N.check(L.sfm_match_ratio(ctx, _stream_ptr(), f1.data_ptr(), n1, f2.data_ptr(), n2, N.DESC_DIM,
                                  float(np.float32(ratio_threshold)), mode, ws.data_ptr(), nbytes,
                                  m.data_ptr(), c.data_ptr(), cnt.data_ptr(), n1), ctx)
        # keep the workspace alive until the stream has consumed it
        ws.record_stream(torch.cuda.current_stream())
    return m, c, cnt


def match_batch_device(desc: torch.Tensor, counts: torch.Tensor, pairs: torch.Tensor, ratio_threshold: float,
                       mode: int = N.SFM_MATCH_AUTO, cap: Optional[int] = None, want_stats: bool = False,
                       ws: Optional[torch.Tensor] = None, prepared: bool = False):
    """sfm_match_ratio_batch.  desc [n_sets,nmax,128] f32, counts [n_sets] i32,
    pairs [n_pairs,2] i32, all CUDA.  Returns (matches [P,cap,2], conf [P,cap],
    count [P][, stats [P,2]]).  `ws` is a caller-kept workspace (uint8, at least
    sfm_match_workspace_bytes); with `prepared` it already holds the per-set
    preparation of an earlier call over the same desc / counts (SFM_MATCH_PREPARED)."""
    if not (desc.is_cuda and counts.is_cuda and pairs.is_cuda):
        raise ValueError("desc, counts and pairs must be CUDA tensors")
    if desc.dtype != torch.float32 or desc.dim() != 3 or desc.shape[2] != N.DESC_DIM:
        raise ValueError("desc must be float32 [n_sets, nmax, 128]")
    desc = desc.contiguous()
    counts = counts.to(torch.int32).contiguous()
    pairs = pairs.to(torch.int32).contiguous()
    n_sets, nmax = desc.shape[0], desc.shape[1]
    P = pairs.shape[0]
    cap = nmax if cap is None else cap
    L = N.load_library()
    ctx = N.get_ctx(desc.device.index)
    with torch.cuda.device(desc.device):
        nbytes = L.sfm_match_workspace_bytes(n_sets, nmax, P)
        if ws is None:
            if prepared:
                raise ValueError("prepared=True needs the workspace of the preparing call")
            ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=desc.device)
        elif ws.numel() < nbytes or ws.dtype != torch.uint8 or not ws.is_cuda:
            raise ValueError(f"workspace too small: {ws.numel()} < {nbytes}")
        m = torch.empty((P, cap, 2), dtype=torch.int32, device=desc.device)
        c = torch.empty((P, cap), dtype=torch.float32, device=desc.device)
        cnt = torch.zeros((P,), dtype=torch.int32, device=desc.device)
        st = torch.zeros((P, 2), dtype=torch.int32, device=desc.device) if want_stats else None
        N.check(L.sfm_match_ratio_batch(ctx, _stream_ptr(), desc.data_ptr(), counts.data_ptr(), n_sets, nmax,
                                        pairs.data_ptr(), P, float(np.float32(ratio_threshold)),
                                        mode | (N.SFM_MATCH_PREPARED if prepared else 0),
                                        ws.data_ptr(), ws.numel(), m.data_ptr(), c.data_ptr(), cnt.data_ptr(),
                                        st.data_ptr() if want_stats else None, cap), ctx)
        ws.record_stream(torch.cuda.current_stream())
    return (m, c, cnt, st) if want_stats else (m, c, cnt)


def match_workspace(n_sets: int, nmax: int, n_pairs: int, device) -> torch.Tensor:
    """A workspace for match_batch_device(ws=...) over up to n_pairs pairs per call."""
    nbytes = N.load_library().sfm_match_workspace_bytes(n_sets, nmax, n_pairs)
    return torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=device)


class NNRatioFeatureMatcher:
    """FeatureMatcher/NNRatioFeatureMatcher.py:4-60 on the GPU.

    match_features_ratio_test(features1, features2) -> (matches (k,2) int64,
    confidences (k,) float32) ordered by confidence ascending; empty results
    are the reference's `np.array([])` pair; features2 with fewer than two rows
    raises IndexError as the reference's `sorted_dists_idx[1]` does."""

    def __init__(self, ratio_threshold=0.8, mode: int = N.SFM_MATCH_AUTO):
        self.ratio_threshold = ratio_threshold
        self._mode = mode

    def match_features_ratio_test(self, features1: np.ndarray, features2: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        features1 = np.asarray(features1)
        features2 = np.asarray(features2)
        if features1.ndim != 2 or features2.ndim != 2 or features1.shape[1] != features2.shape[1]:
            raise ValueError("features must be (n, feat_dim) arrays of equal feat_dim")
        if features1.shape[1] != N.DESC_DIM:
            raise ValueError("libsfmb200 matches 128-d descriptors only")
        if features1.shape[0] == 0:
            return np.array([]), np.array([])
        if features2.shape[0] < 2:
            raise IndexError(f"index 1 is out of bounds for axis 0 with size {features2.shape[0]}")
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).pin_memory().to('cuda', non_blocking=True)
        m, c, cnt = match_device(up(features1), up(features2), self.ratio_threshold, self._mode)
        k = int(cnt.cpu()[0])
        if k == 0:
            return np.array([]), np.array([])
        return m[:k].cpu().numpy().astype(np.int64), c[:k].cpu().numpy()
