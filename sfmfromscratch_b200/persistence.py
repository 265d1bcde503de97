"""Match-graph and model persistence: the containers the reference
keeps its pair results in and the file it writes its reconstruction to.

  Matches          Runner.py:118-125 (per-pair record: matches, confidence, p1, p2, K1, K2)
  MatchGraph       Runner.py:171-172,353-355 (`all_matches[i][j]`, filled symmetrically) plus an
                   `.npz` form so that an all-pairs matching run (BASELINE configs[4]) can be fed
                   back into an SfM run -- the reference itself only ever fills consecutive pairs
  save_model /     Runner.py:357-359,403-416 (`output/<model>.npz` with `p3d`, `frame_idx`,
  load_model       `pt_idx`), key for key and dtype for dtype

Pure host bookkeeping: no arithmetic, hence nothing to run on the GPU; the arrays it stores come
from the device paths (matcher, `geometry.matches_to_coords_device`, `geometry.find_inliers_many`).
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


class Matches:
    """Runner.py:118-125."""

    def __init__(self, matches, confidence, p1, p2, K1, K2):
        self.matches = matches
        self.confidence = confidence
        self.p1 = p1
        self.p2 = p2
        self.K1 = K1
        self.K2 = K2


class MatchGraph:
    """`all_matches`: a (max_img + 1) x (max_img + 1) table of `Matches | None`, image ids starting
    at 1 as the reference's file names do (Runner.py:171-172)."""

    def __init__(self, max_img: int):
        self.max_img = int(max_img)
        self.all_matches: List[List[Optional[Matches]]] = [[None for _ in range(self.max_img + 1)]
                                                           for _ in range(self.max_img + 1)]

    def set_pair(self, i1: int, i2: int, matches, confidence, p1, p2, K1, K2) -> None:
        """Runner.py:353-355: the pair and its mirror image."""
        self.all_matches[i1][i2] = Matches(matches, confidence, p1, p2, K1, K2)
        self.all_matches[i2][i1] = Matches(matches, confidence, p2, p1, K2, K1)

    def __getitem__(self, ij: Tuple[int, int]) -> Optional[Matches]:
        return self.all_matches[ij[0]][ij[1]]

    def pairs(self) -> List[Tuple[int, int]]:
        return [(i, j) for i in range(self.max_img + 1) for j in range(i + 1, self.max_img + 1)
                if self.all_matches[i][j] is not None]

    @classmethod
    def from_batch(cls, max_img: int, pairs: Sequence[Tuple[int, int]], matches, conf, mcount, X, Y, K,
                   num_matches: int = 2500, inliers: Optional[Sequence] = None) -> "MatchGraph":
        """Assemble the graph from one batched matcher call: `pairs[k] = (i1, i2)` (1-based image ids),
        `matches[k]` (cap, 2) / `conf[k]` (cap,) / `mcount[k]` the matcher's host outputs for that pair,
        `X[i]`, `Y[i]` the int64 keypoint coordinates of image i, `K[i]` its intrinsics.  p1/p2 are
        `_convert_matches_to_coords` of the first `num_matches` matches (Runner.py:347) or, when given,
        `inliers[k] = (p1, p2)` from `find_inliers` (Runner.py:351)."""
        g = cls(max_img)
        for k, (i1, i2) in enumerate(pairs):
            n = int(mcount[k])
            m = np.asarray(matches[k][:n]).astype(np.int64)
            c = np.asarray(conf[k][:n])
            if inliers is not None:
                p1, p2 = inliers[k][0], inliers[k][1]
            elif n == 0:
                p1, p2 = np.array([]), np.array([])
            else:
                mm = m[:num_matches]
                p1 = np.column_stack((X[i1][mm[:, 0]], Y[i1][mm[:, 0]]))
                p2 = np.column_stack((X[i2][mm[:, 1]], Y[i2][mm[:, 1]]))
            g.set_pair(i1, i2, m if n else np.array([]), c if n else np.array([]), p1, p2, K[i1], K[i2])
        return g

    # ---- .npz form: ragged per-pair arrays stored concatenated with offset tables
    def save(self, path: str) -> None:
        pairs = self.pairs()
        recs = [self.all_matches[i][j] for i, j in pairs]

        def cat(arrs, width, dtype):
            arrs = [np.asarray(a).reshape(-1, width) if np.asarray(a).size else np.zeros((0, width), dtype) for a in arrs]
            off = np.concatenate([[0], np.cumsum([len(a) for a in arrs])]).astype(np.int64)
            return (np.concatenate(arrs).astype(dtype) if arrs else np.zeros((0, width), dtype)), off
        m, m_off = cat([r.matches for r in recs], 2, np.int64)
        c, _ = cat([r.confidence for r in recs], 1, np.float32)
        p1, p_off = cat([r.p1 for r in recs], 2, np.float64)
        p2, _ = cat([r.p2 for r in recs], 2, np.float64)
        np.savez(path, max_img=self.max_img, pairs=np.array(pairs, np.int64).reshape(-1, 2), matches=m, match_off=m_off,
                 confidence=c[:, 0], p1=p1, p2=p2, point_off=p_off,
                 p_dtype=np.array([str(np.asarray(r.p1).dtype) for r in recs]),
                 K1=np.array([r.K1 for r in recs], np.float64).reshape(-1, 3, 3),
                 K2=np.array([r.K2 for r in recs], np.float64).reshape(-1, 3, 3))

    @classmethod
    def load(cls, path: str) -> "MatchGraph":
        z = np.load(path)
        g = cls(int(z["max_img"]))
        for k, (i, j) in enumerate(z["pairs"]):
            a, b = z["match_off"][k], z["match_off"][k + 1]
            pa, pb = z["point_off"][k], z["point_off"][k + 1]
            dt = np.dtype(str(z["p_dtype"][k]))
            empty = np.array([])
            g.set_pair(int(i), int(j), z["matches"][a:b] if b > a else empty, z["confidence"][a:b] if b > a else empty,
                       z["p1"][pa:pb].astype(dt) if pb > pa else empty, z["p2"][pa:pb].astype(dt) if pb > pa else empty,
                       z["K1"][k], z["K2"][k])
        return g


def save_model(path: str, global_points_3D, frame_indices, point_indices) -> None:
    """Runner.py:357-359 (`np.savez('output/<model>.npz', p3d=..., frame_idx=..., pt_idx=...)`);
    `path` is the full file name."""
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    np.savez(path, p3d=np.array(global_points_3D), frame_idx=np.array(frame_indices), pt_idx=np.array(point_indices))


def load_model(path: str):
    """Runner.py:409-414: the three lists the reference's viewer receives."""
    npz = np.load(path)
    return npz["p3d"].tolist(), npz["frame_idx"].tolist(), npz["pt_idx"].tolist()
