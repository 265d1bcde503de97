"""CPU oracle for the two-view geometry that consumes the matcher's output
(SURVEY.md section 8f rows 2 and 3): match -> coordinate conversion, 8-point
RANSAC (``find_inliers``, ``ransac_camera_motion``) and the brute-force 2-D /
3-D nearest-point association loops.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module; the product package never does.

A restatement (not a copy) of the reference's arithmetic: every hypothesis is
evaluated with the same numpy / LAPACK primitives in the same order as the
reference so that the results are bit-identical to it on one machine
(``tests/test_oracle_geometry_vs_reference.py`` checks that against the live
reference, ``tests/golden/geometry_*.npz`` pins it where the reference cannot
travel).  The hypothesis loop additionally records per-hypothesis data
(fundamental matrices, inlier counts, validity flags) that the reference throws
away, so that the CUDA path can be compared stage by stage.

Reference citations are relative to the reference root.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

RANSAC_SEED = 5            # SFM.py:45,133: np.random.seed(5) at the start of every call


def num_ransac_iterations(prob_success: float, sample_size: int, ind_prob_correct: float) -> int:
    """SFM.py:185-187."""
    return int(np.log(1 - prob_success) / np.log(1 - (ind_prob_correct ** sample_size)))


def convert_matches_to_coords(matches, X1, Y1, X2, Y2, num_matches: int = 2500):
    """Runner.py:423-434: the first ``num_matches`` matches as two (k,2) coordinate arrays."""
    if matches.shape[0] == 0:
        return np.array([]), np.array([])
    m = matches[:num_matches]
    return (np.column_stack((X1[m[:, 0]], Y1[m[:, 0]])),
            np.column_stack((X2[m[:, 1]], Y2[m[:, 1]])))


def sample_indices(n: int, iterations: int, seed: int = RANSAC_SEED) -> np.ndarray:
    """The 8-subsets ``np.random.choice(n, 8, replace=False)`` draws after
    ``np.random.seed(5)`` (SFM.py:45-49,133-137): MT19937 legacy stream, one full
    Fisher-Yates permutation of ``n`` per draw."""
    rs = np.random.RandomState(seed)
    return np.stack([rs.choice(n, 8, replace=False) for _ in range(iterations)]) if iterations else np.zeros((0, 8), np.int64)


def _normalize(points_h: np.ndarray):
    """SFM.py:163-178 (points are (n,3) homogeneous rows)."""
    c = np.mean(points_h[:, :2], axis=0)
    d = np.sqrt((points_h[:, 0] - c[0]) ** 2 + (points_h[:, 1] - c[1]) ** 2)
    s = np.sqrt(2) / np.mean(d)
    T = np.array([[s, 0, -s * c[0]], [0, s, -s * c[1]], [0, 0, 1]])
    return points_h @ T.T, T


def fundamental_8pt(p1: np.ndarray, p2: np.ndarray) -> np.ndarray:
    """SFM.py:189-236: normalised 8-point estimate with the rank-2 projection."""
    n = p1.shape[0]
    a, T1 = _normalize(np.hstack([p1, np.ones((n, 1))]))
    b, T2 = _normalize(np.hstack([p2, np.ones((n, 1))]))
    A = np.zeros((n, 9))
    for i in range(n):
        x1, y1, x2, y2 = a[i, 0], a[i, 1], b[i, 0], b[i, 1]
        A[i] = [x1 * x2, y1 * x2, x2, x1 * y2, y1 * y2, y2, x1, y1, 1]
    F = np.linalg.svd(A)[2][-1, :].reshape(3, 3)
    U, D, Vt = np.linalg.svd(F)
    D[2] = 0
    return T2.T @ np.dot(U, np.dot(np.diag(D), Vt)) @ T1


def epipolar_distances(F: np.ndarray, p1: np.ndarray, p2: np.ndarray) -> np.ndarray:
    """SFM.py:143-151: distance of p2 from the epipolar line F p1."""
    ah = np.column_stack((p1, np.ones(len(p1))))
    bh = np.column_stack((p2, np.ones(len(p2))))
    lb = (F @ ah.T).T
    return np.abs(np.sum(lb * bh, axis=1)) / np.sqrt(lb[:, 0] ** 2 + lb[:, 1] ** 2)


def find_inliers(p1, p2, threshold: float = 1.0, max_iterations: int = 1000, detail: Optional[dict] = None):
    """SFM.py:126-160.  Returns the reference's values (a 4-tuple of None below 8
    points, else the two inlier arrays of the first hypothesis with the largest
    count).  ``detail`` (a dict) receives ``F`` (it,3,3), ``counts`` (it,),
    ``best`` (index or -1) and ``mask`` of the winner."""
    if len(p1) < 8:
        return None, None, None, None
    idx = sample_indices(len(p1), max_iterations)
    best1, best2, best, best_mask = [], [], -1, np.zeros(len(p1), bool)
    Fs = np.zeros((max_iterations, 3, 3))
    counts = np.zeros(max_iterations, np.int64)
    for it in range(max_iterations):
        F = fundamental_8pt(p1[idx[it]], p2[idx[it]])
        mask = epipolar_distances(F, p1, p2) < threshold
        Fs[it], counts[it] = F, np.sum(mask)
        if counts[it] > len(best1):
            best1, best2, best, best_mask = p1[mask], p2[mask], it, mask
    if detail is not None:
        detail.update(F=Fs, counts=counts, best=best, mask=best_mask, samples=idx)
    return np.array(best1), np.array(best2)


def projection_matrix(R, t, K):
    """SFM.py:308-309."""
    return K @ np.hstack([R, t.reshape(-1, 1)])


def triangulate_point(x1, x2, P1, P2):
    """SFM.py:239-253: DLT, smallest right singular vector, dehomogenised."""
    A = np.vstack([x1[0] * P1[2, :] - P1[0, :], x1[1] * P1[2, :] - P1[1, :],
                   x2[0] * P2[2, :] - P2[0, :], x2[1] * P2[2, :] - P2[1, :]])
    X = np.linalg.svd(A)[2][-1]
    X /= X[3]
    return X[:3]


def check_valid_pose(p1, p2, K1, K2, R_base, T_base, R_c, T_c) -> bool:
    """SFM.py:104-124: every correspondence must triangulate in front of both cameras."""
    P1 = projection_matrix(R_base, T_base, K1)
    P2 = projection_matrix(R_c, T_c, K2)
    for i in range(len(p1)):
        X = triangulate_point(np.array([p1[i, 0], p1[i, 1], 1]), np.array([p2[i, 0], p2[i, 1], 1]), P1, P2)
        if (R_base @ X + T_base)[2] < 1e-6 or (R_c @ X + T_c)[2] < 1e-6:
            return False
    return True


def pose_candidates(F, K1, K2):
    """SFM.py:56-80: the four (R, T) decompositions of E = K2^T F K1 in the reference's order."""
    U, _, Vt = np.linalg.svd(K2.T @ F @ K1)
    W = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]])
    R1 = np.dot(U, np.dot(W, Vt))
    R2 = np.dot(U, np.dot(W.T, Vt))
    if np.linalg.det(R1) < 0:
        R1 = R1 * -1
    if np.linalg.det(R2) < 0:
        R2 = R2 * -1
    T = U[:, 2]
    return [(R1, T), (R1, -T), (R2, T), (R2, -T)]


def ransac_camera_motion(p1, p2, K1, K2, R_base, T_base, threshold: float = 1.0, max_iterations: int = 1000,
                         detail: Optional[dict] = None):
    """SFM.py:38-102."""
    if len(p1) < 8:
        return None, None, None, None
    idx = sample_indices(len(p1), max_iterations)
    best1, best2, best_r, best_t, best = [], [], None, None, -1
    Fs = np.zeros((max_iterations, 3, 3))
    counts = np.zeros(max_iterations, np.int64)
    valid = np.zeros((max_iterations, 4), bool)
    for it in range(max_iterations):
        F = fundamental_8pt(p1[idx[it]], p2[idx[it]])
        Fs[it] = F
        mask = None
        for c, (Rc, Tc) in enumerate(pose_candidates(F, K1, K2)):
            if not check_valid_pose(p1, p2, K1, K2, R_base, T_base, Rc, Tc):
                continue
            valid[it, c] = True
            if mask is None:
                mask = epipolar_distances(F, p1, p2) < threshold
                counts[it] = np.sum(mask)
            if counts[it] > len(best1):
                best1, best2, best_r, best_t, best = p1[mask], p2[mask], Rc, Tc, it
    if detail is not None:
        detail.update(F=Fs, counts=counts, best=best, valid=valid, samples=idx)
    return best_r, best_t, np.array(best1), np.array(best2)


# ---------------------------------------------------------------- association

def euclidean_distance(arr1, arr2):
    """SFM.py:376-382."""
    if arr2.shape[0] == 1:
        return np.linalg.norm(arr1 - arr2, axis=1)
    return np.linalg.norm(arr1[:, np.newaxis] - arr2, axis=2)


def associate_prev_frame(points_2d_prev, prev_frame_2d, dist_threshold: float = 5.0):
    """Runner.py:241-247: for every row q of ``prev_frame_2d`` the first nearest row of
    ``points_2d_prev``; kept when that distance is below the threshold.  Returns
    (kept q indices, their nearest indices)."""
    q_idx, n_idx = [], []
    for q in range(prev_frame_2d.shape[0]):
        d = euclidean_distance(points_2d_prev, prev_frame_2d[q:q + 1])
        m = np.argmin(d)
        if d[m] < dist_threshold:
            q_idx.append(q)
            n_idx.append(m)
    return np.array(q_idx, np.int64), np.array(n_idx, np.int64)


def dedup_points(points_3d, existing=None, threshold: float = 1e-6):
    """Runner.py:361-385 (``add_points`` / ``is_new_point`` / ``find_existing_point``): walks
    ``points_3d`` in order; a point at least ``threshold`` away from every stored point is
    appended to the store, otherwise it maps to the first nearest stored point.  Returns
    (point index per input row, the grown store)."""
    store = [] if existing is None else [np.asarray(p) for p in existing]
    out = []
    for p in points_3d:
        if not store:
            store.append(p)
            out.append(0)
            continue
        d = euclidean_distance(np.array(store), p[np.newaxis])
        if np.min(d) >= threshold:
            store.append(p)
            out.append(len(store) - 1)
        else:
            out.append(int(np.argmin(d)))
    return np.array(out, np.int64), np.array(store)
