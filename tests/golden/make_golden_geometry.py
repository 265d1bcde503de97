"""Golden vectors for the two-view RANSAC stage, produced by the
UNMODIFIED reference (`/root/reference/SFM.py`) in the build container:

    python tests/golden/make_golden_geometry.py

Writes tests/golden/geometry_ransac.npz: for each case the input correspondences, the arrays
`CameraPose.find_inliers` / `CameraPose.ransac_camera_motion` return, and (from a second pass that
repeats the reference's loop with its own helpers) the winning iteration and the per-iteration
inlier counts the reference discards.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SFM_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(1, REF)

from SFM import CameraPose  # noqa: E402  (the reference)
from sfmfromscratch_b200.synth import two_view_correspondences  # noqa: E402

FIND = [(600, 0, 0.3, 800), (50, 1, 0.5, 300), (9, 2, 0.0, 50), (2500, 3, 0.4, 200), (8, 4, 0.0, 20)]
POSE = [(40, 5, 0.0, 100), (100, 6, 0.02, 150), (60, 8, 0.3, 100)]


def ref_counts(p1, p2, iterations):
    """The reference's loop (SFM.py:133-158) re-run with its own helpers, keeping what it discards."""
    np.random.seed(5)
    counts, samples = [], []
    for _ in range(iterations):
        idx = np.random.choice(len(p1), 8, replace=False)
        F = CameraPose._compute_fundamental_matrix(p1[idx], p2[idx])
        a = np.column_stack((p1, np.ones(len(p1))))
        b = np.column_stack((p2, np.ones(len(p2))))
        lb = (F @ a.T).T
        d = np.abs(np.sum(lb * b, axis=1)) / np.sqrt(lb[:, 0] ** 2 + lb[:, 1] ** 2)
        counts.append(int(np.sum(d < 1.0)))
        samples.append(idx)
    return np.array(counts), np.array(samples)


def main():
    out = {}
    for k, (n, seed, outl, it) in enumerate(FIND):
        p1, p2, K = two_view_correspondences(n, seed, outl)
        a, b = CameraPose.find_inliers(p1, p2, max_iterations=it)
        counts, samples = ref_counts(p1, p2, it)
        out.update({f"find{k}_p1": p1, f"find{k}_p2": p2, f"find{k}_it": it, f"find{k}_in1": a, f"find{k}_in2": b,
                    f"find{k}_counts": counts, f"find{k}_samples": samples})
    for k, (n, seed, outl, it) in enumerate(POSE):
        p1, p2, K = two_view_correspondences(n, seed, outl)
        R, T, a, b = CameraPose(p1, p2, K, K).ransac_camera_motion(np.eye(3), np.zeros(3), max_iterations=it)
        out.update({f"pose{k}_p1": p1, f"pose{k}_p2": p2, f"pose{k}_K": K, f"pose{k}_it": it,
                    f"pose{k}_R": np.zeros((0,)) if R is None else R, f"pose{k}_T": np.zeros((0,)) if T is None else T,
                    f"pose{k}_in1": a, f"pose{k}_in2": b})
    np.savez_compressed(os.path.join(HERE, "geometry_ransac.npz"), **out)
    print("wrote geometry_ransac.npz:", {k: np.asarray(v).shape for k, v in out.items() if k.endswith("in1")})


if __name__ == "__main__":
    main()
