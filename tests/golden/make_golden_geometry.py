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


# ---------------------------------------------------------------- association (section 8f row 3)

def _runner_module():
    """Runner.py imports matplotlib (absent here) for its plots: stub the GUI modules."""
    import importlib.util
    import types

    class _Stub(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.widgets", "matplotlib.cm", "matplotlib.colors",
                 "mpl_toolkits", "mpl_toolkits.mplot3d"):
        if name not in sys.modules:
            m = _Stub(name)
            m.__path__ = []
            sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("Runner", os.path.join(REF, "Runner.py"))
    R = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(R)
    return R


def assoc_main():
    import types
    R = _runner_module()
    out = {}
    rng = np.random.default_rng(7)
    # Runner.py:241-247 replayed with the reference's own distance helper
    for k, (m, q, thr) in enumerate([(400, 300, 5.0), (50, 80, 2.0), (1, 5, 5.0)]):
        prev = rng.integers(0, 960, (m, 2)).astype(np.int64)
        qry = prev[rng.integers(0, m, q)] + rng.integers(-6, 7, (q, 2))
        if m > 10:
            prev[5] = prev[3]                                   # duplicated triangulated point: first index wins
        rows, near = [], []
        for p_prime in range(qry.shape[0]):
            dist = CameraPose.compute_euclidean_distance(prev, qry[p_prime:p_prime + 1])
            mask = np.argmin(dist)
            if dist[mask] < thr:
                rows.append(p_prime)
                near.append(mask)
        out.update({f"assoc{k}_prev": prev, f"assoc{k}_query": qry, f"assoc{k}_thr": thr,
                    f"assoc{k}_rows": np.array(rows, np.int64), f"assoc{k}_near": np.array(near, np.int64)})
    # Runner.py:361-385: SFMRunner.add_points on a bare object carrying the attributes it touches
    for k, (n1, n2) in enumerate([(120, 90), (40, 60)]):
        self = types.SimpleNamespace(global_points_3D=[], global_points_2D=[], frame_indices=[], point_indices=[])
        for name in ("is_new_point", "find_existing_point", "add_points"):
            setattr(self, name, types.MethodType(getattr(R.SFMRunner, name), self))
        a = rng.normal(size=(n1, 3)) * 5
        a[10] = a[2]; a[11] = a[2] + 3e-7; a[30] = a[29] + 9e-7   # duplicates inside the first batch
        b = rng.normal(size=(n2, 3)) * 5
        b[::3] = a[rng.integers(0, n1, len(b[::3]))]              # re-observed points in the second batch
        b[1] = b[0] + 5e-7
        self.add_points(a, rng.integers(0, 900, (n1, 2)), 0)
        first = len(self.point_indices)
        self.add_points(b, rng.integers(0, 900, (n2, 2)), 1)
        out.update({f"dedup{k}_a": a, f"dedup{k}_b": b, f"dedup{k}_idx_a": np.array(self.point_indices[:first]),
                    f"dedup{k}_idx_b": np.array(self.point_indices[first:]), f"dedup{k}_store": np.array(self.global_points_3D)})
    np.savez_compressed(os.path.join(HERE, "geometry_assoc.npz"), **out)
    print("wrote geometry_assoc.npz:", {k: np.asarray(v).shape for k, v in out.items() if "idx" in k or "rows" in k})


if __name__ == "__main__":
    assoc_main()
