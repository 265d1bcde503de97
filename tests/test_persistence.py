"""Match-graph / model persistence: Runner.py:118-125,171-172,353-359,403-416."""
import os
import sys
import types

import numpy as np
import pytest

from sfmfromscratch_b200.persistence import MatchGraph, Matches, load_model, save_model


def _graph():
    rng = np.random.default_rng(0)
    X = {i: rng.integers(0, 900, 50).astype(np.int64) for i in range(1, 5)}
    Y = {i: rng.integers(0, 500, 50).astype(np.int64) for i in range(1, 5)}
    K = {i: np.array([[800.0 + i, 0, 450], [0, 800.0 + i, 250], [0, 0, 1]]) for i in range(1, 5)}
    pairs = [(1, 2), (2, 3), (1, 4)]
    matches = rng.integers(0, 50, (3, 40, 2)).astype(np.int32)
    conf = np.sort(rng.random((3, 40)).astype(np.float32), axis=1)
    mcount = np.array([40, 0, 17])
    return MatchGraph.from_batch(4, pairs, matches, conf, mcount, X, Y, K, num_matches=25), X, Y, K, matches, conf


def test_match_graph_layout_and_mirror():
    g, X, Y, K, matches, conf = _graph()
    assert len(g.all_matches) == 5 and all(len(r) == 5 for r in g.all_matches)        # (max_img + 1)^2, ids from 1
    a, b = g[1, 2], g[2, 1]
    assert isinstance(a, Matches) and a.matches.dtype == np.int64 and a.matches.shape == (40, 2)
    assert a.p1.shape == (25, 2) and a.p1.dtype == np.int64                            # first num_matches only
    assert np.array_equal(a.p1[:, 0], X[1][matches[0, :25, 0]]) and np.array_equal(a.p2[:, 1], Y[2][matches[0, :25, 1]])
    assert np.array_equal(b.p1, a.p2) and np.array_equal(b.p2, a.p1) and b.K1 is a.K2 and b.K2 is a.K1
    assert g[2, 3].matches.shape == (0,) and g[2, 3].p1.shape == (0,)                  # the reference's np.array([]) pair
    assert g[3, 4] is None and g.pairs() == [(1, 2), (1, 4), (2, 3)]


def test_match_graph_npz_round_trip(tmp_path):
    g, *_ = _graph()
    p = str(tmp_path / "graph.npz")
    g.save(p)
    h = MatchGraph.load(p)
    assert h.pairs() == g.pairs()
    for i, j in g.pairs():
        for (x, y) in ((g[i, j], h[i, j]), (g[j, i], h[j, i])):
            for f in ("matches", "confidence", "p1", "p2", "K1", "K2"):
                u, v = np.asarray(getattr(x, f)), np.asarray(getattr(y, f))
                assert u.shape == v.shape and u.dtype == v.dtype and np.array_equal(u, v), (i, j, f)


def test_model_file_matches_the_references(tmp_path, monkeypatch):
    """The file save_model writes is the file SFMRunner.save_data writes (same keys, dtypes, values),
    and the reference's loader reads ours."""
    rng = np.random.default_rng(1)
    p3d = [rng.normal(size=3) for _ in range(30)]
    frames = [int(v) for v in rng.integers(0, 4, 55)]
    pts = [int(v) for v in rng.integers(0, 30, 55)]
    ours = str(tmp_path / "output" / "m.npz")
    save_model(ours, p3d, frames, pts)
    l3, lf, lp = load_model(ours)
    assert l3 == np.array(p3d).tolist() and lf == frames and lp == pts
    ref_root = os.environ.get("SFM_REFERENCE", "/root/reference")
    if not os.path.exists(os.path.join(ref_root, "Runner.py")):
        return                                                                          # GPU box: format pinned above
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    sys.path.insert(0, ref_root)                                                        # Runner.py imports its siblings by name
    try:
        import make_golden_geometry as MG
        R = MG._runner_module()
    finally:
        sys.path[:] = [q for q in sys.path if q != ref_root]
        for k in [k for k in sys.modules if k.split(".")[0] in ("FeatureExtractor", "FeatureMatcher", "Runner", "SFM", "PoseEstimator",
                                                                "Util", "Visualizer", "matplotlib", "mpl_toolkits", "make_golden_geometry")]:
            del sys.modules[k]
    monkeypatch.chdir(tmp_path)
    self = types.SimpleNamespace(model_name="ref", global_points_3D=p3d, frame_indices=frames, point_indices=pts)
    R.SFMRunner.save_data(self)
    a, b = np.load(str(tmp_path / "output" / "ref.npz")), np.load(ours)
    assert sorted(a.files) == sorted(b.files) == ["frame_idx", "p3d", "pt_idx"]
    for k in a.files:
        assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k])
    seen = {}
    R.V3D = lambda *args: seen.setdefault("args", args)                                # the viewer the loader hands over to
    os.replace(ours, str(tmp_path / "output" / "mine.npz"))
    R.SFMRunner.load("mine")
    assert seen["args"][0] == np.array(p3d).tolist() and seen["args"][1] == frames and seen["args"][2] == pts
