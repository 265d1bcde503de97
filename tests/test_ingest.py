"""Image ingest: Runner.py:33-46."""
import os
import sys
import types

import numpy as np
import pytest

from oracle import oracle as O


def _rgb(h, w, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h // 4 + 2, w // 4 + 2, 3))
    img = np.kron(base, np.ones((4, 4, 1)))[:h, :w] + rng.integers(-20, 21, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("h,w,ow,oh", [(40, 60, 30, 20), (41, 63, 31, 20), (100, 80, 24, 30), (64, 64, 64, 32), (33, 47, 60, 50)])
def test_pil_restatement_matches_pillow(h, w, ow, oh):
    from PIL import Image
    img = _rgb(h, w, h + w)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh)))
    assert np.array_equal(O.pil_bicubic_resize(img, (ow, oh)), ref)


def test_oracle_ingest_matches_reference_functions():
    """The reference's own helpers (Runner.py imports matplotlib, absent here: stubbed)."""
    ref_root = os.environ.get("SFM_REFERENCE", "/root/reference")
    if not os.path.exists(os.path.join(ref_root, "Runner.py")):
        pytest.skip("reference tree not present (GPU box)")
    import importlib.util
    import tempfile
    from PIL import Image
    saved = dict(sys.modules)
    saved_path = list(sys.path)
    try:
        class _Stub(types.ModuleType):                       # any attribute of the GUI modules resolves to a dummy
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
        for k in [k for k in sys.modules if k.split(".")[0] in ("FeatureExtractor", "FeatureMatcher", "Runner", "SFM", "PoseEstimator", "Util", "Visualizer")]:
            del sys.modules[k]
        sys.path.insert(0, ref_root)
        try:
            spec = importlib.util.spec_from_file_location("Runner", os.path.join(ref_root, "Runner.py"))
            R = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(R)
        except Exception as e:                               # other missing GUI bits: nothing to compare against
            pytest.skip(f"reference Runner.py not importable here: {e}")
        img = _rgb(90, 122, 3)
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "a.png")
            Image.fromarray(img).save(path)
            a = R._load_image(path)
            a = R._PIL_resize(a, (int(a.shape[1] * 0.5), int(a.shape[0] * 0.5)))
            ref = R._rgb2gray(a)
        got = O.ingest_gray(img, 0.5)
        assert got.dtype == ref.dtype == np.float32 and np.array_equal(got, ref)
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):                          # drop only the reference's modules and the GUI stubs
            if k not in saved and k.split(".")[0] in ("FeatureExtractor", "FeatureMatcher", "Runner", "SFM", "PoseEstimator",
                                                      "Util", "Visualizer", "matplotlib", "mpl_toolkits"):
                del sys.modules[k]


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,scale", [(96, 128, 0.5), (97, 131, 0.5), (240, 320, 0.3), (60, 50, 1.0), (50, 70, 1.7), (1080, 1920, 0.5)])
def test_gpu_ingest_bit_exact(h, w, scale):
    from sfmfromscratch_b200 import ingest
    img = _rgb(h, w, 7)
    got = ingest.gray_from_rgb8(img, scale)
    ref = O.ingest_gray(img, scale)
    assert got.shape == ref.shape and got.dtype == np.float32
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.mark.gpu
def test_gpu_ingest_file_and_batch(tmp_path):
    import torch
    from PIL import Image
    from sfmfromscratch_b200 import ScaleRotInvSIFT, ingest
    imgs = np.stack([_rgb(120, 160, s) for s in range(3)])
    out = ingest.gray_from_rgb8_device(torch.from_numpy(imgs).cuda(), (60, 80)).cpu().numpy()
    for b in range(3):
        assert np.array_equal(out[b], O.ingest_gray(imgs[b], 0.5))
    p = tmp_path / "img.png"
    Image.fromarray(imgs[0]).save(p)
    g = ingest.load_image_gray(str(p), 0.5)
    assert np.array_equal(g, out[0])
    e, o = ScaleRotInvSIFT(g, {'num_interest_points': 200}), O.ScaleRotInvSIFT(O.ingest_gray(imgs[0], 0.5), {'num_interest_points': 200})
    assert np.array_equal(e.detect_keypoints()[0], o.detect_keypoints()[0])


def test_oracle_ingest_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "ingest_151x203.npz"))
    for sf, key in ((0.5, "gray05"), (0.3, "gray03")):
        got = O.ingest_gray(g["rgb"], sf)
        assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), g[key].view(np.uint32))


@pytest.mark.gpu
def test_gpu_ingest_matches_golden(golden_dir):
    from sfmfromscratch_b200 import ingest
    g = np.load(os.path.join(golden_dir, "ingest_151x203.npz"))
    for sf, key in ((0.5, "gray05"), (0.3, "gray03")):
        got = ingest.gray_from_rgb8(g["rgb"], sf)
        assert np.array_equal(got.view(np.uint32), g[key].view(np.uint32))
