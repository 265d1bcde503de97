"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (reesque/SfmFromScratch, mounted read-only at /root/reference) in the
build container.  The reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so these fixtures are what pins the oracle and the CUDA
path.  Run from the repo root:

    python tests/golden/make_golden.py

Library versions the outputs depend on are recorded in versions.json.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SFM_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(1, REF)

import cv2  # noqa: E402
from FeatureExtractor import NaiveSIFT, ScaleRotInvSIFT  # noqa: E402  (the reference)
from FeatureMatcher import NNRatioFeatureMatcher  # noqa: E402  (the reference)
from sfmfromscratch_b200.synth import second_view, synth_descriptors, synth_image  # noqa: E402


def run_srs(img, params):
    e = ScaleRotInvSIFT(img, params)
    X, Y = e.detect_keypoints()
    return dict(X=np.asarray(X), Y=np.asarray(Y), desc=np.asarray(e.extract_descriptors()),
                pyramid=[np.asarray(p) for p in e._img_pyramid])


def harris_parts(img, params):
    """Re-run the reference's private Harris stage to capture R and the per-level keypoints."""
    e = NaiveSIFT(img, params)
    Ix, Iy = e._compute_image_gradients(img)
    gk = e._generate_gaussian_kernel(e._gaussian_size, e._sigma)
    Sxx = cv2.filter2D(Ix ** 2, ddepth=-1, kernel=gk, borderType=cv2.BORDER_CONSTANT)
    Sxy = cv2.filter2D(Ix * Iy, ddepth=-1, kernel=gk, borderType=cv2.BORDER_CONSTANT)
    Syy = cv2.filter2D(Iy ** 2, ddepth=-1, kernel=gk, borderType=cv2.BORDER_CONSTANT)
    R = (Sxx * Syy) - (Sxy ** 2) - e._alpha * ((Sxx + Syy) ** 2)       # NaiveSIFT.py:71-74
    return R, gk.astype(np.float32)


def main():
    out = {}
    # 1. two-view pair, defaults (configs[0] in miniature)
    p1 = {'num_interest_points': 600}
    a = synth_image(96, 128, 0)
    b = second_view(a, 1)
    ra, rb = run_srs(a, p1), run_srs(b, p1)
    m, c = NNRatioFeatureMatcher(0.8).match_features_ratio_test(ra['desc'], rb['desc'])
    R, gk = harris_parts(a, p1)
    np.savez_compressed(os.path.join(HERE, "two_view_96x128.npz"), img1=a, img2=b,
                        X1=ra['X'], Y1=ra['Y'], D1=ra['desc'], X2=rb['X'], Y2=rb['Y'], D2=rb['desc'],
                        matches=m, conf=c, R1=R, gauss=gk,
                        pyr1=ra['pyramid'][1], pyr2=ra['pyramid'][2], pyr3=ra['pyramid'][3])
    out['two_view_96x128'] = dict(n1=int(len(ra['X'])), n2=int(len(rb['X'])), matches=int(len(m)))
    # 1b. a larger two-view pair (about 550 keypoints per view)
    a = synth_image(240, 320, 0)
    b = second_view(a, 1)
    ra, rb = run_srs(a, {}), run_srs(b, {})
    m, c = NNRatioFeatureMatcher(0.8).match_features_ratio_test(ra['desc'], rb['desc'])
    np.savez_compressed(os.path.join(HERE, "two_view_240x320.npz"), img1=a, img2=b,
                        X1=ra['X'], Y1=ra['Y'], D1=ra['desc'], X2=rb['X'], Y2=rb['Y'], D2=rb['desc'],
                        matches=m, conf=c)
    out['two_view_240x320'] = dict(n1=int(len(ra['X'])), n2=int(len(rb['X'])), matches=int(len(m)))
    # 2. main.py-style parameters: non-integer pyramid factor, ksize 3, feature_width 18
    p2 = {'num_interest_points': 900, 'ksize': 3, 'gaussian_size': 7, 'sigma': 6, 'alpha': 0.05,
          'feature_width': 18, 'pyramid_level': 3, 'pyramid_scale_factor': 1.1}
    img = synth_image(120, 160, 7)
    r2 = run_srs(img, p2)
    np.savez_compressed(os.path.join(HERE, "srs_mainpy_120x160.npz"), img=img, X=r2['X'], Y=r2['Y'], D=r2['desc'],
                        pyr1=r2['pyramid'][1], pyr2=r2['pyramid'][2], params=json.dumps(p2))
    out['srs_mainpy_120x160'] = dict(n=int(len(r2['X'])))
    # 3. odd sizes: the halving is not exact, every level goes through the bilinear resize
    img = synth_image(101, 135, 11)
    r3 = run_srs(img, {'num_interest_points': 400})
    np.savez_compressed(os.path.join(HERE, "srs_odd_101x135.npz"), img=img, X=r3['X'], Y=r3['Y'], D=r3['desc'],
                        pyr1=r3['pyramid'][1], pyr2=r3['pyramid'][2], pyr3=r3['pyramid'][3])
    out['srs_odd_101x135'] = dict(n=int(len(r3['X'])))
    # 4. NaiveSIFT
    img = synth_image(96, 128, 3)
    e = NaiveSIFT(img, {'num_interest_points': 300})
    X, Y = e.detect_keypoints()
    D = e.extract_descriptors()
    np.savez_compressed(os.path.join(HERE, "naive_96x128.npz"), img=img, X=X, Y=Y, D=D, conf=e.confidences)
    out['naive_96x128'] = dict(n=int(len(X)))
    # 5. matcher on synthetic descriptors with planted matches, exact duplicates and an all-zero row
    f1 = synth_descriptors(220, 0, planted=0.6)
    f2 = synth_descriptors(260, 1, planted=0.6)
    f2[5] = f2[17]                 # duplicate train rows: d0 == d1 for any query matching them
    f1[3] = f2[40]                 # exact hit: d0 == 0
    f1[9] = 0.0                    # zero descriptor
    m, c = NNRatioFeatureMatcher(0.8).match_features_ratio_test(f1, f2)
    m2, c2 = NNRatioFeatureMatcher(0.95).match_features_ratio_test(f1, f2)
    np.savez_compressed(os.path.join(HERE, "matcher_220x260.npz"), f1=f1, f2=f2, matches=m, conf=c,
                        matches95=m2, conf95=c2)
    out['matcher_220x260'] = dict(matches=int(len(m)), matches95=int(len(m2)))
    # 6. image ingest: Runner.py's own helpers on an 8-bit RGB image
    import importlib.util, tempfile, types
    from PIL import Image

    class _Stub(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.widgets", "matplotlib.cm", "matplotlib.colors",
                 "mpl_toolkits", "mpl_toolkits.mplot3d"):          # Runner.py imports the (absent) GUI stack
        if name not in sys.modules:
            m = _Stub(name); m.__path__ = []; sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("Runner", os.path.join(REF, "Runner.py"))
    R = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(R)
    rng = np.random.default_rng(42)
    base = rng.integers(0, 256, (40, 52, 3))
    rgb = np.clip(np.kron(base, np.ones((4, 4, 1)))[:151, :203] + rng.integers(-25, 26, (151, 203, 3)), 0, 255).astype(np.uint8)
    grays = {}
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "a.png")
        Image.fromarray(rgb).save(path)
        for sf in (0.5, 0.3):
            a = R._load_image(path)                                                    # Runner.py:33
            a = R._PIL_resize(a, (int(a.shape[1] * sf), int(a.shape[0] * sf)))        # :37-39
            grays[sf] = R._rgb2gray(a)                                                 # :45
    np.savez_compressed(os.path.join(HERE, "ingest_151x203.npz"), rgb=rgb, gray05=grays[0.5], gray03=grays[0.3])
    out['ingest_151x203'] = dict(shape05=list(grays[0.5].shape), shape03=list(grays[0.3].shape), pillow=Image.__version__)
    out['versions'] = dict(numpy=np.__version__, opencv=cv2.__version__,
                           python=sys.version.split()[0], ipp=str(cv2.ipp.getIppVersion()))
    with open(os.path.join(HERE, "versions.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
