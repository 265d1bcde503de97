"""Golden fixture for the drop-in proof through the reference's OWN caller: run the unmodified
Runner.FeatureRunner (Runner.py:22-73: _load_image -> _PIL_resize -> _rgb2gray -> extractor x2 -> matcher) of
/root/reference on two PNG files and record what it computes.  tests/test_dropin_runner.py then runs the same
class with sfmfromscratch_b200/dropin shadowing the FeatureExtractor / FeatureMatcher packages and compares.

    python tests/golden/make_golden_runner.py
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SFM_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from refstub import reference_path  # noqa: E402
from sfmfromscratch_b200.synth import second_view, synth_image  # noqa: E402

MAINPY = {'num_interest_points': 2500, 'ksize': 3, 'gaussian_size': 7, 'sigma': 6, 'alpha': 0.05,
          'feature_width': 18, 'pyramid_level': 3, 'pyramid_scale_factor': 1.1}          # main.py:19-28
CASES = {"mainpy": (MAINPY, 0.85), "defaults": ({'num_interest_points': 1200}, 0.8)}


def rgb_pair(h=240, w=320):
    """Two colour views: three differently weighted mixes of two noise fields, second view warped + noisy."""
    a, b = synth_image(h, w, 21), synth_image(h, w, 22, sigma=3.0)
    v1 = np.stack([0.8 * a + 0.2 * b, 0.5 * a + 0.5 * b, 0.3 * a + 0.7 * b], axis=2)
    v2 = np.stack([second_view(v1[:, :, c].astype(np.float32), 30 + c) for c in range(3)], axis=2)
    return (np.clip(v1, 0, 1) * 255).astype(np.uint8), (np.clip(v2, 0, 1) * 255).astype(np.uint8)


def main():
    import PIL.Image
    im1, im2 = rgb_pair()
    p1, p2 = os.path.join(HERE, "runner_im1.png"), os.path.join(HERE, "runner_im2.png")
    PIL.Image.fromarray(im1).save(p1)
    PIL.Image.fromarray(im2).save(p2)
    out, info = {}, {}
    with reference_path(REF):
        import Runner
        from FeatureExtractor import ScaleRotInvSIFT
        assert Runner.__file__.startswith(REF) and sys.modules["FeatureExtractor"].__file__.startswith(REF)
        for name, (params, thr) in CASES.items():
            with contextlib.redirect_stdout(io.StringIO()):
                fr = Runner.FeatureRunner(p1, p2, scale_factor=0.5, feature_extractor_class=ScaleRotInvSIFT,
                                          extractor_params=dict(params), match_threshold=thr)
            out.update({f"{name}_X1": np.asarray(fr.X1), f"{name}_Y1": np.asarray(fr.Y1), f"{name}_D1": np.asarray(fr.descriptors1),
                        f"{name}_X2": np.asarray(fr.X2), f"{name}_Y2": np.asarray(fr.Y2), f"{name}_D2": np.asarray(fr.descriptors2),
                        f"{name}_matches": np.asarray(fr.matches), f"{name}_conf": np.asarray(fr.confidences),
                        f"{name}_bw1": np.asarray(fr._image1_bw), f"{name}_bw2": np.asarray(fr._image2_bw)})
            info[name] = dict(n1=int(len(fr.X1)), n2=int(len(fr.X2)), matches=int(len(fr.matches)), params=params, thr=thr)
    np.savez_compressed(os.path.join(HERE, "runner_two_view.npz"), cases=json.dumps(CASES), **out)
    print(json.dumps(info))


if __name__ == "__main__":
    main()
