"""The persistent Harris stream (csrc/harris_stream.cuh) on the shapes the default policy keeps away from it.

By default a pyramid level goes to k_harris_stream only when it has >= 24 bands of 64x16 pixels per SM (a batch of
1080p frames); smaller levels run the one-tile-per-CTA kernel.  SFM_OPT_HARRIS_STREAM_MIN_BANDS = 0 sends every
level the stream can take (W % 4 == 0, W >= 76) through it, so these tests cover what a big batch never shows:
one band per CTA, runs of one to three bands, strips cut by the right edge, bands cut by the bottom edge, several
images inside one CTA's range (histogram epochs), the fused next pyramid level from partial chunks.  Everything is
compared with the CPU oracle (NaiveSIFT.py:60-118, ScaleRotInvSIFT.py:24-115): keypoints and responses
bit-identical, descriptors within tolerance with no keypoint needing the bin-edge exemption."""
import numpy as np
import pytest

from parity import DescriptorExplainer, assert_descriptors_close, assert_keypoints_equal

pytestmark = pytest.mark.gpu


@pytest.fixture()
def stream_everywhere():
    from sfmfromscratch_b200 import _native as N
    N.load_library()
    N.set_option(N.SFM_OPT_HARRIS_STREAM_MIN_BANDS, 0)
    yield N
    N.set_option(N.SFM_OPT_HARRIS_STREAM_MIN_BANDS, 24)


def _kernels(N, fn):
    N.profile_enable(True)
    out = fn()
    import torch
    torch.cuda.synchronize()
    st = N.profile_collect()
    N.profile_enable(False)
    return out, st


@pytest.mark.parametrize("h,w,nimg,params", [
    (96, 128, 1, {'num_interest_points': 600}),                       # 12 bands on 12 CTAs: every run is one band
    (240, 320, 3, {}),                                                # levels 240x320 and 120x160 (right-edge strip half empty)
    (135, 240, 5, {'pyramid_level': 2}),                              # bottom band cut (135 = 8 * 16 + 7); level 1 is 67x120
    (250, 332, 2, {'pyramid_level': 3, 'num_interest_points': 900}),  # W % 64 = 12, H % 16 = 10; level 1 is 125x166 (W % 4 != 0: tile kernel)
    (48, 256, 4, {'pyramid_level': 1, 'num_interest_points': 300}),   # runs of three bands
    (16, 512, 2, {'pyramid_level': 1, 'feature_width': 8, 'num_interest_points': 100}),   # one band per strip: every chunk pair is first and last
    (1080, 1920, 1, {}),                                              # 2040 + 510 + 136 bands over 148 CTAs, a single image
])
def test_stream_small_and_odd_shapes_equal_oracle(stream_everywhere, h, w, nimg, params):
    N = stream_everywhere
    from oracle import oracle as O
    from sfmfromscratch_b200 import extractor as S
    from sfmfromscratch_b200.synth import synth_image
    imgs = np.stack([synth_image(h, w, 40 + s) for s in range(nimg)])
    res, st = _kernels(N, lambda: S.extract_batch(imgs, params))
    assert st.get("k_harris_stream", (0, 0.0))[0] >= 1, f"the stream kernel did not run: {sorted(st)}"
    for b in range(nimg if h < 1000 else 1):
        o = O.ScaleRotInvSIFT(imgs[b], params)
        X, Y = o.detect_keypoints()
        assert_keypoints_equal(res[b][0], res[b][1], X, Y)
        g = S.ScaleRotInvSIFT(imgs[b], params)                        # (single-image call: levels, confidences, level coordinates)
        assert np.array_equal(g.confidences.view(np.uint32), o.confidences.view(np.uint32))
        ex = DescriptorExplainer(imgs[b], params, g.levels, g.level_x, g.level_y)
        assert assert_descriptors_close(res[b][2], o.extract_descriptors(), ex) == 0


def test_stream_and_tile_kernels_agree_bitwise(stream_everywhere):
    """The same batch through both Harris kernels: identical keypoints, descriptors bit for bit."""
    N = stream_everywhere
    from sfmfromscratch_b200 import extractor as S
    from sfmfromscratch_b200.synth import synth_image
    imgs = np.stack([synth_image(300, 404, 70 + s) for s in range(4)])
    a, sa = _kernels(N, lambda: S.extract_batch(imgs, {}))
    N.set_option(N.SFM_OPT_HARRIS_STREAM_MIN_BANDS, 1 << 20)
    b, sb = _kernels(N, lambda: S.extract_batch(imgs, {}))
    assert "k_harris_stream" in sa and "k_harris_stream" not in sb
    for ra, rb in zip(a, b):
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1]) and np.array_equal(ra[2], rb[2])


def test_set_option_rejects_what_it_does_not_know(stream_everywhere):
    N = stream_everywhere
    with pytest.raises(N.SfmError):
        N.set_option(999, 1)
    with pytest.raises(N.SfmError):
        N.set_option(N.SFM_OPT_HARRIS_STREAM_MIN_BANDS, -1)
