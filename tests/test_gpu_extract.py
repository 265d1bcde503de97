"""Parity of the CUDA extraction path (through the C ABI) with the oracle and
the reference's golden outputs.  Needs a B200: pytest -m gpu."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from parity import DescriptorExplainer, assert_descriptors_close, assert_keypoints_equal  # noqa: E402


def explainer(e, image, params, pyramid=True):
    """Per-keypoint exemption prover for the extractor object `e` (its level coordinates are the reference's:
    the tests assert the keypoints equal first)."""
    if pyramid:
        return DescriptorExplainer(image, params, e.levels, e.level_x, e.level_y, pyramid=True)
    X, Y = e.detect_keypoints()
    return DescriptorExplainer(image, params, None, X, Y, pyramid=False)


def _mods():
    from oracle import oracle as O
    import sfmfromscratch_b200 as S
    from sfmfromscratch_b200 import extractor as X
    return O, S, X


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("h,w,params", [
    (96, 128, {}), (101, 135, {}), (480, 640, {}), (64, 64, {}), (65, 63, {}), (7, 300, {}),
    (70, 50, {'gaussian_size': 5, 'sigma': 2.0}), (65, 200, {'gaussian_size': 3, 'sigma': 1.0}),
    (130, 131, {'gaussian_size': 9, 'sigma': 3.0, 'alpha': 0.06}), (90, 77, {'gaussian_size': 11, 'sigma': 4.0}),
    (40, 40, {'gaussian_size': 1, 'sigma': 1.0}),
])
def test_harris_response_bit_exact(h, w, params):
    """R map: bit-exact (the 49-tap fmaf chain is order-preserving)."""
    O, S, X = _mods()
    from sfmfromscratch_b200.synth import synth_image
    img = synth_image(h, w, h + w)
    R = X.harris_response(img, params)
    Ro = O.harris_response(img, params.get('gaussian_size', 7), params.get('sigma', 5), params.get('alpha', 0.05))
    assert np.array_equal(R.view(np.uint32), Ro.view(np.uint32))


def test_harris_response_golden(golden_dir):
    _, _, X = _mods()
    g = load(golden_dir, "two_view_96x128.npz")
    assert np.array_equal(X.harris_response(g["img1"], {}).view(np.uint32), g["R1"].view(np.uint32))


@pytest.mark.parametrize("name", ["two_view_96x128.npz", "two_view_240x320.npz"])
def test_two_view_golden(golden_dir, name):
    _, S, _ = _mods()
    g = load(golden_dir, name)
    params = {'num_interest_points': 600} if "96x128" in name else {}
    for i in (1, 2):
        e = S.ScaleRotInvSIFT(g[f"img{i}"], params)
        X, Y = e.detect_keypoints()
        assert_keypoints_equal(X, Y, g[f"X{i}"], g[f"Y{i}"])
        D = e.extract_descriptors()
        assert D.dtype == np.float32 and D.shape == g[f"D{i}"].shape
        assert assert_descriptors_close(D, g[f"D{i}"], explainer(e, g[f"img{i}"], params)) == 0


def test_mainpy_params_golden(golden_dir):
    """main.py:19-28 parameters: ksize 3, feature_width 18, 3 levels, factor 1.1 (bilinear pyramid)."""
    _, S, _ = _mods()
    g = load(golden_dir, "srs_mainpy_120x160.npz")
    prm = json.loads(str(g["params"]))
    e = S.ScaleRotInvSIFT(g["img"], prm)
    assert_keypoints_equal(*e.detect_keypoints(), g["X"], g["Y"])
    assert assert_descriptors_close(e.extract_descriptors(), g["D"], explainer(e, g["img"], prm)) == 0


def test_odd_size_golden(golden_dir):
    _, S, _ = _mods()
    g = load(golden_dir, "srs_odd_101x135.npz")
    e = S.ScaleRotInvSIFT(g["img"], {'num_interest_points': 400})
    assert_keypoints_equal(*e.detect_keypoints(), g["X"], g["Y"])
    assert assert_descriptors_close(e.extract_descriptors(), g["D"], explainer(e, g["img"], {'num_interest_points': 400})) == 0


def test_naive_sift_golden(golden_dir):
    _, S, _ = _mods()
    g = load(golden_dir, "naive_96x128.npz")
    e = S.NaiveSIFT(g["img"], {'num_interest_points': 300})
    with pytest.raises(RuntimeError):
        e.extract_descriptors()
    X, Y = e.detect_keypoints()
    assert_keypoints_equal(X, Y, g["X"], g["Y"])
    assert np.array_equal(e.confidences, g["conf"])           # responses are bit-exact
    assert assert_descriptors_close(e.extract_descriptors(), g["D"], explainer(e, g["img"], {'num_interest_points': 300}, pyramid=False)) == 0


@pytest.mark.parametrize("h,w,seed,params", [
    (480, 640, 0, {}),
    (480, 640, 1, {'num_interest_points': 6000}),
    (333, 517, 2, {'pyramid_level': 3, 'pyramid_scale_factor': 1.5, 'feature_width': 12}),
    (300, 400, 3, {'ksize': 3, 'sigma': 6, 'feature_width': 18, 'pyramid_level': 3, 'pyramid_scale_factor': 1.1}),
    (256, 256, 4, {'ksize': 9, 'gaussian_size': 5, 'sigma': 1.5, 'feature_width': 8, 'pyramid_level': 2}),
    (200, 300, 5, {'pyramid_level': 1}),
    (48, 64, 6, {'pyramid_level': 5}),
    (150, 210, 7, {'ksize': 1, 'pyramid_level': 2}),              # 1x1 NMS window: every pixel at or above the median
    (222, 318, 8, {'ksize': 5, 'gaussian_size': 3}),
    (190, 254, 9, {'ksize': 11, 'gaussian_size': 9, 'num_interest_points': 900}),
])
def test_scale_rot_inv_vs_oracle(h, w, seed, params):
    O, S, _ = _mods()
    from sfmfromscratch_b200.synth import synth_image
    img = synth_image(h, w, seed)
    g, o = S.ScaleRotInvSIFT(img, params), O.ScaleRotInvSIFT(img, params)
    assert_keypoints_equal(*g.detect_keypoints(), *o.detect_keypoints())
    assert np.array_equal(g.levels, o.levels)
    assert np.array_equal(g.confidences.view(np.uint32), o.confidences.view(np.uint32))
    # generic seeded image: every descriptor within tolerance, no keypoint needs the bin-edge exemption
    assert assert_descriptors_close(g.extract_descriptors(), o.extract_descriptors(), explainer(g, img, params)) == 0


def test_batch_equals_single():
    """A batch call returns, per image, exactly what single-image calls return."""
    _, S, _ = _mods()
    from sfmfromscratch_b200.synth import synth_image
    imgs = np.stack([synth_image(240, 320, s) for s in (10, 11, 12)])
    res = S.extract_batch(imgs, {})
    for b in range(3):
        e = S.ScaleRotInvSIFT(imgs[b], {})
        X, Y = e.detect_keypoints()
        assert np.array_equal(res[b][0], X) and np.array_equal(res[b][1], Y)
        assert np.array_equal(res[b][2], e.extract_descriptors())


def test_plateau_image_overflow_retry():
    """Zero image: every pixel ties (R == 0 == median): the candidate buffer
    overflows, the host retries with one slot per pixel; the top-k among equal
    responses is the k lowest pixel indices (canonical tie order)."""
    O, S, _ = _mods()
    z = np.zeros((64, 80), np.float32)
    g = S.NaiveSIFT(z, {'num_interest_points': 100})
    X, Y = g.detect_keypoints()
    o = O.NaiveSIFT(z, {'num_interest_points': 100})
    assert_keypoints_equal(X, Y, *o.detect_keypoints())      # first 100 pixels are on the border: all filtered
    img = np.zeros((64, 80), np.float32)
    img[20:44, 30:60] = 0.5
    g = S.NaiveSIFT(img, {'num_interest_points': 4000})
    o = O.NaiveSIFT(img, {'num_interest_points': 4000})
    gx, gy = g.detect_keypoints()
    ox, oy = o.detect_keypoints()
    assert_keypoints_equal(gx, gy, ox, oy)
    # Descriptors are NOT compared here: an axis-aligned step edge produces gradients at exact
    # multiples of pi/4, which sit exactly on the 8-bin histogram edges, and numpy's float32
    # arctan2(1, 1) is 1 ulp below the correctly rounded pi/4 this path computes -- the bin such a
    # sample lands in is decided by that ulp ("orientation arithmetic").
    D = np.atleast_2d(g.extract_descriptors())
    assert D.shape == (len(gx), 128) and np.isfinite(D).all() and (D >= 0).all() and (D <= 1).all()


def test_degenerate_and_errors():
    _, S, _ = _mods()
    from sfmfromscratch_b200.synth import synth_image
    e = S.ScaleRotInvSIFT(np.zeros((40, 40), np.float32), {'num_interest_points': 40})
    X, Y = e.detect_keypoints()
    assert len(X) == len(e.extract_descriptors())
    with pytest.raises(AssertionError):
        S.ScaleRotInvSIFT(np.zeros((4, 40, 40), np.float32), {})          # 'Image must be grayscale'
    with pytest.raises(Exception):
        S.ScaleRotInvSIFT(synth_image(64, 64, 0), {'gaussian_size': 13})   # unsupported window
    with pytest.raises(Exception):
        S.ScaleRotInvSIFT(synth_image(8, 8, 0), {'pyramid_level': 6})      # empty pyramid level


def test_full_size_properties_1080p():
    """configs[1] at full size through size-independent properties, plus the
    oracle on the same image (the C oracle needs about a second at 1080p)."""
    O, S, _ = _mods()
    from sfmfromscratch_b200.synth import synth_image
    img = synth_image(1080, 1920, 0)
    g = S.ScaleRotInvSIFT(img, {})
    X, Y = g.detect_keypoints()
    D = g.extract_descriptors()
    assert X.dtype == np.int64 and D.shape == (len(X), 128) and len(X) <= 2500
    assert (X >= 0).all() and (X < 1920).all() and (Y >= 0).all() and (Y < 1080).all()
    for l in range(4):                                   # per level: response descending
        c = g.confidences[g.levels == l]
        assert np.all(np.diff(c) <= 0)
    assert np.all(np.diff(g.levels) >= 0)
    nz = D.any(axis=1)
    assert np.allclose((D[nz].astype(np.float64) ** 4).sum(axis=1), 1.0, atol=1e-5)   # sqrt of an L2-normalised vector
    g2 = S.ScaleRotInvSIFT(img, {})
    assert np.array_equal(g2.extract_descriptors(), D) and np.array_equal(g2.detect_keypoints()[0], X)   # deterministic
    o = O.ScaleRotInvSIFT(img, {})
    assert_keypoints_equal(X, Y, *o.detect_keypoints())
    assert assert_descriptors_close(D, o.extract_descriptors(), explainer(g, img, {})) == 0


def test_4k_many_keypoints_vs_oracle():
    """configs[2] extraction half: 3840x2160, num_interest_points 32000."""
    O, S, _ = _mods()
    from sfmfromscratch_b200.synth import synth_image
    img = synth_image(2160, 3840, 1)
    p = {'num_interest_points': 32000}
    g = S.ScaleRotInvSIFT(img, p)
    X, Y = g.detect_keypoints()
    assert len(X) > 15000
    R = O.harris_response(img)
    from sfmfromscratch_b200 import extractor as XX
    assert np.array_equal(XX.harris_response(img, {}).view(np.uint32), R.view(np.uint32))
    # oracle keypoints per level from the C window-max (descriptors are covered at the smaller sizes)
    pyr = O.build_pyramid(img, 4, 2)
    off = 0
    for l, im in enumerate(pyr):
        fw = max(int(16 / 2 ** l), 3)
        x, y, c = O.harris_interest_points(im, 8000, fw)
        n = len(x)
        assert np.array_equal(g.level_x[off:off + n], x) and np.array_equal(g.level_y[off:off + n], y)
        assert np.all(g.levels[off:off + n] == l)
        off += n
    assert off == len(X)


def test_run_host_pipeline_equals_plain_calls():
    """FeaturePipeline.run_host (chunked, 3 streams, incremental matching) returns exactly what the
    plain batch calls return."""
    import torch
    _, S, _ = _mods()
    from sfmfromscratch_b200 import pipeline as PL
    from sfmfromscratch_b200.synth import synth_image
    B, H, W, cap = 11, 120, 160, 2500
    imgs = np.stack([synth_image(H, W, 40 + s) for s in range(B)])
    host = torch.from_numpy(imgs).pin_memory()
    pairs = PL.consecutive_pairs(B)
    out = {'x': torch.zeros((B, cap), dtype=torch.int32).pin_memory(), 'y': torch.zeros((B, cap), dtype=torch.int32).pin_memory(),
           'desc': torch.zeros((B, cap, 128), dtype=torch.float32).pin_memory(), 'count': torch.zeros((B,), dtype=torch.int32).pin_memory(),
           'matches': torch.zeros((len(pairs), cap, 2), dtype=torch.int32).pin_memory(),
           'conf': torch.zeros((len(pairs), cap), dtype=torch.float32).pin_memory(),
           'mcount': torch.zeros((len(pairs),), dtype=torch.int32).pin_memory()}
    pipe = PL.FeaturePipeline({}, 0.8)
    order = pipe.run_host(host, pairs, out, chunk=4)
    ref = S.extract_batch(imgs, {})
    for b in range(B):
        n = int(out['count'][b])
        assert n == len(ref[b][0])
        assert np.array_equal(out['x'][b, :n].numpy(), ref[b][0]) and np.array_equal(out['y'][b, :n].numpy(), ref[b][1])
        assert np.array_equal(out['desc'][b, :n].numpy(), ref[b][2])
    assert sorted(map(tuple, order.tolist())) == sorted(map(tuple, pairs.tolist()))
    for k, (i, j) in enumerate(order.tolist()):
        m, c = S.NNRatioFeatureMatcher(0.8).match_features_ratio_test(ref[i][2], ref[j][2])
        n = int(out['mcount'][k])
        assert n == len(m)
        if n:
            assert np.array_equal(out['matches'][k, :n].numpy().astype(np.int64), m)
            assert np.array_equal(out['conf'][k, :n].numpy(), c)


def test_stream_host_equals_run_host():
    """FeaturePipeline.stream_host (no host wait between batches, two device slots recycled by events)
    delivers, for every batch of a sequence, exactly what run_host delivers for that batch alone."""
    import torch
    from sfmfromscratch_b200 import pipeline as PL
    from sfmfromscratch_b200.synth import synth_image
    B, H, W, cap = 6, 96, 128, 2500
    pairs = PL.consecutive_pairs(B)

    def outs():
        return {'x': torch.zeros((B, cap), dtype=torch.int32).pin_memory(), 'y': torch.zeros((B, cap), dtype=torch.int32).pin_memory(),
                'desc': torch.zeros((B, cap, 128), dtype=torch.float32).pin_memory(), 'count': torch.zeros((B,), dtype=torch.int32).pin_memory(),
                'matches': torch.zeros((len(pairs), cap, 2), dtype=torch.int32).pin_memory(),
                'conf': torch.zeros((len(pairs), cap), dtype=torch.float32).pin_memory(),
                'mcount': torch.zeros((len(pairs),), dtype=torch.int32).pin_memory()}
    batches = [torch.from_numpy(np.stack([synth_image(H, W, 100 * k + s) for s in range(B)])).pin_memory() for k in range(5)]
    pipe = PL.FeaturePipeline({}, 0.8)
    got = [outs() for _ in batches]
    for hb, o in zip(batches, got):
        pipe.stream_host(hb, pairs, o, chunk=4)
    assert pipe.drain()
    single = PL.FeaturePipeline({}, 0.8)
    for hb, o in zip(batches, got):
        ref = outs()
        single.run_host(hb, pairs, ref, chunk=4)
        assert np.array_equal(o['count'].numpy(), ref['count'].numpy()) and np.array_equal(o['mcount'].numpy(), ref['mcount'].numpy())
        for b in range(B):
            n = int(ref['count'][b])
            for k in ('x', 'y', 'desc'):
                assert np.array_equal(o[k][b, :n].numpy(), ref[k][b, :n].numpy())
        for k in range(len(pairs)):
            n = int(ref['mcount'][k])
            assert np.array_equal(o['matches'][k, :n].numpy(), ref['matches'][k, :n].numpy())
            assert np.array_equal(o['conf'][k, :n].numpy(), ref['conf'][k, :n].numpy())


def test_stream_host_reports_candidate_overflow():
    """A plateau image (all responses tie) overflows the density-sized candidate buffer: stream_host does not
    wait on the host, so the flag travels to pinned memory with the results and drain() reports it."""
    import torch
    from sfmfromscratch_b200 import pipeline as PL
    from sfmfromscratch_b200.synth import synth_image
    B, H, W, cap = 4, 96, 128, 2500
    pairs = PL.consecutive_pairs(B)
    good = np.stack([synth_image(H, W, s) for s in range(B)])
    bad = good.copy()
    bad[2] = 0.0
    out = {'x': torch.zeros((B, cap), dtype=torch.int32).pin_memory(), 'y': torch.zeros((B, cap), dtype=torch.int32).pin_memory(),
           'desc': torch.zeros((B, cap, 128), dtype=torch.float32).pin_memory(), 'count': torch.zeros((B,), dtype=torch.int32).pin_memory(),
           'matches': torch.zeros((len(pairs), cap, 2), dtype=torch.int32).pin_memory(),
           'conf': torch.zeros((len(pairs), cap), dtype=torch.float32).pin_memory(),
           'mcount': torch.zeros((len(pairs),), dtype=torch.int32).pin_memory()}
    pipe = PL.FeaturePipeline({}, 0.8)
    pipe.stream_host(torch.from_numpy(good).pin_memory(), pairs, out, chunk=2)
    assert pipe.drain()
    pipe.stream_host(torch.from_numpy(bad).pin_memory(), pairs, out, chunk=2)
    pipe.stream_host(torch.from_numpy(good).pin_memory(), pairs, out, chunk=2)
    assert not pipe.drain()                                     # the plateau batch is reported, even when retired early
    pipe.params.cand_full = 1
    pipe.stream_host(torch.from_numpy(bad).pin_memory(), pairs, out, chunk=2)
    assert pipe.drain()
