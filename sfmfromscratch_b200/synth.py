"""Seeded synthetic inputs for the feature hot path.

Everything here is plain element-wise numpy in float64 with a fixed operation
order, so the same seed yields the same bytes on every x86 host (no BLAS, no
OpenCV, no reduction whose order depends on the SIMD width).  Used by the
tests, ``bench.py`` and the golden-vector generator.
"""
from __future__ import annotations

import numpy as np


def _gauss_taps(sigma: float) -> np.ndarray:
    r = int(np.ceil(4.0 * sigma))
    t = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-(t * t) / (2.0 * sigma * sigma))
    s = 0.0
    for v in k:            # fixed-order sum
        s += float(v)
    return k / s


def _blur_axis(a: np.ndarray, taps: np.ndarray, axis: int) -> np.ndarray:
    r = len(taps) // 2
    pad = [(0, 0), (0, 0)]
    pad[axis] = (r, r)
    p = np.pad(a, pad, mode="reflect")
    out = np.zeros_like(a)
    n = a.shape[axis]
    for i, w in enumerate(taps):
        sl = [slice(None), slice(None)]
        sl[axis] = slice(i, i + n)
        out += w * p[tuple(sl)]
    return out


def synth_image(h: int, w: int, seed: int, sigma: float = 2.0) -> np.ndarray:
    """Uniform noise -> Gaussian blur (sigma 2) -> min-max normalise to [0,1].
    float32 (h, w), generic enough that Harris responses do not tie."""
    rng = np.random.default_rng(seed)
    a = rng.random((h, w), dtype=np.float64)
    taps = _gauss_taps(sigma)
    a = _blur_axis(_blur_axis(a, taps, 0), taps, 1)
    lo, hi = a.min(), a.max()
    return ((a - lo) / (hi - lo)).astype(np.float32)


def sequence_canvas(h: int, w: int, total: int, seed: int = 7, step: int = 24, sigma: float = 2.0) -> np.ndarray:
    """The blurred-noise canvas (float64, h x (w + step * (total - 1)), in [0, 1]) a `total`-frame synthetic
    video pans over; depends on (h, w, total, seed, step) only, so every rank computes the same one."""
    cw = w + step * (total - 1)
    rng = np.random.default_rng(seed)
    a = rng.random((h, cw), dtype=np.float64)
    taps = _gauss_taps(sigma)
    a = _blur_axis(_blur_axis(a, taps, 0), taps, 1)
    lo, hi = a.min(), a.max()
    return (a - lo) / (hi - lo)


def sequence_frame(canvas: np.ndarray, f: int, w: int, seed: int = 7, step: int = 24, noise: float = 0.005) -> np.ndarray:
    """Frame `f` of the video over `canvas`: the window starting `f * step` columns in, plus the frame's own
    N(0, noise) sensor noise (generator seeded by the frame index), clipped to [0, 1].  float32 (h, w)."""
    h = canvas.shape[0]
    r = np.random.default_rng(1000 * seed + 17 + f)
    return np.clip(canvas[:, f * step: f * step + w] + r.normal(0.0, noise, size=(h, w)), 0.0, 1.0).astype(np.float32)


def frame_sequence(h: int, w: int, first: int, count: int, total: int, seed: int = 7, step: int = 24,
                   noise: float = 0.005, threads: int = 1) -> np.ndarray:
    """Frames `first .. first + count - 1` of a `total`-frame synthetic video: a camera panning over one
    blurred-noise canvas by `step` pixels per frame.  Consecutive frames overlap by (w - step) columns, so the
    consecutive-pair matching the reference does (Runner.py:183-191) finds true correspondences, as it does
    on real footage.  float32 (count, h, w)."""
    canvas = sequence_canvas(h, w, total, seed, step)
    out = np.empty((count, h, w), np.float32)

    def one(k):
        out[k] = sequence_frame(canvas, first + k, w, seed, step, noise)
    if threads > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, range(count)))
    else:
        for k in range(count):
            one(k)
    return out


def second_view(img: np.ndarray, seed: int, noise: float = 0.005) -> np.ndarray:
    """Affine warp [[1, .02, 3.5], [-.02, 1, -2.25]] (bilinear, reflect border)
    plus N(0, noise) -- the second camera of the synthetic two-view pair."""
    h, w = img.shape
    a = img.astype(np.float64)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    sx = xx + 0.02 * yy + 3.5
    sy = -0.02 * xx + yy - 2.25

    def refl(v, n):
        v = np.abs(v)
        period = 2.0 * (n - 1)
        v = np.mod(v, period)
        return np.where(v > n - 1, period - v, v)

    sx, sy = refl(sx, w), refl(sy, h)
    x0 = np.minimum(np.floor(sx).astype(np.int64), w - 2)
    y0 = np.minimum(np.floor(sy).astype(np.int64), h - 2)
    fx, fy = sx - x0, sy - y0
    top = a[y0, x0] * (1 - fx) + a[y0, x0 + 1] * fx
    bot = a[y0 + 1, x0] * (1 - fx) + a[y0 + 1, x0 + 1] * fx
    out = top * (1 - fy) + bot * fy
    rng = np.random.default_rng(seed)
    out = out + rng.normal(0.0, noise, size=out.shape)
    return np.clip(out, 0.0, 1.0).astype(np.float32)


def _rootsift_like(h: np.ndarray) -> np.ndarray:
    """sqrt(l2normalise(h)) row-wise in float64 with a fixed-order norm."""
    sq = h * h
    s = np.zeros(h.shape[0], np.float64)
    for j in range(h.shape[1]):
        s += sq[:, j]
    return np.sqrt(h / np.sqrt(s)[:, None])


def synth_descriptor_base(n: int, seed: int = 12345, dim: int = 128) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.gamma(0.5, 1.0, size=(n, dim))


def synth_descriptors(n: int, image_id: int, base: np.ndarray | None = None,
                      planted: float = 0.5, noise: float = 0.05, dim: int = 128) -> np.ndarray:
    """Config-5 style descriptors: a `planted` fraction of rows are noisy copies
    of a shared base set (so true matches exist between images), the rest are
    fresh Gamma(0.5) rows; RootSIFT-shaped (non-negative, sqrt of an
    L2-normalised histogram).  float32 (n, dim)."""
    if base is None:
        base = synth_descriptor_base(n, dim=dim)
    rng = np.random.default_rng(1000003 * (image_id + 1))
    h = rng.gamma(0.5, 1.0, size=(n, dim))
    npl = int(n * planted)
    rows = rng.permutation(n)[:npl]
    src = rng.permutation(base.shape[0])[:npl]
    h[rows] = base[src] + np.abs(rng.normal(0.0, noise, size=(npl, dim)))
    return _rootsift_like(h).astype(np.float32)


def two_view_correspondences(n: int, seed: int, outlier_frac: float = 0.3, width: int = 960, height: int = 540,
                             noise_px: float = 0.4):
    """Seeded integer-pixel correspondences of a two-view scene for the RANSAC stage (SURVEY.md
    section 8f row 2): `n` random 3-D points seen by two cameras with the same intrinsics K (a
    small rotation about y plus a sideways translation), Gaussian pixel noise in the second view,
    a fraction replaced by uniformly random outliers, coordinates rounded to int64 as the
    extractor's keypoints are.  Returns (p1 (n,2) int64, p2 (n,2) int64, K (3,3) float64)."""
    rng = np.random.default_rng(seed)
    K = np.array([[800.0, 0, width / 2], [0, 800.0, height / 2], [0, 0, 1]])
    X = np.column_stack([rng.uniform(-4, 4, n), rng.uniform(-2.5, 2.5, n), rng.uniform(4, 12, n)])
    ang = 0.1
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([-1.0, 0.05, 0.1])
    x1 = (K @ X.T).T
    x1 = x1[:, :2] / x1[:, 2:]
    x2 = (K @ (R @ X.T + t[:, None])).T
    x2 = x2[:, :2] / x2[:, 2:]
    x2 = x2 + rng.normal(0, noise_px, x2.shape)
    m = rng.random(n) < outlier_frac
    x2[m] = np.column_stack([rng.uniform(0, width, int(m.sum())), rng.uniform(0, height, int(m.sum()))])
    return np.round(x1).astype(np.int64), np.round(x2).astype(np.int64), K
