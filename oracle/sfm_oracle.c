/*
 * sfm_oracle.c -- CPU restatement of the SfmFromScratch feature hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product
 * path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product
 * (sfmfromscratch_b200/csrc) never links or calls it and has no CPU fallback.
 *
 * Parity status: the reference ships no tests or golden vectors (SURVEY.md
 * section 4), so this restatement is pinned against the reference itself run
 * in the build container (tests/golden/make_golden.py imports
 * /root/reference and commits its outputs; tests/test_oracle_vs_golden.py and
 * tests/test_oracle_vs_reference.py compare).  The third-party arithmetic the
 * reference delegates to (opencv-python filter2D / resize, numpy sum) is
 * restated here from its observed behaviour:
 *   - cv2.filter2D(float32, BORDER_CONSTANT): per pixel, acc = 0, then for
 *     every non-zero tap in row-major order acc = fmaf(k, p, acc);
 *   - cv2.resize INTER_LINEAR at an exact 2x reduction takes the INTER_AREA
 *     fast path: ((p00 + p01) + (p10 + p11)) * 0.25f; any other size goes to
 *     IPP's bilinear (three fmaf lerps, see orc_resize_bilinear);
 *   - np.sum(axis=-1) over 128 contiguous float32: 8 strided accumulators and
 *     a fixed pairwise tree.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference root).  Compile with -ffp-contract=off: every rounding below
 * is explicit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline float px(const float *img, int H, int W, int y, int x) {
    return (y < 0 || y >= H || x < 0 || x >= W) ? 0.0f : img[(size_t)y * W + x];
}

/* FeatureExtractor/SIFT/NaiveSIFT.py:201-213 (_compute_image_gradients):
 * cv2.filter2D with the 3x3 Sobel kernels of NaiveSIFT.py:23-31, correlation,
 * zero border.  Taps with coefficient 0 are skipped; the remaining six are
 * accumulated in row-major order starting from 0. */
ORC_API void orc_sobel(const float *img, int H, int W, float *Ix, float *Iy) {
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            float a = px(img, H, W, y - 1, x - 1), b = px(img, H, W, y - 1, x), c = px(img, H, W, y - 1, x + 1);
            float d = px(img, H, W, y, x - 1), f = px(img, H, W, y, x + 1);
            float g = px(img, H, W, y + 1, x - 1), h = px(img, H, W, y + 1, x), i = px(img, H, W, y + 1, x + 1);
            float sx = 0.0f;
            sx = fmaf(-1.0f, a, sx); sx = fmaf(1.0f, c, sx);
            sx = fmaf(-2.0f, d, sx); sx = fmaf(2.0f, f, sx);
            sx = fmaf(-1.0f, g, sx); sx = fmaf(1.0f, i, sx);
            float sy = 0.0f;
            sy = fmaf(-1.0f, a, sy); sy = fmaf(-2.0f, b, sy); sy = fmaf(-1.0f, c, sy);
            sy = fmaf(1.0f, g, sy);  sy = fmaf(2.0f, h, sy);  sy = fmaf(1.0f, i, sy);
            Ix[(size_t)y * W + x] = sx;
            Iy[(size_t)y * W + x] = sy;
        }
    }
}

/* cv2.filter2D(float32 plane, g x g float32 kernel, BORDER_CONSTANT) as used at
 * NaiveSIFT.py:67-69: row-major fmaf chain over all taps (zero taps of the
 * kernel are skipped, as OpenCV's filter engine drops them). */
ORC_API void orc_filter2d(const float *src, int H, int W, const float *k, int g, float *dst) {
    int r = g / 2;
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            float acc = 0.0f;
            for (int dy = 0; dy < g; ++dy) {
                for (int dx = 0; dx < g; ++dx) {
                    float kv = k[dy * g + dx];
                    if (kv == 0.0f) continue;
                    acc = fmaf(kv, px(src, H, W, y + dy - r, x + dx - r), acc);
                }
            }
            dst[(size_t)y * W + x] = acc;
        }
    }
}

/* NaiveSIFT.py:60-74: second moments and the Harris response
 * R = (Sxx*Syy - Sxy**2) - float32(alpha) * (Sxx+Syy)**2, every operation
 * rounded to float32 separately.  Returns 0 on success, -1 on allocation
 * failure.  gk is the g x g kernel already cast to float32. */
ORC_API int orc_harris_response(const float *img, int H, int W, const float *gk, int g, float alpha, float *R) {
    size_t n = (size_t)H * W;
    float *buf = (float *)malloc(n * 8 * sizeof(float));
    if (!buf) return -1;
    float *Ix = buf, *Iy = buf + n, *xx = buf + 2 * n, *yy = buf + 3 * n, *xy = buf + 4 * n;
    float *sxx = buf + 5 * n, *syy = buf + 6 * n, *sxy = buf + 7 * n;
    orc_sobel(img, H, W, Ix, Iy);
    for (size_t i = 0; i < n; ++i) {
        xx[i] = Ix[i] * Ix[i];
        yy[i] = Iy[i] * Iy[i];
        xy[i] = Ix[i] * Iy[i];
    }
    orc_filter2d(xx, H, W, gk, g, sxx);
    orc_filter2d(xy, H, W, gk, g, sxy);
    orc_filter2d(yy, H, W, gk, g, syy);
    for (size_t i = 0; i < n; ++i) {
        float det = sxx[i] * syy[i];
        float sq = sxy[i] * sxy[i];
        det = det - sq;
        float tr = sxx[i] + syy[i];
        float tr2 = tr * tr;
        float at = alpha * tr2;
        R[i] = det - at;
    }
    free(buf);
    return 0;
}

/* NaiveSIFT.py:77-88: clipped (2*half+1)^2 window maximum (the reference's
 * Python double loop). */
ORC_API void orc_maxpool(const float *R, int H, int W, int half, float *out) {
    for (int y = 0; y < H; ++y) {
        int y0 = y - half < 0 ? 0 : y - half, y1 = y + half + 1 > H ? H : y + half + 1;
        for (int x = 0; x < W; ++x) {
            int x0 = x - half < 0 ? 0 : x - half, x1 = x + half + 1 > W ? W : x + half + 1;
            float m = R[(size_t)y0 * W + x0];
            for (int yy = y0; yy < y1; ++yy)
                for (int xx = x0; xx < x1; ++xx) {
                    float v = R[(size_t)yy * W + xx];
                    if (v > m) m = v;
                }
            out[(size_t)y * W + x] = m;
        }
    }
}

/* FeatureExtractor/SIFT/ScaleRotInvSIFT.py:109-115 at scale factor 2 with even
 * source dimensions: cv2.resize(INTER_LINEAR) dispatches to the INTER_AREA 2x2
 * fast path. */
ORC_API void orc_resize_half(const float *src, int H, int W, float *dst) {
    int h = H / 2, w = W / 2;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const float *p = src + (size_t)(2 * y) * W + 2 * x;
            float top = p[0] + p[1];
            float bot = p[W] + p[W + 1];
            float s = top + bot;
            dst[(size_t)y * w + x] = s * 0.25f;
        }
}

/* ScaleRotInvSIFT.py:109-115, general factor: cv2.resize(INTER_LINEAR) on
 * float32.  In the build container opencv-python 4.13 routes this call to
 * IPP (ippiResizeLinear_32f); its observed arithmetic, reproduced bit-exactly
 * here, is: source coordinate f = (d + 0.5) * (src/dst) - 0.5 in double,
 * s = floor(f), fraction (double) cast to float, clamped at both borders, then
 *   t = fmaf(fx, p01 - p00, p00); b = fmaf(fx, p11 - p10, p10);
 *   out = fmaf(fy, b - t, t). */
ORC_API void orc_resize_bilinear(const float *src, int H, int W, float *dst, int h, int w) {
    double sxs = (double)W / w, sys = (double)H / h;
    for (int dy = 0; dy < h; ++dy) {
        double fyd = (dy + 0.5) * sys - 0.5;
        int sy = (int)floor(fyd);
        double fyr = fyd - sy;
        if (sy < 0) { sy = 0; fyr = 0; }
        if (sy >= H - 1) { sy = H - 1; fyr = 0; }
        int sy1 = sy + 1 < H ? sy + 1 : H - 1;
        float fy = (float)fyr;
        for (int dx = 0; dx < w; ++dx) {
            double fxd = (dx + 0.5) * sxs - 0.5;
            int sx = (int)floor(fxd);
            double fxr = fxd - sx;
            if (sx < 0) { sx = 0; fxr = 0; }
            if (sx >= W - 1) { sx = W - 1; fxr = 0; }
            int sx1 = sx + 1 < W ? sx + 1 : W - 1;
            float fx = (float)fxr;
            float p00 = src[(size_t)sy * W + sx], p01 = src[(size_t)sy * W + sx1];
            float p10 = src[(size_t)sy1 * W + sx], p11 = src[(size_t)sy1 * W + sx1];
            float t = fmaf(fx, p01 - p00, p00);
            float b = fmaf(fx, p11 - p10, p10);
            dst[(size_t)dy * w + dx] = fmaf(fy, b - t, t);
        }
    }
}

/* FeatureMatcher/NNRatioFeatureMatcher.py:31-34: one entry of
 * dists = sqrt(sum((a-b)**2, axis=2)) for D == 128 ... any D.  numpy's
 * pairwise_sum over a contiguous float32 run of n <= 128: 8 strided
 * accumulators, tree ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail; n < 8
 * is a plain loop; n > 128 recurses on halves (n2 = n/2 rounded down to a
 * multiple of 8). */
static float pairwise_sq(const float *a, const float *b, int n) {
    if (n < 8) {
        float res = 0.f;  /* numpy starts from -0.0; x*x >= 0 so the value is identical */
        for (int i = 0; i < n; ++i) { float t = a[i] - b[i]; t = t * t; res = res + t; }
        return res;
    } else if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; ++j) { float t = a[j] - b[j]; r[j] = t * t; }
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) { float t = a[i + j] - b[i + j]; t = t * t; r[j] = r[j] + t; }
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) { float t = a[i] - b[i]; t = t * t; res = res + t; }
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        float lo = pairwise_sq(a, b, n2);
        float hi = pairwise_sq(a + n2, b + n2, n - n2);
        return lo + hi;
    }
}

ORC_API float orc_dist(const float *a, const float *b, int D) {
    return sqrtf(pairwise_sq(a, b, D));
}

/* NNRatioFeatureMatcher.py:31-51 for rows [r0, r1) of features1: the two
 * smallest distances per row (value semantics of argsort()[0], [1]: smallest
 * value, and second smallest value counting duplicates) and the index of the
 * smallest.  When the two smallest tie the reference's index is
 * implementation-defined (unstable argsort) but the ratio is 1, see
 * SURVEY.md section 8a; idx0 is then the lowest such column. */
ORC_API void orc_match_top2(const float *f1, int r0, int r1, const float *f2, int n2, int D,
                            int64_t *idx0, float *d0, float *d1) {
    for (int i = r0; i < r1; ++i) {
        float b0 = INFINITY, b1 = INFINITY;
        int64_t j0 = -1;
        const float *a = f1 + (size_t)i * D;
        for (int j = 0; j < n2; ++j) {
            float d = orc_dist(a, f2 + (size_t)j * D, D);
            if (d < b0) { b1 = b0; b0 = d; j0 = j; }
            else if (d < b1) { b1 = d; }
        }
        idx0[i - r0] = j0; d0[i - r0] = b0; d1[i - r0] = b1;
    }
}

/* Full float32 distance matrix (NNRatioFeatureMatcher.py:31-34), small cases. */
ORC_API void orc_dist_matrix(const float *f1, int n1, const float *f2, int n2, int D, float *out) {
    for (int i = 0; i < n1; ++i)
        for (int j = 0; j < n2; ++j)
            out[(size_t)i * n2 + j] = orc_dist(f1 + (size_t)i * D, f2 + (size_t)j * D, D);
}
