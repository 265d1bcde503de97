// ingest.cu -- image ingest in front of the feature path:
// Runner.py:33-46 = _load_image (:551-563) -> _PIL_resize (:481-493) -> _rgb2gray (:467-478).
//
// Input: the decoded 8-bit RGB image (PIL's decode stays on the host).  Output:
// the float32 grayscale image the extractor receives, bit for bit:
//   - the reference round-trips through float32/255 and back (`img *= 255`,
//     np.uint8): an identity for every 8-bit value, so the bytes go straight in;
//   - PIL.Image.resize default filter on RGB = BICUBIC: Pillow's two-pass
//     convolution (libImaging/Resample.c), horizontal then vertical, 22-bit fixed
//     point coefficients normalised per output coordinate, each pass rounded and
//     clipped to uint8.  Pure integer arithmetic: reproduced exactly;
//   - float32(u8) / 255, then R*0.299f + G*0.587f + B*0.114f with the
//     reference's operation order, every op rounded.
// The coefficient tables are computed on the host in double precision exactly
// as Pillow's precompute_coeffs / normalize_coeffs_8bpc do and copied into the
// caller's workspace.
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

double bicubic_filter(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

// Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box.
int precompute(int in_size, int out_size, std::vector<int32_t>& bounds, std::vector<int32_t>& kk) {
    double scale = (double)in_size / out_size, filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;
    const int ksize = (int)std::ceil(support) * 2 + 1;
    bounds.assign((size_t)out_size * 2, 0);
    kk.assign((size_t)out_size * ksize, 0);
    std::vector<double> k((size_t)ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        const double ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) k[x] /= ww;
            const double v = k[x] * (double)(1 << PRECISION_BITS);
            kk[(size_t)xx * ksize + x] = (k[x] < 0) ? (int32_t)(-0.5 + v) : (int32_t)(0.5 + v);
        }
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
    return ksize;
}

__device__ __forceinline__ uint32_t clip8(int v) {
    v >>= PRECISION_BITS;                  // arithmetic shift, as Pillow's lookup index
    return (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: rgb [B][H][W][3] -> tmp [B][H][ow][3]
__global__ void __launch_bounds__(256) k_ingest_h(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ tmp, int H, int W,
                                                  int ow, const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk,
                                                  int ksize) {
    const int xx = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int b = blockIdx.z;
    if (xx >= ow || y >= H) return;
    const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
    const int32_t* k = kk + (size_t)xx * ksize;
    const uint8_t* row = rgb + ((size_t)b * H + y) * W * 3 + (size_t)xmin * 3;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < n; ++x) {
        const int c = __ldg(k + x);
        s0 += (int)__ldg(row + 3 * x + 0) * c;
        s1 += (int)__ldg(row + 3 * x + 1) * c;
        s2 += (int)__ldg(row + 3 * x + 2) * c;
    }
    uint8_t* o = tmp + (((size_t)b * H + y) * ow + xx) * 3;
    o[0] = (uint8_t)clip8(s0); o[1] = (uint8_t)clip8(s1); o[2] = (uint8_t)clip8(s2);
}

// vertical pass + float32/255 + rgb2gray: tmp [B][H][ow][3] -> gray [B][oh][ow]
__global__ void __launch_bounds__(256) k_ingest_v_gray(const uint8_t* __restrict__ tmp, float* __restrict__ gray, int H, int ow,
                                                       int oh, const int32_t* __restrict__ bounds,
                                                       const int32_t* __restrict__ kk, int ksize) {
    const int xx = blockIdx.x * 64 + (threadIdx.x & 63);
    const int yy = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int b = blockIdx.z;
    if (xx >= ow || yy >= oh) return;
    const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
    const int32_t* k = kk + (size_t)yy * ksize;
    const uint8_t* col = tmp + (((size_t)b * H + ymin) * ow + xx) * 3;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int y = 0; y < n; ++y) {
        const int c = __ldg(k + y);
        const uint8_t* p = col + (size_t)y * ow * 3;
        s0 += (int)__ldg(p + 0) * c;
        s1 += (int)__ldg(p + 1) * c;
        s2 += (int)__ldg(p + 2) * c;
    }
    // _PIL_image_to_numpy_arr (:495-509): astype(float32) / 255; _rgb2gray (:476-478)
    const float r = __fdiv_rn((float)clip8(s0), 255.0f);
    const float g = __fdiv_rn((float)clip8(s1), 255.0f);
    const float bl = __fdiv_rn((float)clip8(s2), 255.0f);
    const float v = __fadd_rn(__fadd_rn(__fmul_rn(r, 0.299f), __fmul_rn(g, 0.587f)), __fmul_rn(bl, 0.114f));
    gray[((size_t)b * oh + yy) * ow + xx] = v;
}

struct IngestWs { size_t bx, kx, by, ky, tmp, total; int ksx, ksy; };

void layout(int B, int H, int W, int oh, int ow, IngestWs& ws) {
    auto ks = [](int in, int out) {
        double fs = (double)in / out;
        if (fs < 1.0) fs = 1.0;
        return (int)std::ceil(2.0 * fs) * 2 + 1;
    };
    ws.ksx = ks(W, ow); ws.ksy = ks(H, oh);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 256); return at; };
    ws.bx = take(sizeof(int32_t) * 2 * ow);
    ws.kx = take(sizeof(int32_t) * (size_t)ow * ws.ksx);
    ws.by = take(sizeof(int32_t) * 2 * oh);
    ws.ky = take(sizeof(int32_t) * (size_t)oh * ws.ksy);
    ws.tmp = take((size_t)B * H * ow * 3);
    ws.total = o;
}

}  // namespace

extern "C" {

size_t sfm_ingest_workspace_bytes(int B, int H, int W, int out_h, int out_w) {
    if (B < 1 || H < 1 || W < 1 || out_h < 1 || out_w < 1) return 0;
    IngestWs ws;
    layout(B, H, W, out_h, out_w, ws);
    return ws.total;
}

int sfm_ingest_rgb8(SfmCtx* ctx, void* stream, const uint8_t* rgb_dev, int B, int H, int W, int out_h, int out_w,
                    void* workspace_dev, size_t workspace_bytes, float* gray_out) {
    if (!ctx) return SFM_ERR_BAD_ARG;
    if (!rgb_dev || !workspace_dev || !gray_out) return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "NULL pointer");
    if (B < 1 || H < 1 || W < 1 || out_h < 1 || out_w < 1) return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "bad sizes");
    IngestWs ws;
    layout(B, H, W, out_h, out_w, ws);
    if (workspace_bytes < ws.total)
        return sfm_set_error(ctx, SFM_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, ws.total);
    std::vector<int32_t> bx, kx, by, ky;
    const int ksx = precompute(W, out_w, bx, kx), ksy = precompute(H, out_h, by, ky);
    if (ksx != ws.ksx || ksy != ws.ksy) return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "coefficient size mismatch");
    cudaStream_t st = (cudaStream_t)stream;
    char* w = (char*)workspace_dev;
    SFM_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    // pageable sources: cudaMemcpyAsync stages them before returning, the vectors may die afterwards
    SFM_CUDA_CHECK(ctx, cudaMemcpyAsync(w + ws.bx, bx.data(), bx.size() * 4, cudaMemcpyHostToDevice, st));
    SFM_CUDA_CHECK(ctx, cudaMemcpyAsync(w + ws.kx, kx.data(), kx.size() * 4, cudaMemcpyHostToDevice, st));
    SFM_CUDA_CHECK(ctx, cudaMemcpyAsync(w + ws.by, by.data(), by.size() * 4, cudaMemcpyHostToDevice, st));
    SFM_CUDA_CHECK(ctx, cudaMemcpyAsync(w + ws.ky, ky.data(), ky.size() * 4, cudaMemcpyHostToDevice, st));
    uint8_t* tmp = (uint8_t*)(w + ws.tmp);
    SFM_LAUNCH(ctx, st, "k_ingest_h",
               k_ingest_h<<<dim3(ceil_div(out_w, 64), ceil_div(H, 4), B), 256, 0, st>>>(
                   rgb_dev, tmp, H, W, out_w, (const int32_t*)(w + ws.bx), (const int32_t*)(w + ws.kx), ksx));
    SFM_LAUNCH(ctx, st, "k_ingest_v_gray",
               k_ingest_v_gray<<<dim3(ceil_div(out_w, 64), ceil_div(out_h, 4), B), 256, 0, st>>>(
                   tmp, gray_out, H, out_w, out_h, (const int32_t*)(w + ws.by), (const int32_t*)(w + ws.ky), ksy));
    return SFM_OK;
}

}  // extern "C"
