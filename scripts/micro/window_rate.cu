// Microbenchmark: the 49-tap window stage of k_harris alone (copied from csrc/extract.cu) on a static product tile,
// to see how many warps per scheduler it takes to fill the FMA pipe.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define SFM_MAX_GAUSS 11
#define SFM_HIST1_BINS 4096
#define SFM_GW_PITCH 16
struct __align__(16) GaussWeights {
    float w[SFM_MAX_GAUSS * SFM_GW_PITCH];
    // wp[jj][dx] = (w[jj][dx], w[jj-1][dx]) for jj = 1..G-1: the packed (upper row, lower row) weight
    // pair of product row jj, consumed as one 64-bit uniform operand by the FFMA2 variant
    float2 wp[SFM_MAX_GAUSS * SFM_GW_PITCH];
};
constexpr int HT = 64;            // output tile width

// TH: output tile height; a thread owns 8 consecutive pixels on 2 adjacent rows.
template <int G, int TH> struct HarrisCfg {
    static constexpr int THREADS = 8 * (TH / 2);
    static constexpr int R = G / 2;
    static constexpr int RA = (R + 1 + 3) & ~3;        // image tile starts RA columns left of the output tile (16-byte aligned)
    static constexpr int OFF = RA - (R + 1);           // product column c reads image tile columns c+OFF .. c+OFF+2
    static constexpr int PW = HT + 2 * R;              // product tile width
    static constexpr int PH = TH + 2 * R;
    static constexpr int NV = 8 + 2 * R;               // product values a thread needs per row
    static constexpr int NCH = (NV + 3) / 4;           // ... in 16-byte chunks
    static constexpr int PCH = (14 + NCH + 1) & ~1;    // chunks per product row (even: XOR swizzle stays in range)
    static constexpr int PPITCH = PCH * 4;             // floats
    static constexpr int IPITCH = (PCH * 4 + OFF + 2 + 3) & ~3;   // image tile row pitch (strips may overrun into padding)
    static constexpr int IH = TH + 2 * R + 2;
    static constexpr int IMG_WORDS = IPITCH * IH;
    static constexpr int PROD_WORDS = 3 * PH * PPITCH;
    static constexpr size_t smem_bytes = sizeof(float) * ((size_t)IMG_WORDS + PROD_WORDS);
    static_assert(PROD_WORDS >= SFM_HIST1_BINS, "histogram aliases the product planes");
};

// packed float32 pair arithmetic (Blackwell FFMA2): both lanes IEEE round-to-nearest
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
    unsigned long long d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// 3. G x G window sums as row-major fmaf chains (what cv2.filter2D does), then
//    R = (Sxx*Syy - Sxy^2) - alpha * (Sxx+Syy)^2 with every op rounded
//    (NaiveSIFT.py:71-74).  r[q][p]: row 2*ty+q, pixel 8*tx+p of the tile.
template <int G, int TH, bool F2>
__device__ __forceinline__ void harris_window(const float* s_prod, const GaussWeights& gw, float alpha, float (&r)[2][8],
                                              int tid = threadIdx.x) {
    using C = HarrisCfg<G, TH>;
    const int tx = tid & 7, ty = tid >> 3;
    float S[3][2][8];
    auto load_row = [&](const float* row, float (&v)[4 * C::NCH]) {
#pragma unroll
        for (int j = 0; j < C::NCH; ++j) {
            const int c = 2 * tx + j;
            const float4 q4 = *reinterpret_cast<const float4*>(row + (c ^ ((c >> 3) & 1)) * 4);
            v[4 * j + 0] = q4.x; v[4 * j + 1] = q4.y; v[4 * j + 2] = q4.z; v[4 * j + 3] = q4.w;
        }
    };
    // Product row jj feeds tap row dy = jj of the upper output row and dy = jj - 1 of the lower one.
    // Rows 0 and G are peeled (one output row each); rows 1..G-1 run as a ROLLED loop with the two
    // weight rows fetched from the constant bank by index -- fully unrolled, the 3 x 2 x 49 x 8 FMAs
    // are ~58 KB of code and the kernel stalls on instruction fetch.
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
        const float* plane = s_prod + pl * C::PH * C::PPITCH + 2 * ty * C::PPITCH;
        float acc[2][8];
        float v[4 * C::NCH];
#pragma unroll
        for (int p = 0; p < 8; ++p) { acc[0][p] = 0.0f; acc[1][p] = 0.0f; }
        load_row(plane, v);
#pragma unroll
        for (int dx = 0; dx < G; ++dx)
#pragma unroll
            for (int p = 0; p < 8; ++p) acc[0][p] = __fmaf_rn(gw.w[dx], v[p + dx], acc[0][p]);   // tap row 0
        if constexpr (F2) {
            // both output rows in one packed FFMA2 per tap: (upper, lower) accumulators, the product
            // value broadcast to both lanes, the weight pair as one 64-bit constant operand
            unsigned long long acc2[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) acc2[p] = f2_pack(acc[0][p], 0.0f);
#pragma unroll 1
            for (int jj = 1; jj < G; ++jj) {
                load_row(plane + jj * C::PPITCH, v);
                const float2* wp = gw.wp + jj * SFM_GW_PITCH;
#pragma unroll
                for (int dx = 0; dx < G; ++dx) {
                    const unsigned long long ww = f2_pack(wp[dx].x, wp[dx].y);
#pragma unroll
                    for (int p = 0; p < 8; ++p) acc2[p] = f2_fma(f2_pack(v[p + dx], v[p + dx]), ww, acc2[p]);
                }
            }
#pragma unroll
            for (int p = 0; p < 8; ++p) f2_unpack(acc2[p], acc[0][p], acc[1][p]);
        } else {
#pragma unroll 1
            for (int jj = 1; jj < G; ++jj) {
                load_row(plane + jj * C::PPITCH, v);
                const float* w0 = gw.w + jj * SFM_GW_PITCH;
                const float* w1 = w0 - SFM_GW_PITCH;
#pragma unroll
                for (int dx = 0; dx < G; ++dx) {
                    const float a0 = w0[dx], a1 = w1[dx];
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        acc[0][p] = __fmaf_rn(a0, v[p + dx], acc[0][p]);
                        acc[1][p] = __fmaf_rn(a1, v[p + dx], acc[1][p]);
                    }
                }
            }
        }
        load_row(plane + G * C::PPITCH, v);
#pragma unroll
        for (int dx = 0; dx < G; ++dx)
#pragma unroll
            for (int p = 0; p < 8; ++p) acc[1][p] = __fmaf_rn(gw.w[(G - 1) * SFM_GW_PITCH + dx], v[p + dx], acc[1][p]);
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int p = 0; p < 8; ++p) S[pl][q][p] = acc[q][p];
    }
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const float sxx = S[0][q][p], sxy = S[1][q][p], syy = S[2][q][p];
            const float det = __fsub_rn(__fmul_rn(sxx, syy), __fmul_rn(sxy, sxy));
            const float tr = __fadd_rn(sxx, syy);
            r[q][p] = __fsub_rn(det, __fmul_rn(alpha, __fmul_rn(tr, tr)));
        }
}

// 3b. The same arithmetic with the tap ROWS unrolled and the three PLANES rolled: every weight is then a
//     compile-time offset into the constant bank (a direct operand of FFMA / FFMA2 -- no indexed constant
//     loads into vector registers, two fewer register-file reads per FFMA2), and the body stays ~8 KB.
template <int G, int TH>
__device__ __forceinline__ void harris_window_rows(const float* s_prod, const GaussWeights& gw, float alpha, float (&r)[2][8],
                                                   int tid = threadIdx.x) {
    using C = HarrisCfg<G, TH>;
    const int tx = tid & 7, ty = tid >> 3;
    int coff[C::NCH];
#pragma unroll
    for (int j = 0; j < C::NCH; ++j) { const int c = 2 * tx + j; coff[j] = (c ^ ((c >> 3) & 1)) * 4; }
    auto load_row = [&](const float* row, float (&v)[4 * C::NCH]) {
#pragma unroll
        for (int j = 0; j < C::NCH; ++j) {
            const float4 q4 = *reinterpret_cast<const float4*>(row + coff[j]);
            v[4 * j + 0] = q4.x; v[4 * j + 1] = q4.y; v[4 * j + 2] = q4.z; v[4 * j + 3] = q4.w;
        }
    };
    float S0[2][8], S1[2][8], S2[2][8];
#pragma unroll 1
    for (int pl = 0; pl < 3; ++pl) {
        const float* plane = s_prod + pl * C::PH * C::PPITCH + 2 * ty * C::PPITCH;
        // two row buffers: the loads of product row jj+1 are issued before the FMAs of row jj
        float va[4 * C::NCH], vb[4 * C::NCH];
        float lo[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) lo[p] = 0.0f;
        load_row(plane, va);
        load_row(plane + C::PPITCH, vb);
#pragma unroll
        for (int dx = 0; dx < G; ++dx)
#pragma unroll
            for (int p = 0; p < 8; ++p) lo[p] = __fmaf_rn(gw.w[dx], va[p + dx], lo[p]);               // tap row 0, upper output row
        unsigned long long acc2[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) acc2[p] = f2_pack(lo[p], 0.0f);
#pragma unroll
        for (int jj = 1; jj < G; ++jj) {
            float (&cur)[4 * C::NCH] = (jj & 1) ? vb : va;
            float (&nxt)[4 * C::NCH] = (jj & 1) ? va : vb;
            load_row(plane + (jj + 1) * C::PPITCH, nxt);
#pragma unroll
            for (int dx = 0; dx < G; ++dx) {
                const float2 w2 = gw.wp[jj * SFM_GW_PITCH + dx];
                const unsigned long long ww = f2_pack(w2.x, w2.y);
#pragma unroll
                for (int p = 0; p < 8; ++p) acc2[p] = f2_fma(f2_pack(cur[p + dx], cur[p + dx]), ww, acc2[p]);
            }
        }
        float a0[8], a1[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) f2_unpack(acc2[p], a0[p], a1[p]);
        float (&last)[4 * C::NCH] = (G & 1) ? vb : va;
#pragma unroll
        for (int dx = 0; dx < G; ++dx)
#pragma unroll
            for (int p = 0; p < 8; ++p) a1[p] = __fmaf_rn(gw.w[(G - 1) * SFM_GW_PITCH + dx], last[p + dx], a1[p]);   // tap row G-1, lower row
        if (pl == 0) {
#pragma unroll
            for (int p = 0; p < 8; ++p) { S0[0][p] = a0[p]; S0[1][p] = a1[p]; }
        } else if (pl == 1) {
#pragma unroll
            for (int p = 0; p < 8; ++p) { S1[0][p] = a0[p]; S1[1][p] = a1[p]; }
        } else {
#pragma unroll
            for (int p = 0; p < 8; ++p) { S2[0][p] = a0[p]; S2[1][p] = a1[p]; }
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const float sxx = S0[q][p], sxy = S1[q][p], syy = S2[q][p];
            const float det = __fsub_rn(__fmul_rn(sxx, syy), __fmul_rn(sxy, sxy));
            const float tr = __fadd_rn(sxx, syy);
            r[q][p] = __fsub_rn(det, __fmul_rn(alpha, __fmul_rn(tr, tr)));
        }
}

template <int WARPS_PER_CTA, int VAR>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k(const __grid_constant__ GaussWeights gw, float* out, int iters) {
    using C = HarrisCfg<7, 32>;
    extern __shared__ float s_prod[];
    for (int i = threadIdx.x; i < C::PROD_WORDS; i += blockDim.x) s_prod[i] = 1.0f + (i % 7) * 0.125f;
    __syncthreads();
    float r[2][8], acc = 0.f;
    const int tid = threadIdx.x & 127;
    for (int it = 0; it < iters; ++it) {
        if (VAR == 0) harris_window<7, 32, true>(s_prod, gw, 0.05f + it * 1e-9f, r, tid);
        else harris_window_rows<7, 32>(s_prod, gw, 0.05f + it * 1e-9f, r, tid);
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int p = 0; p < 8; ++p) acc += r[q][p];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int W, int VAR> void run(int ctas_per_sm) {
    using C = HarrisCfg<7, 32>;
    GaussWeights gw; for (int i = 0; i < 11 * 16; ++i) { gw.w[i] = 0.01f * (i % 13); gw.wp[i] = make_float2(0.01f * (i % 5), 0.02f * (i % 3)); }
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    size_t smem = C::PROD_WORDS * 4;
    cudaFuncSetAttribute(k<W, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int iters = 40;
    k<W, VAR><<<148 * ctas_per_sm, W * 32, smem>>>(gw, out, 2);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<W, VAR><<<148 * ctas_per_sm, W * 32, smem>>>(gw, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double px = (double)148 * ctas_per_sm * W * 32 * 16 * iters;
    printf("variant %d warps/CTA %d CTAs/SM %d (warps/SMSP %.1f): %.3f ms, %.2f Gpx/s, FMA-pipe share %.1f%% (147 FMA/px, 128/clk/SM @1.965GHz) err=%s\n", VAR, W, ctas_per_sm,
           W * ctas_per_sm / 4.0, ms, px / ms / 1e6, 100.0 * px * 147 / (ms * 1e-3) / (148.0 * 128 * 1.965e9), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() { run<4,0>(1); run<4,1>(1); run<4,0>(2); run<4,1>(2); run<4,0>(4); run<4,1>(4); run<8,1>(2); run<8,1>(3); return 0; }
