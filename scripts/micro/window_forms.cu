// Microbenchmark (round 2): instruction ORDER and work-per-thread forms of the 49-tap window stage of k_harris.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o window_forms window_forms.cu && ./window_forms
// Part 1 (registers only): 56 FFMA2 per iteration on 8 packed accumulators and 14 row values,
//   order 0 = tap-column outer (dx, then pixel): the row-value operand changes every instruction
//   order 1 = row-value outer (j, then dx): the same row value feeds up to 7 consecutive FFMA2 (.reuse)
// Part 2 (full window stage off a static shared-memory tile): 2 rows x 8 px per thread in both orders,
//   4 rows x 8 px per thread (each product row loaded once for four output rows).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define GP 16
struct __align__(16) GaussWeights {
    float w[11 * GP];
    float2 wp[11 * GP];     // wp[jj][dx] = (w[jj][dx], w[jj-1][dx])
};
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) { unsigned long long d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
// volatile: ptxas keeps the issue order of volatile asm statements (it re-sorts the plain form into tap-column order)
__device__ __forceinline__ unsigned long long f2_fma_v(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// ---------------------------------------------------------------- part 1
// MODE 0: shipped inner-loop form (row value broadcast, weight pair from the constant bank by index -> UR)
// MODE 1: the same with ONE row value for all 56 (is the changing 32-bit operand what costs?)
// MODE 2: weight pairs held in vector registers (no uniform operand)
// MODE 3: scalar FFMA, two accumulator sets (112 FFMA)
// MODE 4: MODE 0 plus 4 LDS.128 per iteration into the row values (what the real loop does)
template <int MODE>
__global__ void k_reg(const __grid_constant__ GaussWeights gw, float* out, int iters, float seed) {
    __shared__ float4 s_v[4 * 256];
    unsigned long long acc2[8];
    float accs[2][8];
    float v[16];
    for (int i = 0; i < 8; ++i) { acc2[i] = f2_pack(threadIdx.x + i, i); accs[0][i] = threadIdx.x + i; accs[1][i] = i; }
    for (int i = 0; i < 16; ++i) v[i] = seed + i + threadIdx.x * 0.001f;
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) s_v[i] = make_float4(seed, seed + 1, seed + 2, seed + 3);
    __syncthreads();
    unsigned long long wr[7];
    for (int dx = 0; dx < 7; ++dx) wr[dx] = f2_pack(seed + dx, seed - dx);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const float2* wp = gw.wp + (1 + (it & 3)) * GP;
        const float* w0 = gw.w + (1 + (it & 3)) * GP;
        if (MODE == 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 q = s_v[((it + j) & 3) * 256 + threadIdx.x % 256];
                v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
            }
        }
#pragma unroll
        for (int dx = 0; dx < 7; ++dx) {
            const unsigned long long ww = (MODE == 2) ? wr[dx] : f2_pack(wp[dx].x, wp[dx].y);
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                if (MODE == 3) {
                    accs[0][p] = __fmaf_rn(w0[dx], v[p + dx], accs[0][p]);
                    accs[1][p] = __fmaf_rn(w0[dx - GP], v[p + dx], accs[1][p]);
                } else if (MODE == 1) acc2[p] = f2_fma(f2_pack(v[0], v[0]), ww, acc2[p]);
                else acc2[p] = f2_fma(f2_pack(v[p + dx], v[p + dx]), ww, acc2[p]);
            }
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) { float a, b; f2_unpack(acc2[i], a, b); s += a + b + accs[0][i] + accs[1][i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------- part 2
constexpr int G = 7, R = 3, HT = 64;
constexpr int NV = 8 + 2 * R, NCH = 4, PCH = 20, PPITCH = PCH * 4;
template <int TH> struct Cfg { static constexpr int PH = TH + 2 * R; static constexpr int WORDS = 3 * PH * PPITCH; };

__device__ __forceinline__ void load_row(const float* row, int tx, float (&v)[16]) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int c = 2 * tx + j;
        const float4 q4 = *reinterpret_cast<const float4*>(row + (c ^ ((c >> 3) & 1)) * 4);
        v[4 * j + 0] = q4.x; v[4 * j + 1] = q4.y; v[4 * j + 2] = q4.z; v[4 * j + 3] = q4.w;
    }
}
template <int ORDER>
__device__ __forceinline__ void taps_f2(const float (&v)[16], const float2* wp, unsigned long long (&acc2)[8]) {
    unsigned long long ww[7];
#pragma unroll
    for (int dx = 0; dx < 7; ++dx) ww[dx] = f2_pack(wp[dx].x, wp[dx].y);
    if (ORDER == 0) {
#pragma unroll
        for (int dx = 0; dx < 7; ++dx)
#pragma unroll
            for (int p = 0; p < 8; ++p) acc2[p] = f2_fma(f2_pack(v[p + dx], v[p + dx]), ww[dx], acc2[p]);
    } else {
#pragma unroll
        for (int j = 0; j < 14; ++j)
#pragma unroll
            for (int dx = 6; dx >= 0; --dx) {
                const int p = j - dx;
                if (p >= 0 && p < 8) acc2[p] = f2_fma_v(f2_pack(v[j], v[j]), ww[dx], acc2[p]);
            }
    }
}
template <int ORDER>
__device__ __forceinline__ void taps_f1(const float (&v)[16], const float* w, float (&acc)[8]) {
    if (ORDER == 0) {
#pragma unroll
        for (int dx = 0; dx < 7; ++dx)
#pragma unroll
            for (int p = 0; p < 8; ++p) acc[p] = __fmaf_rn(w[dx], v[p + dx], acc[p]);
    } else {
#pragma unroll
        for (int j = 0; j < 14; ++j)
#pragma unroll
            for (int dx = 6; dx >= 0; --dx) {
                const int p = j - dx;
                if (p >= 0 && p < 8) acc[p] = __fmaf_rn(w[dx], v[j], acc[p]);
            }
    }
}
__device__ __forceinline__ float harris_r(float sxx, float sxy, float syy, float alpha) {
    const float det = __fsub_rn(__fmul_rn(sxx, syy), __fmul_rn(sxy, sxy));
    const float tr = __fadd_rn(sxx, syy);
    return __fsub_rn(det, __fmul_rn(alpha, __fmul_rn(tr, tr)));
}

// 2 rows x 8 px per thread (the shipped mapping), tile height 32, 128 threads
template <int ORDER>
__device__ __forceinline__ float window2(const float* s_prod, const GaussWeights& gw, float alpha, int tid) {
    using C = Cfg<32>;
    const int tx = tid & 7, ty = tid >> 3;
    float S[3][2][8];
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
        const float* plane = s_prod + pl * C::PH * PPITCH + 2 * ty * PPITCH;
        float acc[2][8], v[16];
#pragma unroll
        for (int p = 0; p < 8; ++p) { acc[0][p] = 0.0f; acc[1][p] = 0.0f; }
        load_row(plane, tx, v);
        taps_f1<ORDER>(v, gw.w, acc[0]);
        unsigned long long acc2[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) acc2[p] = f2_pack(acc[0][p], 0.0f);
#pragma unroll 1
        for (int jj = 1; jj < G; ++jj) {
            load_row(plane + jj * PPITCH, tx, v);
            taps_f2<ORDER>(v, gw.wp + jj * GP, acc2);
        }
#pragma unroll
        for (int p = 0; p < 8; ++p) f2_unpack(acc2[p], acc[0][p], acc[1][p]);
        load_row(plane + G * PPITCH, tx, v);
        taps_f1<ORDER>(v, gw.w + (G - 1) * GP, acc[1]);
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int p = 0; p < 8; ++p) S[pl][q][p] = acc[q][p];
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int p = 0; p < 8; ++p) s += harris_r(S[0][q][p], S[1][q][p], S[2][q][p], alpha);
    return s;
}

// 4 rows x 8 px per thread, tile height 64, 128 threads: product row jj (0..9) feeds
//   pair A (output rows 0,1) with tap rows (jj, jj-1), pair B (rows 2,3) with tap rows (jj-2, jj-3).
// Steady rows 3..6 run rolled (both pairs packed); the six ramp rows are peeled.
template <int ORDER>
__device__ __forceinline__ float window4(const float* s_prod, const GaussWeights& gw, float alpha, int tid) {
    using C = Cfg<64>;
    const int tx = tid & 7, ty = tid >> 3;
    float R4[4][8];
    float S[2][4][8];          // planes 0,1 kept; plane 2 consumed immediately
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
        const float* plane = s_prod + pl * C::PH * PPITCH + 4 * ty * PPITCH;
        float a0[8], v[16];
        unsigned long long A[8], B[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) a0[p] = 0.0f;
        load_row(plane, tx, v);                                   // jj = 0: row 0 tap 0
        taps_f1<ORDER>(v, gw.w, a0);
#pragma unroll
        for (int p = 0; p < 8; ++p) A[p] = f2_pack(a0[p], 0.0f);
        load_row(plane + PPITCH, tx, v);                          // jj = 1: A (1, 0)
        taps_f2<ORDER>(v, gw.wp + 1 * GP, A);
        load_row(plane + 2 * PPITCH, tx, v);                      // jj = 2: A (2, 1), row 2 tap 0
        taps_f2<ORDER>(v, gw.wp + 2 * GP, A);
#pragma unroll
        for (int p = 0; p < 8; ++p) a0[p] = 0.0f;
        taps_f1<ORDER>(v, gw.w, a0);
#pragma unroll
        for (int p = 0; p < 8; ++p) B[p] = f2_pack(a0[p], 0.0f);
#pragma unroll 1
        for (int jj = 3; jj < G; ++jj) {                          // jj = 3..6: A (jj, jj-1), B (jj-2, jj-3)
            load_row(plane + jj * PPITCH, tx, v);
            taps_f2<ORDER>(v, gw.wp + jj * GP, A);
            taps_f2<ORDER>(v, gw.wp + (jj - 2) * GP, B);
        }
        float r0[8], r1[8], r2[8], r3[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) f2_unpack(A[p], r0[p], r1[p]);
        load_row(plane + 7 * PPITCH, tx, v);                      // jj = 7: row 1 tap 6, B (5, 4)
        taps_f1<ORDER>(v, gw.w + 6 * GP, r1);
        taps_f2<ORDER>(v, gw.wp + 5 * GP, B);
        load_row(plane + 8 * PPITCH, tx, v);                      // jj = 8: B (6, 5)
        taps_f2<ORDER>(v, gw.wp + 6 * GP, B);
#pragma unroll
        for (int p = 0; p < 8; ++p) f2_unpack(B[p], r2[p], r3[p]);
        load_row(plane + 9 * PPITCH, tx, v);                      // jj = 9: row 3 tap 6
        taps_f1<ORDER>(v, gw.w + 6 * GP, r3);
        if (pl < 2) {
#pragma unroll
            for (int p = 0; p < 8; ++p) { S[pl][0][p] = r0[p]; S[pl][1][p] = r1[p]; S[pl][2][p] = r2[p]; S[pl][3][p] = r3[p]; }
        } else {
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                R4[0][p] = harris_r(S[0][0][p], S[1][0][p], r0[p], alpha);
                R4[1][p] = harris_r(S[0][1][p], S[1][1][p], r1[p], alpha);
                R4[2][p] = harris_r(S[0][2][p], S[1][2][p], r2[p], alpha);
                R4[3][p] = harris_r(S[0][3][p], S[1][3][p], r3[p], alpha);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int p = 0; p < 8; ++p) s += R4[q][p];
    return s;
}

template <int ROWS, int ORDER>
__global__ void __launch_bounds__(128) k_win(const __grid_constant__ GaussWeights gw, float* out, int iters) {
    extern __shared__ float s_prod[];
    constexpr int WORDS = (ROWS == 2) ? Cfg<32>::WORDS : Cfg<64>::WORDS;
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) s_prod[i] = 1.0f + (i % 7) * 0.125f;
    __syncthreads();
    float acc = 0.f;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (ROWS == 2) acc += window2<ORDER>(s_prod, gw, 0.05f + it * 1e-9f, threadIdx.x);
        else acc += window4<ORDER>(s_prod, gw, 0.05f + it * 1e-9f, threadIdx.x);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

static GaussWeights make_w() {
    GaussWeights gw;
    for (int i = 0; i < 11 * GP; ++i) { gw.w[i] = 0.01f * (i % 13); gw.wp[i] = make_float2(0.01f * (i % 5), 0.02f * (i % 3)); }
    return gw;
}
template <int ORDER> void run_reg(int warps) {
    static const char* names[] = {"shipped form", "one row value", "weights in vregs", "scalar FFMA x112", "shipped + 4 LDS.128"};
    GaussWeights gw = make_w();
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    int iters = 4000;
    k_reg<ORDER><<<148, 32 * warps>>>(gw, out, 10, 1.0f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k_reg<ORDER><<<148, 32 * warps>>>(gw, out, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)148 * 32 * warps * iters * 56 * 2;
    printf("reg-only mode %d (%-20s) warps/SM %2d: %.3f ms = %.1f FMA/clk/SM  (%s)\n", ORDER, names[ORDER], warps, ms, fmas / (ms * 1e-3) / 148 / 1.965e9,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
template <int ROWS, int ORDER> void run_win(int ctas_per_sm) {
    GaussWeights gw = make_w();
    float* out; cudaMalloc(&out, 148 * 16 * 128 * 4);
    size_t smem = ((ROWS == 2) ? Cfg<32>::WORDS : Cfg<64>::WORDS) * 4;
    cudaFuncSetAttribute(k_win<ROWS, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int iters = 40;
    k_win<ROWS, ORDER><<<148 * ctas_per_sm, 128, smem>>>(gw, out, 2);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k_win<ROWS, ORDER><<<148 * ctas_per_sm, 128, smem>>>(gw, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double px = (double)148 * ctas_per_sm * 128 * 8 * ROWS * iters;
    printf("window rows/thread %d order %d CTAs/SM %d (warps/SMSP %d, smem %zu KB/CTA): %.3f ms, FMA-pipe share %.1f%%  (%s)\n", ROWS, ORDER, ctas_per_sm,
           ctas_per_sm, smem / 1024, ms, 100.0 * px * 147 / (ms * 1e-3) / (148.0 * 128 * 1.965e9), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() {
    for (int w : {4, 8, 16, 20}) { run_reg<0>(w); run_reg<1>(w); run_reg<2>(w); run_reg<3>(w); run_reg<4>(w); }
    for (int c : {1, 2, 3, 4, 5}) { run_win<2, 0>(c); run_win<2, 1>(c); }
    for (int c : {1, 2, 3, 4}) { run_win<4, 0>(c); run_win<4, 1>(c); }
    return 0;
}
