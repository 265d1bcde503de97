// Microbenchmark: FFMA2 operand forms (packed x2 everywhere vs scalar-broadcast .F32 operand) on B200.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) { unsigned long long d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
struct W { float2 w[64]; };
template <int MODE>
__global__ void k(const __grid_constant__ W cw, float* out, int iters, float seed) {
    unsigned long long acc2[8];
    float v[14];
    for (int i = 0; i < 8; ++i) acc2[i] = f2_pack(threadIdx.x + i, i);
    for (int i = 0; i < 14; ++i) v[i] = seed + i + threadIdx.x * 0.001f;
    for (int it = 0; it < iters; ++it) {
        unsigned long long ww[7];
#pragma unroll
        for (int dx = 0; dx < 7; ++dx) {
            if (MODE == 2) { const float2 t = cw.w[(it & 7) * 7 + dx]; ww[dx] = f2_pack(t.x, t.y); }   // indexed constant load
            else ww[dx] = f2_pack(seed + dx, seed - dx);
        }
#pragma unroll
        for (int dx = 0; dx < 7; ++dx)
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                if (MODE == 0) acc2[p] = f2_fma(f2_pack(v[p + dx], v[(p + dx + 1) % 14]), ww[dx], acc2[p]);       // true pairs
                else acc2[p] = f2_fma(f2_pack(v[p + dx], v[p + dx]), ww[dx], acc2[p]);                            // broadcast
            }
#pragma unroll
        for (int i = 0; i < 14; ++i) v[i] += 1.0f;   // 14 FADD per 56 FFMA2 (stands in for the row loads)
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += (float)(acc2[i] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int warps) {
    W cw; for (int i = 0; i < 64; ++i) cw.w[i] = make_float2(0.001f * i, 0.002f * i);
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    int iters = 2000;
    k<MODE><<<148, 32 * warps>>>(cw, out, 10, 1.0f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<148, 32 * warps>>>(cw, out, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)148 * 32 * warps * iters * 56 * 2;
    printf("%-34s warps/SM %2d: %.3f ms = %.1f FMA/clk/SM\n", name, warps, ms, fmas / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}
int main() {
    for (int w : {4, 8, 16}) { run<0>("packed pairs, reg weights", w); run<1>("scalar broadcast, reg weights", w); run<2>("scalar broadcast, LDC weights", w); }
    return 0;
}
