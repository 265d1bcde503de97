// Microbenchmark: FP32 FMA issue rate on B200 for scalar FFMA and packed FFMA2 (fma.rn.f32x2).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
template <int MODE, int CH>
__global__ void k(float* out, int iters, float a, float b) {
    float acc[CH]; unsigned long long acc2[CH];
    for (int i = 0; i < CH; ++i) { acc[i] = threadIdx.x + i; acc2[i] = ((unsigned long long)(threadIdx.x + i) << 32) | i; }
    unsigned long long a2 = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
    unsigned long long b2 = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (MODE == 0) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(a), "f"(b));
                else acc2[i] = f2_fma(a2, acc2[i], b2);
            }
    }
    float s = 0; for (int i = 0; i < CH; ++i) s += acc[i] + (float)(acc2[i] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int CH> void run(const char* name, int warps_per_sm) {
    float* out; cudaMalloc(&out, 148 * 1024 * 8 * 4);
    int iters = 4000; dim3 grid(148), block(32 * warps_per_sm);
    k<MODE, CH><<<grid, block>>>(out, 10, 1.0001f, 0.5f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE, CH><<<grid, block>>>(out, iters, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)148 * 32 * warps_per_sm * iters * 8.0 * CH * (MODE ? 2 : 1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s warps/SM %2d chains %2d: %.3f ms  %.1f TFMA/s  = %.1f FMA/clk/SM at %d MHz\n", name, warps_per_sm, CH, ms, fmas / ms / 1e9,
           fmas / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
    cudaFree(out);
}
int main() {
    for (int w : {4, 8, 16, 32}) { run<0, 8>("FFMA", w); run<1, 8>("FFMA2", w); }
    run<0, 4>("FFMA", 8); run<1, 4>("FFMA2", 8); run<0, 16>("FFMA", 8); run<1, 16>("FFMA2", 8);
    run<1, 2>("FFMA2", 8); run<1, 2>("FFMA2", 16);
    return 0;
}
