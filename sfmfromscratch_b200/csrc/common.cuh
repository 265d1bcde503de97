// common.cuh -- shared host/device helpers for libsfmb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sfmb200.h"

struct SfmCtx {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t smem_optin = 0;
    std::mutex mu;
    std::string err;
    void* tmap_encode = nullptr;   // cuTensorMapEncodeTiled, resolved lazily
    std::atomic<int> harris_stream_min_bands{24};   // SFM_OPT_HARRIS_STREAM_MIN_BANDS
    // launch accounting / optional per-kernel CUDA-event timing (sfm_profile_*)
    std::atomic<unsigned long long> launches{0};
    bool prof_on = false;
    struct ProfRec { const char* name; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
};

int sfm_prof_begin(SfmCtx* ctx, cudaStream_t st, const char* name);
void sfm_prof_end(SfmCtx* ctx, cudaStream_t st, int idx);

// Launch a kernel with accounting: SFM_LAUNCH(ctx, stream, "name", kernel<<<g, b, s, stream>>>(args));
#define SFM_LAUNCH(ctx, st, name, ...)                                                    \
    do {                                                                                  \
        int _pi = sfm_prof_begin((ctx), (st), (name));                                    \
        __VA_ARGS__;                                                                      \
        sfm_prof_end((ctx), (st), _pi);                                                   \
        SFM_LAUNCH_CHECK((ctx), (name));                                                  \
    } while (0)

int sfm_set_error(SfmCtx* ctx, int code, const char* fmt, ...);

#define SFM_CUDA_CHECK(ctx, expr)                                                         \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            return sfm_set_error((ctx), SFM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,     \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);             \
    } while (0)

#define SFM_LAUNCH_CHECK(ctx, name)                                                       \
    do {                                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess)                                                            \
            return sfm_set_error((ctx), SFM_ERR_CUDA, "launch of %s failed: %s", (name),  \
                                 cudaGetErrorString(_e));                                 \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Order-preserving float <-> uint32 key (ascending key == ascending float).
__host__ __device__ __forceinline__ uint32_t f32_to_key(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float key_to_f32(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
