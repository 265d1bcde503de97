// api.cu -- context management for libsfmb200 (C ABI in include/sfmb200.h).
#include <cstring>

#include "common.cuh"

// Per-launch accounting.  With profiling off this is one atomic increment; with
// it on every launch is bracketed by two CUDA events on the launching stream.
int sfm_prof_begin(SfmCtx* ctx, cudaStream_t st, const char* name) {
    ctx->launches.fetch_add(1);
    if (!ctx->prof_on) return -1;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaEvent_t ev[2];
    for (int i = 0; i < 2; ++i) {
        if (!ctx->ev_pool.empty()) { ev[i] = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); }
        else if (cudaEventCreate(&ev[i]) != cudaSuccess) return -1;
    }
    cudaEventRecord(ev[0], st);
    ctx->prof.push_back({name, ev[0], ev[1]});
    return (int)ctx->prof.size() - 1;
}

void sfm_prof_end(SfmCtx* ctx, cudaStream_t st, int idx) {
    if (idx < 0) return;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (idx < (int)ctx->prof.size()) cudaEventRecord(ctx->prof[idx].b, st);
}

extern "C" {

int sfm_version(void) { return SFM_API_VERSION; }

static thread_local std::string g_noctx_err;

int sfm_ctx_create(int device, SfmCtx** out) {
    if (!out) return SFM_ERR_BAD_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        g_noctx_err = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count is 0") +
                      " (libsfmb200 has no CPU fallback)";
        return SFM_ERR_CUDA;
    }
    if (device < 0 || device >= n) { g_noctx_err = "device index out of range"; return SFM_ERR_BAD_ARG; }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { g_noctx_err = cudaGetErrorString(e); return SFM_ERR_CUDA; }
    if (prop.major != 10 || prop.minor != 0) {       // the library holds sm_100a SASS only (no PTX): sm_103 cannot run it
        g_noctx_err = "libsfmb200 is built for sm_100a (B200) only; device is sm_" + std::to_string(prop.major) +
                      std::to_string(prop.minor);
        return SFM_ERR_UNSUPPORTED;
    }
    SfmCtx* c = new SfmCtx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    *out = c;
    return SFM_OK;
}

void sfm_ctx_destroy(SfmCtx* ctx) {
    if (!ctx) return;
    for (auto& r : ctx->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    delete ctx;
}

const char* sfm_last_error(SfmCtx* ctx) {
    if (!ctx) return g_noctx_err.c_str();
    std::lock_guard<std::mutex> g(ctx->mu);
    // keep the returned pointer stable for the calling thread
    static thread_local std::string copy;
    copy = ctx->err;
    return copy.c_str();
}

int sfm_ctx_sm_count(SfmCtx* ctx) { return ctx ? ctx->sm_count : 0; }

unsigned long long sfm_ctx_launch_count(SfmCtx* ctx) { return ctx ? ctx->launches.load() : 0ull; }

int sfm_ctx_set_option(SfmCtx* ctx, int option, int value) {
    if (!ctx) return SFM_ERR_BAD_ARG;
    switch (option) {
        case SFM_OPT_HARRIS_STREAM_MIN_BANDS:
            if (value < 0) return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "SFM_OPT_HARRIS_STREAM_MIN_BANDS must be >= 0");
            ctx->harris_stream_min_bands.store(value);
            return SFM_OK;
    }
    return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "unknown option %d", option);
}

int sfm_profile_enable(SfmCtx* ctx, int on) {
    if (!ctx) return SFM_ERR_BAD_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    for (auto& r : ctx->prof) { ctx->ev_pool.push_back(r.a); ctx->ev_pool.push_back(r.b); }
    ctx->prof.clear();
    ctx->prof_on = on != 0;
    return SFM_OK;
}

int sfm_profile_collect(SfmCtx* ctx, SfmKernelStat* out, int cap) {
    if (!ctx || (!out && cap > 0)) return SFM_ERR_BAD_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    int n = 0;
    for (auto& r : ctx->prof) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) ms = 0.f;
        int k = 0;
        for (; k < n; ++k) if (strncmp(out[k].name, r.name, sizeof(out[k].name) - 1) == 0) break;
        if (k == n) {
            if (n >= cap) continue;
            memset(&out[n], 0, sizeof(SfmKernelStat));
            strncpy(out[n].name, r.name, sizeof(out[n].name) - 1);
            ++n;
        }
        out[k].launches += 1;
        out[k].total_ms += ms;
        ctx->ev_pool.push_back(r.a);
        ctx->ev_pool.push_back(r.b);
    }
    ctx->prof.clear();
    return n;
}

}  // extern "C"
