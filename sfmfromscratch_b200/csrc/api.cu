// api.cu -- context management for libsfmb200 (C ABI in include/sfmb200.h).
#include "common.cuh"

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
    if (prop.major != 10) {
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

void sfm_ctx_destroy(SfmCtx* ctx) { delete ctx; }

const char* sfm_last_error(SfmCtx* ctx) {
    if (!ctx) return g_noctx_err.c_str();
    std::lock_guard<std::mutex> g(ctx->mu);
    // keep the returned pointer stable for the calling thread
    static thread_local std::string copy;
    copy = ctx->err;
    return copy.c_str();
}

int sfm_ctx_sm_count(SfmCtx* ctx) { return ctx ? ctx->sm_count : 0; }

}  // extern "C"
