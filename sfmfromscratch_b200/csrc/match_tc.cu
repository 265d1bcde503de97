// match_tc.cu -- tcgen05 / TMEM / TMA candidate pass of the NN-ratio matcher.
//
// For a work unit (pair, 256 query rows, column split) the CTA computes the
// 256 x N tile products  acc = A16 * B16^T  on the 5th-generation tensor cores
// (fp16 inputs, fp32 accumulation in TMEM), 256 train columns at a time: per
// column tile one M=128 x N=256 x K=16 UMMA chain for each of the two 128-row
// halves of the unit, each into its own 256-column TMEM accumulator.  N = 256
// is what keeps the tensor pipe fed: both operands come from shared memory, and
// an M128 N128 K16 instruction reads 8 KB per 64 cycles -- all of the 128 B/clk
// the SM's shared memory has, with the TMA fill and the epilogue still to
// serve -- while N256 reads 12 KB per 128 cycles.  The two accumulators double
// as the pipeline: the epilogue drains half 0 while the tensor pipe works on
// half 1 of the same column tile, and vice versa.
// The epilogue warps turn each accumulator straight into the ranking key
// |b|^2 - 2 a.b  and keep, per query row, the 4 smallest groups of GRP (4 or 8)
// columns.  The distance matrix never leaves TMEM / registers; the only global
// output is 16 bytes per (row, list).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator +
// single-thread MMA issuer, warps 2..9 = epilogue (TMEM lane quarter = warp % 4,
// column half of the 256-wide tile = (warp - 2) / 4).
// Pipelines: smem ring full/empty (TMA <-> MMA), A-tile full/empty per unit,
// TMEM accumulator full/empty per row half (MMA <-> epilogue).
#include <cuda.h>

#include "match.cuh"

namespace {

constexpr int TC_THREADS = 320;
constexpr int TC_STAGES = 2;                       // B smem ring
constexpr int TC_ACC = 2;                          // TMEM accumulators: one per 128-row half of the unit (256 columns each)
constexpr int TC_KBLK = 64;                        // fp16 elements per 128-byte swizzle row
constexpr uint32_t TC_SUB_BYTES = 128 * 128;       // one [128 rows][64 halves] box = 16 KB
constexpr uint32_t TC_A_BYTES = 4 * TC_SUB_BYTES;  // 2 row tiles x 2 k-blocks
constexpr uint32_t TC_B_BYTES = 4 * TC_SUB_BYTES;  // 2 k-blocks x [256 rows][64 halves] (two boxes each)
constexpr int TC_HALF = MT_COLS / 2;               // columns per epilogue warp and tile
constexpr uint32_t TC_NB_BYTES = 8 * 2 * TC_HALF * 4;   // per epilogue warp: 2 buffers of 128 norms
static_assert(MT_COLS == 256 && MT_ROWS == 256, "tile geometry");
constexpr uint32_t TC_SMEM = 1024 + TC_A_BYTES + TC_STAGES * TC_B_BYTES + TC_NB_BYTES + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
// wait::ld names the registers of the load it completes, so that nothing that reads them is scheduled above it
__device__ __forceinline__ void tc_wait_ld(uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.wait::ld.sync.aligned;"
        : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
          "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
          "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
          "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
        :
        : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4 in [0,14), LBO (unused for swizzled K-major, 1) in [16,30),
// SBO = 8 rows * 128 B = 1024 B >> 4 in [32,46), version 1 in [46,48),
// layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// cute::UMMA::InstrDescriptor for kind::f16: D = F32 (bits [4,6) = 1), A = B = F16
// (0), both K-major (0), N >> 3 in [17,23), M >> 4 in [24,29).
constexpr uint32_t TC_IDESC = (1u << 4) | ((uint32_t)(MT_COLS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

struct UnitInfo { int p, rb, split, qa, qb, n1, n2, t0, t1; bool active; };

__device__ __forceinline__ UnitInfo decode_unit(const MatchPlan& P, int unit) {
    UnitInfo u;
    const int rowblocks = P.nmax_pad / MT_ROWS;
    u.split = unit % P.n_splits;
    const int q = unit / P.n_splits;
    u.rb = q % rowblocks;
    u.p = P.p0 + q / rowblocks;
    u.qa = P.pairs[2 * u.p]; u.qb = P.pairs[2 * u.p + 1];
    u.n1 = P.set_cnt[u.qa]; u.n2 = P.set_cnt[u.qb];
    const int ntiles = (u.n2 + MT_COLS - 1) / MT_COLS;
    u.t0 = u.split * P.tiles_per_split;
    u.t1 = min(u.t0 + P.tiles_per_split, ntiles);
    u.active = (u.rb * MT_ROWS < u.n1);
    return u;
}

__device__ __forceinline__ void top4_insert(float (&m)[4], float k) {
    const float n3 = fminf(m[3], fmaxf(m[2], k));
    const float n2 = fminf(m[2], fmaxf(m[1], k));
    const float n1 = fminf(m[1], fmaxf(m[0], k));
    m[0] = fminf(m[0], k);
    m[1] = n1; m[2] = n2; m[3] = n3;
}

template <int GRP>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_match_tc(const __grid_constant__ MatchPlan P, const __grid_constant__ CUtensorMap tmap, int n_units) {
    using G = MtG<GRP>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;      // SWIZZLE_128B tiles need 1024 B alignment
    unsigned char* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t sA = sbase;
    const uint32_t sB = sA + TC_A_BYTES;
    float* s_nb = reinterpret_cast<float*>(sgen + TC_A_BYTES + TC_STAGES * TC_B_BYTES);
    const uint32_t sbar = sB + TC_STAGES * TC_B_BYTES + TC_NB_BYTES;
    // barrier slots (8 bytes each)
    const uint32_t bar_full = sbar;                         // [TC_STAGES]
    const uint32_t bar_empty = sbar + 8 * TC_STAGES;        // [TC_STAGES]
    const uint32_t bar_afull = sbar + 16 * TC_STAGES;
    const uint32_t bar_aempty = bar_afull + 8;
    const uint32_t bar_tfull = bar_aempty + 8;              // [TC_ACC]
    const uint32_t bar_tempty = bar_tfull + 8 * TC_ACC;     // [TC_ACC]
    const uint32_t tmem_slot = bar_tempty + 8 * TC_ACC;
    uint32_t* tmem_slot_gen = reinterpret_cast<uint32_t*>(sgen + (tmem_slot - sbase));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < TC_STAGES; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        mbar_init(bar_afull, 1);
        mbar_init(bar_aempty, 1);
        for (int i = 0; i < TC_ACC; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0, aphase = 0;
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
                const UnitInfo u = decode_unit(P, unit);
                if (!u.active || u.t0 >= u.t1) continue;
                mbar_wait(bar_aempty, aphase ^ 1);
                mbar_expect_tx(bar_afull, TC_A_BYTES);
                const int arow = u.qa * P.nmax_pad + u.rb * MT_ROWS;
#pragma unroll
                for (int rt = 0; rt < 2; ++rt)
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
                        tma_load_2d(sA + (rt * 2 + kb) * TC_SUB_BYTES, &tmap, bar_afull, kb * TC_KBLK, arow + rt * 128);
                aphase ^= 1;
                for (int t = u.t0; t < u.t1; ++t) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    mbar_expect_tx(bar_full + 8 * stage, TC_B_BYTES);
                    const int brow = u.qb * P.nmax_pad + t * MT_COLS;
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int rh = 0; rh < 2; ++rh)
                            tma_load_2d(sB + stage * TC_B_BYTES + (kb * 2 + rh) * TC_SUB_BYTES, &tmap, bar_full + 8 * stage,
                                        kb * TC_KBLK, brow + rh * 128);
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0, accphase = 0, aphase = 0;
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
                const UnitInfo u = decode_unit(P, unit);
                if (!u.active || u.t0 >= u.t1) continue;
                mbar_wait(bar_afull, aphase);
                aphase ^= 1;
                for (int t = u.t0; t < u.t1; ++t) {
                    mbar_wait(bar_full + 8 * stage, phase);
#pragma unroll
                    for (int rt = 0; rt < 2; ++rt) {
                        mbar_wait(bar_tempty + 8 * rt, accphase ^ 1);
                        tc_fence_after();
                        const uint32_t d = tmem_base + (uint32_t)(rt * MT_COLS);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint64_t ad = umma_desc(sA + (uint32_t)(rt * 2 + (k >> 2)) * TC_SUB_BYTES + (uint32_t)(k & 3) * 32u);
                            const uint64_t bd = umma_desc(sB + (uint32_t)stage * TC_B_BYTES + (uint32_t)(k >> 2) * 2u * TC_SUB_BYTES +
                                                          (uint32_t)(k & 3) * 32u);
                            tc_mma_f16(d, ad, bd, TC_IDESC, k > 0 ? 1u : 0u);
                        }
                        tc_commit(bar_tfull + 8 * rt);             // this half's accumulator is ready for the epilogue
                    }
                    tc_commit(bar_empty + 8 * stage);              // smem stage free once both halves' MMAs retire
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    accphase ^= 1;
                }
                tc_commit(bar_aempty);                             // A tile free for the next unit
            }
        }
    } else {
        // ------------------------------------------------ epilogue warps
        const int ew = warp - 2;
        const int q = warp & 3;                                // TMEM lane quarter this warp may read
        const int half = ew >> 2;                              // column half of every 256-column tile
        float* my_nb = s_nb + ew * (2 * TC_HALF);              // 2 buffers x 128 floats
        uint32_t accphase = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const UnitInfo u = decode_unit(P, unit);
            if (!u.active) continue;
            float m[2][4];
#pragma unroll
            for (int rt = 0; rt < 2; ++rt)
#pragma unroll
                for (int e = 0; e < 4; ++e) m[rt][e] = 3.0e38f;
            const float* nbrow = P.nb + (size_t)u.qb * P.nmax_pad + half * TC_HALF + lane;
            int buf = 0;
            float pre[4] = {0.f, 0.f, 0.f, 0.f};
            if (u.t0 < u.t1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) pre[i] = nbrow[u.t0 * MT_COLS + 32 * i];
            }
            for (int t = u.t0; t < u.t1; ++t) {
                float* nbs = my_nb + buf * TC_HALF;
#pragma unroll
                for (int i = 0; i < 4; ++i) nbs[lane + 32 * i] = pre[i];
                __syncwarp();
                if (t + 1 < u.t1) {                             // prefetch the next tile's norms
#pragma unroll
                    for (int i = 0; i < 4; ++i) pre[i] = nbrow[(t + 1) * MT_COLS + 32 * i];
                }
                const uint32_t tile_bits = (uint32_t)(t - u.t0) << G::GROUP_BITS;
                // 32 accumulator columns -> 32 / G::GROUP candidate groups of row (q, lane) of half rt
                auto process = [&](const uint32_t (&v)[32], int rt, int c) {
#pragma unroll
                    for (int g = 0; g < 32 / G::GROUP; ++g) {
                        float gm = 3.0e38f;
#pragma unroll
                        for (int e = 0; e < G::GROUP; e += 4) {
                            const float4 nb4 = *reinterpret_cast<const float4*>(nbs + c * 32 + g * G::GROUP + e);
                            const float k0 = __fmaf_rn(__uint_as_float(v[G::GROUP * g + e + 0]), -2.0f, nb4.x);
                            const float k1 = __fmaf_rn(__uint_as_float(v[G::GROUP * g + e + 1]), -2.0f, nb4.y);
                            const float k2 = __fmaf_rn(__uint_as_float(v[G::GROUP * g + e + 2]), -2.0f, nb4.z);
                            const float k3 = __fmaf_rn(__uint_as_float(v[G::GROUP * g + e + 3]), -2.0f, nb4.w);
                            const float g4 = fminf(fminf(k0, k1), fminf(k2, k3));
                            gm = (e == 0) ? g4 : fminf(gm, g4);
                        }
                        const uint32_t pk = (__float_as_uint(gm) & ~MT_IDX_MASK) | tile_bits | (uint32_t)(c * (32 / G::GROUP) + g);
                        top4_insert(m[rt], __uint_as_float(pk));
                    }
                };
#pragma unroll
                for (int rt = 0; rt < 2; ++rt) {
                    mbar_wait(bar_tfull + 8 * rt, accphase);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(rt * MT_COLS + half * TC_HALF);
                    uint32_t va[32], vb[32];
                    // the next 32 columns are in flight while the current ones are ranked
                    tc_ld32(taddr, va);
                    tc_wait_ld(va);
                    tc_ld32(taddr + 32, vb);
                    process(va, rt, 0);
                    tc_wait_ld(vb);
                    tc_ld32(taddr + 64, va);
                    process(vb, rt, 1);
                    tc_wait_ld(va);
                    tc_ld32(taddr + 96, vb);
                    process(va, rt, 2);
                    tc_wait_ld(vb);
                    tc_fence_before();                          // this half's accumulator is in registers: hand it back
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * rt);
                    process(vb, rt, 3);
                }
                buf ^= 1;
                accphase ^= 1;
            }
            // 16 bytes per (row, list)
            const int list = u.split * 2 + half;
#pragma unroll
            for (int rt = 0; rt < 2; ++rt) {
                const size_t row = (size_t)u.p * P.nmax_pad + (size_t)(u.rb * MT_ROWS + rt * 128 + q * 32 + lane);
                uint4 o;
                o.x = __float_as_uint(m[rt][0]); o.y = __float_as_uint(m[rt][1]);
                o.z = __float_as_uint(m[rt][2]); o.w = __float_as_uint(m[rt][3]);
                *reinterpret_cast<uint4*>(P.cands + (row * P.n_lists + list) * MT_TOPK) = o;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

int launch_match_tc(SfmCtx* ctx, cudaStream_t st, const MatchPlan& P) {
    {
        std::lock_guard<std::mutex> g(ctx->mu);
        if (!ctx->tmap_encode) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
            if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
                ctx->err = "cuTensorMapEncodeTiled not available from the driver";
                return SFM_ERR_CUDA;
            }
            ctx->tmap_encode = fn;
        }
    }
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)SFM_DESC_DIM, (cuuint64_t)P.n_sets * (cuuint64_t)P.nmax_pad};
    const cuuint64_t gstride[1] = {(cuuint64_t)SFM_DESC_DIM * sizeof(__half)};
    const cuuint32_t box[2] = {(cuuint32_t)TC_KBLK, 128u};
    const cuuint32_t estr[2] = {1u, 1u};
    CUresult r = ((PFN_encodeTiled)ctx->tmap_encode)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)P.h16, gdim, gstride,
                                                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return sfm_set_error(ctx, SFM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    const int n_units = P.pn * (P.nmax_pad / MT_ROWS) * P.n_splits;
    const int grid = n_units < ctx->sm_count ? n_units : ctx->sm_count;
    if (P.group == 4) {
        SFM_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_match_tc<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        SFM_LAUNCH(ctx, st, "k_match_tc", k_match_tc<4><<<grid, TC_THREADS, TC_SMEM, st>>>(P, tmap, n_units));
    } else {
        SFM_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_match_tc<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        SFM_LAUNCH(ctx, st, "k_match_tc", k_match_tc<8><<<grid, TC_THREADS, TC_SMEM, st>>>(P, tmap, n_units));
    }
    return SFM_OK;
}
