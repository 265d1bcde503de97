// harris_stream.cuh -- k_harris_stream: the Harris response of one pyramid level as a persistent, warp-specialised
// stream (included by extract.cu after its helpers: sobel_chain, f2_*, harris_pair_taps).
//
// Why.  R must be bit-identical to cv2.filter2D's row-major fmaf chains (NaiveSIFT.py:60-74): 3 x 49 dependent FMAs
// per pixel, an FP32-pipe floor no separable or tensor-core form can replace.  The one-tile-per-CTA kernel (k_harris)
// reaches 46 % of that floor for two reasons measured in round 2 (scripts/micro/window_forms.cu, profiles/):
//   * its window stage reads G+1 product rows from shared memory per 2 output rows and is bound by that operand
//     traffic at ~72 % of the FMA rate at any occupancy; with 4 output rows per thread (G+3 rows per 4) the same
//     chains run at 85-89 %, but
//   * 4 rows per thread in a tile kernel leaves 8 warps per SM, too few to hide the tile load, the Sobel/product
//     phase and the barriers between them: measured slower (0.89 ms against 0.77 per 32 x 1080p).
// So the phases are decoupled instead of interleaved by occupancy.  One persistent CTA per SM, 16 warps:
//   warps 8-15  CONSUMERS: the window stage only, 4 rows x 8 px per thread (a warp = one 64 x 16 band), product rows
//               read from a shared-memory ring, R stored from registers, first radix-select histogram;
//   warps 0-7   PRODUCERS (two groups of four, alternating chunks): Sobel + the three products (NaiveSIFT.py:61-64)
//               of one 16-row chunk at a time, from a TMA-staged image tile (cp.async.bulk.tensor, zero fill outside
//               the image == BORDER_CONSTANT, issued ISTAGES - NGROUP chunks ahead by the group's warps in turn) into
//               the ring; they also emit the strip's part of pyramid level l+1 (the exact 2x2 mean) from that tile.
// Registers are re-dealt between the roles with setmaxnreg; the warp index is taken through __shfl_sync so that ptxas
// knows it is warp-uniform (role-local loop counters and constant-bank indices then stay on the uniform datapath).
// The image is cut into vertical strips of 64 columns and each strip into bands of 16 rows; the launch's bands
// (image-major, strip, band) are dealt to the CTAs as equal contiguous ranges.  Inside a range, consecutive bands of
// one strip form a run: band k needs product rows [16k-R, 16k+16+R) = chunk k plus the first 2R rows of chunk k+1,
// so a run of m bands costs m+1 chunks (the last one short) and NO vertical halo is recomputed inside a run.
// Ring slots are handed over with mbarriers (full: the producer warps; empty: the two bands that read a chunk).
#pragma once

namespace hs {

constexpr int SW = 64;                 // strip width
constexpr int BH = 16;                 // band height == chunk height
constexpr int NCONS = 8, NPROD = 4;    // consumer warps; producer warps PER GROUP
constexpr int NGROUP = 2;              // producer groups: group g computes the chunks q with q % NGROUP == g
constexpr int THREADS = 32 * (NCONS + NPROD * NGROUP);
constexpr int REGS_PROD = 64, REGS_CONS = 192;   // setmaxnreg: 16 warps x 128 registers re-dealt between the roles
constexpr int RING = 12;               // product chunks in flight (8 bands being read need 9)
constexpr int ISTAGES = 8;             // staged image tiles (each producer group cycles through ISTAGES / NGROUP of them)

template <int G> struct Cfg {
    static constexpr int R = G / 2;
    static constexpr int RA = (R + 1 + 3) & ~3;          // image tile starts RA columns left of the strip (16-byte aligned)
    static constexpr int OFF = RA - (R + 1);             // product column c reads tile columns c+OFF .. c+OFF+2
    static constexpr int NV = 8 + 2 * R;                 // product values a consumer thread needs per row
    static constexpr int NCH = (NV + 3) / 4;             // ... in 16-byte chunks
    static constexpr int PCH = (14 + NCH + 1) & ~1;      // 16-byte chunks per product row (even: the XOR swizzle stays in range)
    static constexpr int PPITCH = PCH * 4;               // floats per product row
    static constexpr int IPITCH = (PCH * 4 + OFF + 2 + 3) & ~3;
    static constexpr int IH = BH + 2;                    // image rows per chunk
    // Ring layout [plane][ring row][PPITCH]: the chunk in slot s is ring rows 16 s .. 16 s + 15 of every plane, so the
    // 16 + 2R product rows a band reads are CONSECUTIVE ring rows and a consumer thread addresses them as one base
    // register per 16-byte chunk plus an immediate.  The one exception, a band whose first chunk sits in the last slot,
    // is removed by a shadow: rows 0 .. 2R-1 of slot 0 are also written behind the last slot.
    static constexpr int RROWS = RING * BH + 2 * R;      // ring rows per plane, shadow included
    static constexpr int PSTRIDE = RROWS * PPITCH;       // floats per plane
    static constexpr int LASTV = NV - 4 * (NCH - 1);     // floats a thread needs of its last 16-byte chunk
    static constexpr int ITILE = IH * IPITCH;            // floats per staged image tile (the TMA box)
    static constexpr int ISTRIDE = (ITILE + 31) & ~31;   // stage stride: TMA destinations are 128-byte aligned
    static constexpr int EMIT = (R + 1) & ~1;            // chunk kc emits next-level rows from image row 16 kc - EMIT on (even, inside the tile)
    static constexpr size_t ring_bytes = (sizeof(float) * 3 * (size_t)PSTRIDE + 127) / 128 * 128;   // the TMA stages behind it are 128-byte aligned
    static constexpr size_t img_bytes = sizeof(float) * (size_t)ISTRIDE * ISTAGES;
    static constexpr size_t hist_bytes = sizeof(uint32_t) * SFM_HIST1_BINS;
    static constexpr size_t bar_bytes = 8 * (2 * RING + 2 * ISTAGES);
    static constexpr size_t smem_bytes = ring_bytes + img_bytes + hist_bytes + bar_bytes;
    static_assert(2 * R <= BH, "a band reads at most two chunks");
    static_assert(EMIT <= R + 1 && EMIT >= R - 1, "the emitted rows lie inside the staged tile");
};

struct Geo {             // launch geometry, the same on every thread
    int H, W, S, K;      // level size, strips per image, bands per strip
    int n0, n1;          // this CTA's band range (the launcher checks that B * S * K fits an int)
};

__device__ __forceinline__ void decompose(const Geo& g, int n, int& b, int& s, int& k) {
    const int strip = n / g.K;
    k = n - strip * g.K;
    b = strip / g.S;
    s = strip - b * g.S;
}
// ring sequence number of band n's first chunk: one chunk per band plus one (short) chunk per finished run
__device__ __forceinline__ int chunk_seq(const Geo& g, int n) { return (n - g.n0) + (n / g.K - g.n0 / g.K); }

}  // namespace hs

template <int G>
__global__ void __launch_bounds__(hs::THREADS, 1)
k_harris_stream(const __grid_constant__ ExtractPlan P, const __grid_constant__ GaussWeights gw, int l, int fuse_next,
                const __grid_constant__ CUtensorMap tmap) {
    using C = hs::Cfg<G>;
    using namespace sfm_tma;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_ring = reinterpret_cast<float*>(smem_raw);
    float* s_img = reinterpret_cast<float*>(smem_raw + C::ring_bytes);
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw + C::ring_bytes + C::img_bytes);
    const uint32_t bar0 = smem_u32(smem_raw + C::ring_bytes + C::img_bytes + C::hist_bytes);
    auto bar_full = [&](int slot) { return bar0 + 8u * (uint32_t)slot; };
    auto bar_empty = [&](int slot) { return bar0 + 8u * (uint32_t)(hs::RING + slot); };
    auto bar_ifull = [&](int st) { return bar0 + 8u * (uint32_t)(2 * hs::RING + st); };
    auto bar_iempty = [&](int st) { return bar0 + 8u * (uint32_t)(2 * hs::RING + hs::ISTAGES + st); };

    const LevelInfo& lv = P.lv[l];
    hs::Geo g;
    g.H = lv.H; g.W = lv.W;
    g.S = (g.W + hs::SW - 1) / hs::SW;
    g.K = (g.H + hs::BH - 1) / hs::BH;
    const long long NB = (long long)P.B * g.S * g.K;
    g.n0 = (int)(NB * blockIdx.x / gridDim.x);
    g.n1 = (int)(NB * (blockIdx.x + 1) / gridDim.x);
    // the warp index through a shuffle: ptxas then knows it is warp-uniform and keeps the role branches, loop counters and
    // constant-bank indices below on the uniform datapath
    const int t = threadIdx.x, warp = __shfl_sync(0xffffffffu, t >> 5, 0), lane = t & 31;

    if (t == 0) {
        for (int i = 0; i < hs::RING; ++i) { mbar_init(bar_full(i), hs::NPROD); mbar_init(bar_empty(i), 2); }
        for (int i = 0; i < hs::ISTAGES; ++i) { mbar_init(bar_ifull(i), 1); mbar_init(bar_iempty(i), hs::NPROD); }
        mbar_fence_init();
    }
    for (int i = t; i < SFM_HIST1_BINS; i += hs::THREADS) s_hist[i] = 0;
    __syncthreads();
    if (g.n0 >= g.n1) return;

    // Producers take the LOWEST warp ids: the warp scheduler favours older warps, and with the consumers in front the
    // producers only ran in the gaps both consumers of their sub-partition left (2 270 clk per chunk: the bottleneck).
    if (warp >= hs::NPROD * hs::NGROUP) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(hs::REGS_CONS));
        const int cw = warp - hs::NPROD * hs::NGROUP;      // consumer warp 0 .. 7
        // ================================================================== consumers
        const int tx = lane & 7, ty = lane >> 3;
        int coff[C::NCH];
#pragma unroll
        for (int j = 0; j < C::NCH; ++j) { const int c = 2 * tx + j; coff[j] = (c ^ ((c >> 3) & 1)) * 4; }
        const uint32_t ring_u32 = smem_u32(s_ring);
        int b0i, s0i, k0i;
        hs::decompose(g, g.n0, b0i, s0i, k0i);
        int epoch = 0;                                     // images of this CTA's range whose histogram has been flushed
        const int ct = t - 32 * hs::NPROD * hs::NGROUP;    // 0 .. 255 among the consumers
        auto flush = [&](int b) {
            // every consumer warp has finished image b's bands: push the shared histogram to the segment's and clear it
            bar_sync(1, 32 * hs::NCONS);
            uint32_t* ghist = P.hist1 + (size_t)(b * P.L + l) * SFM_HIST1_BINS;
            for (int i = ct; i < SFM_HIST1_BINS / 4; i += 32 * hs::NCONS) {
                const uint4 c = reinterpret_cast<const uint4*>(s_hist)[i];
                if (c.x | c.y | c.z | c.w) {
                    if (c.x) atomicAdd(ghist + 4 * i + 0, c.x);
                    if (c.y) atomicAdd(ghist + 4 * i + 1, c.y);
                    if (c.z) atomicAdd(ghist + 4 * i + 2, c.z);
                    if (c.w) atomicAdd(ghist + 4 * i + 3, c.w);
                    reinterpret_cast<uint4*>(s_hist)[i] = make_uint4(0, 0, 0, 0);
                }
            }
            bar_sync(1, 32 * hs::NCONS);
        };
        for (int n = g.n0 + cw; ; n += hs::NCONS) {
            int b, s, k;
            const bool live = n < g.n1;
            hs::decompose(g, live ? n : g.n1 - 1, b, s, k);
            // histogram epochs: one per image of the range; every warp crosses every boundary exactly once
            const int target = live ? (b - b0i) : (b - b0i + 1);
            while (epoch < target) { flush(b0i + epoch); ++epoch; }
            if (!live) break;
            const int cs = hs::chunk_seq(g, n);
            const int slot0 = cs % hs::RING, slot1 = (cs + 1) % hs::RING;
            mbar_wait(bar_full(slot0), (uint32_t)((cs / hs::RING) & 1));
            mbar_wait(bar_full(slot1), (uint32_t)(((cs + 1) / hs::RING) & 1));
            // product row jj of this thread's 4 output rows is ring row 16 slot0 + 4 ty + jj of the plane (consecutive rows:
            // see Cfg): one base register per 16-byte chunk and plane, the row as an immediate offset
            uint32_t pbase[C::NCH];
            auto set_plane = [&](int pl) {
#pragma unroll
                for (int j = 0; j < C::NCH; ++j)
                    pbase[j] = ring_u32 + 4u * (uint32_t)(pl * C::PSTRIDE + (slot0 * hs::BH + 4 * ty) * C::PPITCH + coff[j]);
            };
            auto load_at = [&](uint32_t rowoff_bytes, auto imm_tag, float (&v)[4 * C::NCH]) {
                constexpr int IMM = decltype(imm_tag)::value;
#pragma unroll
                for (int j = 0; j < C::NCH; ++j) {
                    if (j == C::NCH - 1 && C::LASTV <= 2) {
                        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v[4 * j]), "=f"(v[4 * j + 1]) : "r"(pbase[j] + rowoff_bytes), "n"(IMM));
                        v[4 * j + 2] = 0.0f; v[4 * j + 3] = 0.0f;
                    } else {
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
                                     : "=f"(v[4 * j]), "=f"(v[4 * j + 1]), "=f"(v[4 * j + 2]), "=f"(v[4 * j + 3]) : "r"(pbase[j] + rowoff_bytes), "n"(IMM));
                    }
                }
            };
            unsigned long long A[8], Bq[8];
            unsigned long long S0a[8], S0b[8], S1a[8], S1b[8];        // the sums of the two planes before the current one
#pragma unroll
            for (int p = 0; p < 8; ++p) { S0a[p] = S0b[p] = S1a[p] = S1b[p] = 0ull; }
#pragma unroll 1
            for (int pl = 0; pl < 3; ++pl) {
                float v[4 * C::NCH];
                set_plane(pl);
#pragma unroll
                for (int p = 0; p < 8; ++p) { A[p] = 0ull; Bq[p] = 0ull; }
                auto step = [&](auto jj_tag) {
                    constexpr int jj = decltype(jj_tag)::value;
                    constexpr bool AU = (jj < G), AL = (jj >= 1 && jj <= G), BU = (jj >= 2 && jj < G + 2), BL = (jj >= 3 && jj <= G + 2);
                    load_at(0u, std::integral_constant<int, jj * C::PPITCH * 4>{}, v);
                    harris_pair_taps<G, AU, AL>(v, gw.wp + (AU && AL ? jj : 0) * SFM_GW_PITCH, gw.w + (AU ? jj : 0) * SFM_GW_PITCH,
                                                gw.w + (AL ? jj - 1 : 0) * SFM_GW_PITCH, A);
                    harris_pair_taps<G, BU, BL>(v, gw.wp + (BU && BL ? jj - 2 : 0) * SFM_GW_PITCH, gw.w + (BU ? jj - 2 : 0) * SFM_GW_PITCH,
                                                gw.w + (BL ? jj - 3 : 0) * SFM_GW_PITCH, Bq);
                };
                step(std::integral_constant<int, 0>{});
                step(std::integral_constant<int, 1>{});
                step(std::integral_constant<int, 2>{});
                // rows 3 .. G-1 (both pairs packed) peeled as well: with a compile-time row every weight pair is a
                // uniform-register operand fetched by LDCU; left rolled, the row index lives in a vector register
                // (ptxas does not treat it as warp-uniform inside the role branch) and the 14 weight pairs of a row
                // arrive by indexed LDC into vector registers, ~60 clk ahead of the row's first FFMA2
                if constexpr (G >= 5) step(std::integral_constant<int, 3>{});
                if constexpr (G >= 5) step(std::integral_constant<int, 4>{});
                if constexpr (G >= 7) step(std::integral_constant<int, 5>{});
                if constexpr (G >= 7) step(std::integral_constant<int, 6>{});
                static_assert(G <= 7, "peel rows 7 .. G-1 for larger windows");
                if constexpr (G >= 3) {
                    step(std::integral_constant<int, G>{});
                    step(std::integral_constant<int, G + 1>{});
                    step(std::integral_constant<int, G + 2>{});
                } else {
                    step(std::integral_constant<int, 3>{});
                }
                if (pl < 2) {                              // rotate: (S0, S1) <- (S1, this plane)
#pragma unroll
                    for (int p = 0; p < 8; ++p) { S0a[p] = S1a[p]; S0b[p] = S1b[p]; S1a[p] = A[p]; S1b[p] = Bq[p]; }
                }
            }
            // the ring is no longer needed by this band: hand the two chunks back before the stores
            __syncwarp();
            if (lane == 0) {
                const bool first_in_run = (k == 0) || (n == g.n0);
                const bool last_in_run = (k == g.K - 1) || (n == g.n1 - 1);
                mbar_arrive(bar_empty(slot0));
                if (first_in_run) mbar_arrive(bar_empty(slot0));
                mbar_arrive(bar_empty(slot1));
                if (last_in_run) mbar_arrive(bar_empty(slot1));
            }
            // R (NaiveSIFT.py:71-74, every op rounded; the row pairs stay packed: mul / add / sub .f32x2 round each lane
            // exactly as the scalar forms do), store, first radix-select histogram.  W % 4 == 0 (launch condition).
            float* Rout = P.R + (size_t)b * P.r_stride + lv.r_off;
            const int gx = s * hs::SW + 8 * tx;
            const int gy0 = k * hs::BH + 4 * ty;
            const unsigned long long alpha2 = f2_pack(P.alpha, P.alpha);
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {                 // rows (0, 1) then (2, 3)
                float r0[8], r1[8];
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const unsigned long long sxx = h2 ? S0b[p] : S0a[p], sxy = h2 ? S1b[p] : S1a[p], syy = h2 ? Bq[p] : A[p];
                    const unsigned long long det = f2_sub(f2_mul(sxx, syy), f2_mul(sxy, sxy));
                    const unsigned long long tr = f2_add(sxx, syy);
                    f2_unpack(f2_sub(det, f2_mul(alpha2, f2_mul(tr, tr))), r0[p], r1[p]);
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float (&r)[8] = q ? r1 : r0;
                    const int gy = gy0 + 2 * h2 + q;
                    if (gy < g.H) {
                        float* o = Rout + (size_t)gy * g.W + gx;
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf)
                            if (gx + 4 * hf + 4 <= g.W) {
                                reinterpret_cast<float4*>(o)[hf] = make_float4(r[4 * hf], r[4 * hf + 1], r[4 * hf + 2], r[4 * hf + 3]);
#pragma unroll
                                for (int p = 4 * hf; p < 4 * hf + 4; ++p) atomicAdd(&s_hist[f32_to_key(r[p]) >> 20], 1u);
                            }
                    }
                }
            }
        }
    } else {
        // ================================================================== producers
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(hs::REGS_PROD));
        const int pw = warp & (hs::NPROD - 1);             // 0 .. 3 inside the group
        const int grp = warp / hs::NPROD;
        const LevelInfo& nx = P.lv[(l + 1 < P.L) ? l + 1 : l];
        // A cursor over the chunk sequence of this CTA's band range.  Chunk q lives in ring slot q % RING and image
        // stage q % ISTAGES; the phase bits flip at every wrap.  Advancing costs a handful of integer instructions.
        struct Cur { int n, b, s, k, extra, slot, sph, st, iph; };
        auto last_in_run = [&](const Cur& c) { return (c.k == g.K - 1) || (c.n == g.n1 - 1); };
        auto advance = [&](Cur& c) {
            if (c.extra == 0 && last_in_run(c)) c.extra = 1;
            else {
                c.extra = 0; ++c.n;
                if (++c.k == g.K) { c.k = 0; if (++c.s == g.S) { c.s = 0; ++c.b; } }
            }
            if (++c.slot == hs::RING) { c.slot = 0; c.sph ^= 1; }
            if (++c.st == hs::ISTAGES) { c.st = 0; c.iph ^= 1; }
        };
        Cur c;
        hs::decompose(g, g.n0, c.b, c.s, c.k);
        c.n = g.n0; c.extra = 0; c.slot = 0; c.sph = 0; c.st = 0; c.iph = 0;
        for (int i = 0; i < grp && c.n < g.n1; ++i) advance(c);          // group g starts at chunk g
        Cur f = c;                                                       // the chunk being fetched: ISTAGES / NGROUP of this group's ahead
        int fturn = 0;                                                   // the group's warps issue the loads in turn
        auto fetch = [&]() {
            if (f.n >= g.n1) return;
            if (fturn == pw) {
                // the stage was this group's, ISTAGES chunks ago: wait until its four warps have let go of it (a fresh
                // barrier reports the phase before its first as complete)
                mbar_wait(bar_iempty(f.st), (uint32_t)(f.iph ^ 1));
                if (lane == 0) {
                    mbar_expect_tx(bar_ifull(f.st), (uint32_t)(C::ITILE * sizeof(float)));
                    tma_load_3d(smem_u32(s_img + (size_t)f.st * C::ISTRIDE), &tmap, bar_ifull(f.st), f.s * hs::SW - C::RA,
                                (f.k + f.extra) * hs::BH - C::R - 1, f.b);
                }
                __syncwarp();
            }
            fturn = (fturn + 1) & (hs::NPROD - 1);
#pragma unroll
            for (int i = 0; i < hs::NGROUP; ++i) if (f.n < g.n1) advance(f);
        };
#pragma unroll 1
        for (int i = 0; i < hs::ISTAGES / hs::NGROUP - 1; ++i) fetch();
#pragma unroll 1
        while (c.n < g.n1) {
            fetch();
            const int kc = c.k + c.extra;
            const bool first_in_run = (c.k == 0) || (c.n == g.n0);
            const int rows = c.extra ? 2 * C::R : hs::BH;          // the run's closing chunk: only its first 2R rows are read
            mbar_wait(bar_empty(c.slot), (uint32_t)(c.sph ^ 1));   // (a fresh barrier reports the phase before its first as complete)
            mbar_wait(bar_ifull(c.st), (uint32_t)c.iph);
            const float* tile = s_img + (size_t)c.st * C::ISTRIDE;
            float* dst = s_ring + c.slot * hs::BH * C::PPITCH;
            const int shadow = (c.slot == 0) ? hs::RING * hs::BH * C::PPITCH : 0;      // rows 0 .. 2R-1 of slot 0 go behind the last slot too
            const int x0 = c.s * hs::SW, y0 = kc * hs::BH;         // product row py is image row y0 - R + py
            // PCH * BH = 288 strip tasks of 4 pixels: two per thread, and a third for one warp of the group in turn
            const int ntask = C::PCH * rows;
            auto task = [&](int ti) {
                const int py = ti / C::PCH, c4 = ti - py * C::PCH;
                const float* ip = tile + py * C::IPITCH + 4 * c4 + C::OFF;
                float w0[6], w1[6], w2[6];
                if constexpr ((C::OFF & 3) == 0) {
                    const float4 a = *reinterpret_cast<const float4*>(ip);
                    const float2 a2 = *reinterpret_cast<const float2*>(ip + 4);
                    const float4 b4 = *reinterpret_cast<const float4*>(ip + C::IPITCH);
                    const float2 b2 = *reinterpret_cast<const float2*>(ip + C::IPITCH + 4);
                    const float4 d4 = *reinterpret_cast<const float4*>(ip + 2 * C::IPITCH);
                    const float2 d2 = *reinterpret_cast<const float2*>(ip + 2 * C::IPITCH + 4);
                    w0[0] = a.x; w0[1] = a.y; w0[2] = a.z; w0[3] = a.w; w0[4] = a2.x; w0[5] = a2.y;
                    w1[0] = b4.x; w1[1] = b4.y; w1[2] = b4.z; w1[3] = b4.w; w1[4] = b2.x; w1[5] = b2.y;
                    w2[0] = d4.x; w2[1] = d4.y; w2[2] = d4.z; w2[3] = d4.w; w2[4] = d2.x; w2[5] = d2.y;
                } else {
#pragma unroll
                    for (int e = 0; e < 6; ++e) { w0[e] = ip[e]; w1[e] = ip[C::IPITCH + e]; w2[e] = ip[2 * C::IPITCH + e]; }
                }
                float xx[4], xy[4], yy[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float sx, sy;
                    sobel_chain(w0[e], w0[e + 1], w0[e + 2], w1[e], w1[e + 2], w2[e], w2[e + 1], w2[e + 2], sx, sy);
                    xx[e] = __fmul_rn(sx, sx);
                    yy[e] = __fmul_rn(sy, sy);
                    xy[e] = __fmul_rn(sx, sy);
                }
                // outside the image the PRODUCTS are zero: the window filter pads the product planes, not the image
                const int gy = y0 - C::R + py, gx0 = x0 - C::R + 4 * c4;
                if (!(gy >= 0 && gy < g.H && gx0 >= 0 && gx0 + 3 < g.W)) {
                    const bool rowok = (gy >= 0 && gy < g.H);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (!(rowok && gx0 + e >= 0 && gx0 + e < g.W)) { xx[e] = 0.0f; xy[e] = 0.0f; yy[e] = 0.0f; }
                }
                float* o = dst + py * C::PPITCH + (c4 ^ ((c4 >> 3) & 1)) * 4;
                *reinterpret_cast<float4*>(o) = make_float4(xx[0], xx[1], xx[2], xx[3]);
                *reinterpret_cast<float4*>(o + C::PSTRIDE) = make_float4(xy[0], xy[1], xy[2], xy[3]);
                *reinterpret_cast<float4*>(o + 2 * C::PSTRIDE) = make_float4(yy[0], yy[1], yy[2], yy[3]);
                if (shadow && py < 2 * C::R) {
                    *reinterpret_cast<float4*>(o + shadow) = make_float4(xx[0], xx[1], xx[2], xx[3]);
                    *reinterpret_cast<float4*>(o + shadow + C::PSTRIDE) = make_float4(xy[0], xy[1], xy[2], xy[3]);
                    *reinterpret_cast<float4*>(o + shadow + 2 * C::PSTRIDE) = make_float4(yy[0], yy[1], yy[2], yy[3]);
                }
            };
            const int t0i = pw * 32 + lane;
            if (t0i < ntask) task(t0i);
            if (t0i + 32 * hs::NPROD < ntask) task(t0i + 32 * hs::NPROD);
            if (((pw + c.slot) & (hs::NPROD - 1)) == 0 && 2 * 32 * hs::NPROD + lane < ntask) task(2 * 32 * hs::NPROD + lane);   // warp-uniform
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full(c.slot));
            if (fuse_next) {
                // ScaleRotInvSIFT.py:109-115 -> cv2.resize at an exact halving -> INTER_AREA 2x2 mean, from the staged tile.
                // A run over bands [ka, kb] owns image rows [16 ka, 16 (kb + 1)); its chunk kc holds image rows
                // 16 kc - R - 1 .. 16 kc + 16 - R and emits the owned rows among [16 kc - EMIT, 16 kc - EMIT + 16)
                // (an even start, inside the tile).  A task = two rows x 4 columns -> two output pixels; 128 tasks.
                float* dstl = P.pyr + (size_t)c.b * P.pyr_stride + nx.img_off;
                const int ylo = (c.extra == 0 && first_in_run) ? kc * hs::BH : kc * hs::BH - C::EMIT;
                const int yhi = c.extra ? kc * hs::BH : kc * hs::BH - C::EMIT + hs::BH;      // exclusive
                const int i = pw * 32 + lane;
                const int pr = i / (hs::SW / 4), oq = i - pr * (hs::SW / 4);
                const int y = kc * hs::BH - C::EMIT + 2 * pr;                          // image row of the pair's upper row
                const int oy = y >> 1, gx = (x0 >> 1) + 2 * oq;
                if (y >= ylo && y < yhi && oy < nx.H && gx < nx.W) {
                    const float* pp = tile + (y - (kc * hs::BH - C::R - 1)) * C::IPITCH + C::RA + 4 * oq;
                    const float4 u = *reinterpret_cast<const float4*>(pp);
                    const float4 d = *reinterpret_cast<const float4*>(pp + C::IPITCH);
                    const float o0 = __fmul_rn(__fadd_rn(__fadd_rn(u.x, u.y), __fadd_rn(d.x, d.y)), 0.25f);
                    const float o1 = __fmul_rn(__fadd_rn(__fadd_rn(u.z, u.w), __fadd_rn(d.z, d.w)), 0.25f);
                    float* o = dstl + (size_t)oy * nx.W + gx;
                    if (gx + 1 < nx.W && (nx.W & 1) == 0) *reinterpret_cast<float2*>(o) = make_float2(o0, o1);
                    else { o[0] = o0; if (gx + 1 < nx.W) o[1] = o1; }
                }
                __syncwarp();
            }
            if (lane == 0) mbar_arrive(bar_iempty(c.st));
#pragma unroll
            for (int i = 0; i < hs::NGROUP; ++i) if (c.n < g.n1) advance(c);
        }
    }
}
