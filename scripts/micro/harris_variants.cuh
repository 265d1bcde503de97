// harris_variants.cuh -- NOT part of libsfmb200.so.
//
// Two measured dead ends of the Harris kernel, kept for reference only (round 1):
//   k_harris_p   persistent CTAs walking runs of tiles, cp.async tile prefetch under the window stage
//                (0.96 ms per 32 x 1080p against 0.80 for the shipped one-tile-per-CTA kernel);
//   k_harris_ws  warp-specialised persistent kernel, TMA producer warps feeding window warps through
//                mbarriers (0.80 ms: the same; the window stage, not the phase serialisation, is the limit).
// They compile inside sfmfromscratch_b200/csrc/extract.cu after the HarrisCfg / harris_* stage helpers
// (that is where they were cut from); they are not built, tested or shipped.
// ---- persistent variant (rows 16-byte aligned: W % 4 == 0).  Each CTA walks a
// contiguous run of tiles of one level across the whole batch.  The haloed
// image tile arrives by cp.async (16-byte chunks, zero-filled outside the image
// -- BORDER_CONSTANT for free); the copy for tile i+1 is issued as soon as the
// products of tile i are in the planes, so it overlaps the 49-tap window
// stage.  The radix histogram lives in shared memory for the CTA's lifetime
// and is flushed only when the run crosses into the next image.
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

template <int G, int TH>
__global__ void __launch_bounds__(HarrisCfg<G, TH>::THREADS, (TH == 64 ? 2 : 3))
k_harris_p(const __grid_constant__ ExtractPlan P, const __grid_constant__ GaussWeights gw, int l, int tiles_x,
           int tiles_y, int tiles_per_cta) {
    using C = HarrisCfg<G, TH>;
    constexpr int NT_ = C::THREADS;
    constexpr int V = C::IPITCH / 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_img = reinterpret_cast<float*>(smem_raw);
    float* s_prod = s_img + C::IMG_WORDS;
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_prod + C::PROD_WORDS);
    const LevelInfo& lv = P.lv[l];
    const int H = lv.H, W = lv.W, t = threadIdx.x;
    const int per_img = tiles_x * tiles_y;
    const int n_tiles = per_img * P.B;
    const int first = blockIdx.x * tiles_per_cta;
    const int last = min(first + tiles_per_cta, n_tiles);
    if (first >= last) return;

    auto issue = [&](int tile) {
        const int b = tile / per_img, rem = tile - b * per_img;
        const int tyi = rem / tiles_x, txi = rem - tyi * tiles_x;
        const int ix0 = txi * HT - C::RA, iy0 = tyi * TH - C::R - 1;
        const float* img = level_image(P, b, l);
        for (int i = t; i < V * C::IH; i += NT_) {
            const int ty = i / V, tv = i - ty * V;
            const int gy = iy0 + ty, gx = ix0 + 4 * tv;
            const bool ok = (gy >= 0 && gy < H && gx >= 0 && gx < W);     // W % 4 == 0: chunks are all-in or all-out
            cp_async16(s_img + ty * C::IPITCH + 4 * tv, img + (ok ? (size_t)gy * W + gx : 0), ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    for (int i = t; i < SFM_HIST1_BINS; i += NT_) s_hist[i] = 0;
    int hist_b = first / per_img;
    issue(first);
    for (int tile = first; tile < last; ++tile) {
        const int b = tile / per_img, rem = tile - b * per_img;
        const int tyi = rem / tiles_x, txi = rem - tyi * tiles_x;
        const int x0 = txi * HT, y0 = tyi * TH;
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();                                          // tile landed; previous tile's histogram updates done
        if (b != hist_b) {                                        // crossed into the next image: flush
            uint32_t* gh = P.hist1 + (size_t)(hist_b * P.L + l) * SFM_HIST1_BINS;
            for (int i = t; i < SFM_HIST1_BINS; i += NT_) {
                const uint32_t c = s_hist[i];
                if (c) { atomicAdd(gh + i, c); s_hist[i] = 0; }
            }
            hist_b = b;
            __syncthreads();
        }
        const bool interior = (x0 - C::RA >= 0) && (x0 - C::RA + C::IPITCH <= W) && (y0 - C::R - 1 >= 0) &&
                              (y0 - C::R - 1 + C::IH <= H);
        if (interior) harris_products<G, TH, true>(s_img, s_prod, x0, y0, H, W);
        else harris_products<G, TH, false>(s_img, s_prod, x0, y0, H, W);
        __syncthreads();                                          // planes complete, image tile free
        if (tile + 1 < last) issue(tile + 1);                     // overlaps the window stage below
        float r[2][8];
        harris_window<G, TH, false>(s_prod, gw, P.alpha, r);
        float* Rout = P.R + (size_t)b * P.r_stride + lv.r_off;
        if (interior) harris_store<G, TH, true>(r, Rout, s_hist, x0, y0, H, W);
        else harris_store<G, TH, false>(r, Rout, s_hist, x0, y0, H, W);
    }
    __syncthreads();
    uint32_t* gh = P.hist1 + (size_t)(hist_b * P.L + l) * SFM_HIST1_BINS;
    for (int i = t; i < SFM_HIST1_BINS; i += NT_) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(gh + i, c);
    }
}

// ---- warp-specialised persistent variant (development knob SFM_HARRIS_VARIANT=5; rows 16-byte aligned).
//
// Written to test whether the one-tile-per-CTA kernel loses its time to phase serialisation (tile
// load, products, window, histogram/store behind barriers).  It does not: this variant measures the
// same 0.80 ms per 32 x 1080p.  A micro-benchmark of the window stage alone (scripts/micro/
// window_rate.cu) tops out at ~70-77 % of the FMA-pipe peak at any occupancy -- the packed FFMA2
// stream with a fresh scalar and accumulator pair per instruction is limited by operand delivery,
// not by latency -- so hiding the other phases cannot buy more than it already does in the
// shipped kernel.  Kept as the TMA reference implementation.  A CTA is two groups of four warps:
//   producers  issue the TMA load of tile i+1 (cp.async.bulk.tensor, 3-D map [B][H][W]; elements
//              outside the image arrive as zeros = BORDER_CONSTANT), emit the next pyramid level from
//              the tile, and turn tile i into the three product planes (ring of 2);
//   consumers  run the window chains of tile i off the planes, store R, and count it into a
//              histogram that stays in shared memory until the CTA's run leaves the image.
// full/empty mbarriers connect them; each CTA walks a contiguous run of tiles of the batch.
template <int G>
struct HarrisWs {
    using C = HarrisCfg<G, 32>;
    static constexpr int IMG_BYTES = (C::IPITCH * C::IH * 4 + 127) & ~127;
    static constexpr int PROD_BYTES = (C::PROD_WORDS * 4 + 127) & ~127;
    static constexpr int OFF_IMG = 0;
    static constexpr int OFF_PROD = 2 * IMG_BYTES;
    static constexpr int OFF_HIST = OFF_PROD + 2 * PROD_BYTES;
    static constexpr int OFF_BAR = OFF_HIST + SFM_HIST1_BINS * 4;
    static constexpr int SMEM = OFF_BAR + 64 + 128;             // + slack for the 128-byte base alignment
};

template <int G>
__global__ void __launch_bounds__(256, 2)
k_harris_ws(const __grid_constant__ ExtractPlan P, const __grid_constant__ GaussWeights gw,
            const __grid_constant__ CUtensorMap tmap, int l, int tiles_x, int tiles_y, int tiles_per_cta, int fuse_next) {
    using namespace sfm_tma;
    using C = HarrisCfg<G, 32>;
    using WS = HarrisWs<G>;
    constexpr int TH = 32;
    extern __shared__ unsigned char smem_ws[];
    unsigned char* sb = smem_ws + ((128u - (smem_u32(smem_ws) & 127u)) & 127u);
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(sb + WS::OFF_HIST);
    const uint32_t bar0 = smem_u32(sb + WS::OFF_BAR);
    // barriers: img_full[2] @0,8; prod_full[2] @16,24; prod_empty[2] @32,40
    const LevelInfo& lv = P.lv[l];
    const int H = lv.H, W = lv.W;
    const int per_img = tiles_x * tiles_y;
    const int n_tiles = per_img * P.B;
    const int first = blockIdx.x * tiles_per_cta;
    const int last = min(first + tiles_per_cta, n_tiles);
    if (first >= last) return;
    const int n_my = last - first;
    const int t = threadIdx.x;
    if (t == 0) {
        mbar_init(bar0 + 0, 1); mbar_init(bar0 + 8, 1);
        mbar_init(bar0 + 16, 1); mbar_init(bar0 + 24, 1);
        mbar_init(bar0 + 32, 128); mbar_init(bar0 + 40, 128);
        mbar_fence_init();
    }
    for (int i = t; i < SFM_HIST1_BINS; i += 256) s_hist[i] = 0;
    __syncthreads();

    if (t >= 128) {
        // ------------------------------------------------------------ producers
        const int tp = t - 128;
        auto issue = [&](int i) {                                    // one thread: tile first+i -> image slot i & 1
            const int tile = first + i;
            const int b = tile / per_img, rem = tile - b * per_img;
            const int tyi = rem / tiles_x, txi = rem - tyi * tiles_x;
            const uint32_t bar = bar0 + 8 * (i & 1);
            mbar_expect_tx(bar, (uint32_t)(C::IPITCH * C::IH * 4));
            tma_load_3d(smem_u32(sb + WS::OFF_IMG + (i & 1) * WS::IMG_BYTES), &tmap, bar, txi * HT - C::RA, tyi * TH - C::R - 1, b);
        };
        if (tp == 0) issue(0);
        for (int i = 0; i < n_my; ++i) {
            const int s = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            // image slot s^1 was last read by iteration i-1, which ended with the producers' barrier
            if (tp == 0 && i + 1 < n_my) issue(i + 1);
            const int tile = first + i;
            const int b = tile / per_img, rem = tile - b * per_img;
            const int tyi = rem / tiles_x, txi = rem - tyi * tiles_x;
            const int x0 = txi * HT, y0 = tyi * TH;
            const float* s_img = reinterpret_cast<const float*>(sb + WS::OFF_IMG + s * WS::IMG_BYTES);
            float* s_prod = reinterpret_cast<float*>(sb + WS::OFF_PROD + s * WS::PROD_BYTES);
            mbar_wait(bar0 + 8 * s, ph);                             // tile landed
            if (fuse_next) {                                         // ScaleRotInvSIFT.py:109-115: exact 2x2 mean
                const LevelInfo& nx = P.lv[l + 1];
                float* dst = P.pyr + (size_t)b * P.pyr_stride + nx.img_off;
                for (int q = tp; q < (HT / 2) * (TH / 2); q += 128) {
                    const int oy = q / (HT / 2), ox = q - oy * (HT / 2);
                    const int gy = y0 / 2 + oy, gx = x0 / 2 + ox;
                    if (gy < nx.H && gx < nx.W) {
                        const float* p = s_img + (2 * oy + C::R + 1) * C::IPITCH + 2 * ox + C::RA;
                        const float top = __fadd_rn(p[0], p[1]);
                        const float bot = __fadd_rn(p[C::IPITCH], p[C::IPITCH + 1]);
                        dst[(size_t)gy * nx.W + gx] = __fmul_rn(__fadd_rn(top, bot), 0.25f);
                    }
                }
            }
            mbar_wait_relaxed(bar0 + 32 + 8 * s, ph ^ 1u);           // consumers are done with plane slot s
            const bool interior = (x0 - C::R >= 0) && (x0 - C::R + C::PCH * 4 <= W) && (y0 - C::R >= 0) && (y0 - C::R + C::PH <= H);
            if (interior) harris_products<G, TH, true>(s_img, s_prod, x0, y0, H, W, tp);
            else harris_products<G, TH, false>(s_img, s_prod, x0, y0, H, W, tp);
            bar_sync(1, 128);                                        // planes complete, image slot s free
            if (tp == 0) mbar_arrive(bar0 + 16 + 8 * s);
        }
    } else {
        // ------------------------------------------------------------ consumers
        int hist_b = first / per_img;
        for (int i = 0; i < n_my; ++i) {
            const int s = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            const int tile = first + i;
            const int b = tile / per_img, rem = tile - b * per_img;
            const int tyi = rem / tiles_x, txi = rem - tyi * tiles_x;
            const int x0 = txi * HT, y0 = tyi * TH;
            if (b != hist_b) {                                       // the run crossed into the next image: flush
                bar_sync(2, 128);
                uint32_t* gh = P.hist1 + (size_t)(hist_b * P.L + l) * SFM_HIST1_BINS;
                for (int q = t; q < SFM_HIST1_BINS; q += 128) {
                    const uint32_t c = s_hist[q];
                    if (c) { atomicAdd(gh + q, c); s_hist[q] = 0; }
                }
                hist_b = b;
                bar_sync(2, 128);
            }
            const float* s_prod = reinterpret_cast<const float*>(sb + WS::OFF_PROD + s * WS::PROD_BYTES);
            mbar_wait(bar0 + 16 + 8 * s, ph);                        // planes of this tile are complete
            float r[2][8];
            harris_window<G, TH, true>(s_prod, gw, P.alpha, r, t);
            mbar_arrive(bar0 + 32 + 8 * s);                          // this thread no longer reads plane slot s
            float* Rout = P.R + (size_t)b * P.r_stride + lv.r_off;
            const bool full = (x0 + HT <= W) && (y0 + TH <= H);
            if (full) harris_store<G, TH, true>(r, Rout, s_hist, x0, y0, H, W, t);
            else harris_store<G, TH, false>(r, Rout, s_hist, x0, y0, H, W, t);
        }
        bar_sync(2, 128);
        uint32_t* gh = P.hist1 + (size_t)(hist_b * P.L + l) * SFM_HIST1_BINS;
        for (int q = t; q < SFM_HIST1_BINS; q += 128) {
            const uint32_t c = s_hist[q];
            if (c) atomicAdd(gh + q, c);
        }
    }
}

