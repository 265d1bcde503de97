// match.cu -- NN-ratio matcher (NNRatioFeatureMatcher.py:8-60) on sm_100a.
//
//   k_match_setup    set / pair tables
//   k_match_prep     float32 -> fp16 copy, |b|^2, rounding-residual norms
//   k_match_tc       (match_tc.cu) tcgen05 fp16 GEMM tiles, fused top-4 groups
//   k_match_recheck  exact float32 distances (numpy summation order) of the
//                    candidate groups + error-bound certificate
//   k_match_rescan   rows the certificate rejected: float32 dot-product scan,
//                    exact arithmetic for the columns within the bound
//   k_match_exact    exact tile scan of every row (SFM_MATCH_EXACT mode)
//   k_match_merge    partial results of either scan
//   k_match_emit     d1 > 0, d0/d1 <= thr            (:46-51)
//   k_match_sort     order by confidence             (:56-58)
//
// "Exact" means the reference's float32 arithmetic: t = a - b, t * t, numpy's
// pairwise sum over 128 contiguous values (8 strided accumulators and the tree
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))), sqrt, divide.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "match.cuh"

int launch_match_tc(SfmCtx* ctx, cudaStream_t st, const MatchPlan& P);   // match_tc.cu

// ------------------------------------------------------------------ setup / prep

__global__ void k_match_setup(const __grid_constant__ MatchPlan P) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (P.desc == nullptr) {
        if (t == 0) {
            P.set_ptr[0] = P.f1; P.set_ptr[1] = P.f2;
            P.set_cnt[0] = P.n1; P.set_cnt[1] = P.n2;
            P.pairs[0] = 0; P.pairs[1] = 1;
        }
    } else {
        if (t < P.n_sets) {
            P.set_ptr[t] = P.desc + (size_t)t * P.nmax * SFM_DESC_DIM;
            int c = P.counts_in[t];
            P.set_cnt[t] = c < 0 ? 0 : (c > P.nmax ? P.nmax : c);
        }
        if (t < P.n_pairs) {
            // set ids outside [0, n_sets) would index the set tables out of bounds: such a pair is redirected to an
            // empty set pair (no rows, no matches) and counted so the caller can see it (stats are optional)
            const int a = P.pairs_in[2 * t], b = P.pairs_in[2 * t + 1];
            const bool ok = a >= 0 && a < P.n_sets && b >= 0 && b < P.n_sets;
            P.pairs[2 * t] = ok ? a : P.n_sets;
            P.pairs[2 * t + 1] = ok ? b : P.n_sets;
        }
        if (t == 0) { P.set_ptr[P.n_sets] = P.desc; P.set_cnt[P.n_sets] = 0; }    // the empty set bad pair ids point at
    }
}

// Warp per row: fp16 copy for the tensor-core pass, |b|^2 and the norms the
// re-check's error bound needs.
__global__ void __launch_bounds__(256) k_match_prep(const __grid_constant__ MatchPlan P) {
    __shared__ float s_max[3][8];
    const int s = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + warp;
    const int cnt = P.set_cnt[s];
    float hs = 0.f, rs = 0.f, nbv = 0.f;
    if (r < P.nmax_pad) {
        const size_t orow = (size_t)s * P.nmax_pad + r;
        __half2* out = reinterpret_cast<__half2*>(P.h16 + orow * SFM_DESC_DIM + lane * 4);
        if (r >= cnt) {
            out[0] = __floats2half2_rn(0.f, 0.f);
            out[1] = __floats2half2_rn(0.f, 0.f);
            if (lane == 0) { P.nb[orow] = MT_SENTINEL; P.hatn[orow] = 0.f; P.resn[orow] = 0.f; }
        } else {
            const float4 v = *reinterpret_cast<const float4*>(P.set_ptr[s] + (size_t)r * SFM_DESC_DIM + lane * 4);
            const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
            out[0] = h01; out[1] = h23;
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            float nb = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            float hn = f01.x * f01.x + f01.y * f01.y + f23.x * f23.x + f23.y * f23.y;
            float e0 = v.x - f01.x, e1 = v.y - f01.y, e2 = v.z - f23.x, e3 = v.w - f23.y;
            float rn = e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
            for (int o = 16; o > 0; o >>= 1) {
                nb += __shfl_xor_sync(0xffffffffu, nb, o);
                hn += __shfl_xor_sync(0xffffffffu, hn, o);
                rn += __shfl_xor_sync(0xffffffffu, rn, o);
            }
            hs = sqrtf(hn); rs = sqrtf(rn); nbv = nb;
            if (lane == 0) { P.nb[orow] = nb; P.hatn[orow] = hs; P.resn[orow] = rs; }
        }
    }
    // one atomic per CTA and statistic (per-row atomics on three addresses serialise in L2)
    if (lane == 0) { s_max[0][warp] = hs; s_max[1][warp] = rs; s_max[2][warp] = nbv; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float m = 0.f;
        for (int w = 0; w < 8; ++w) m = fmaxf(m, s_max[threadIdx.x][w]);
        // non-negative floats order like their bit patterns
        if (m > 0.f) atomicMax(reinterpret_cast<int*>(P.setmax + 4 * s + threadIdx.x), __float_as_int(m));
    }
}

// ------------------------------------------------------------------ exact arithmetic

struct Top2 { float d0; int i0; float d1; };

__device__ __forceinline__ Top2 top2_merge(Top2 a, Top2 b) {
    const bool bwins = (b.d0 < a.d0) || (b.d0 == a.d0 && b.i0 >= 0 && (a.i0 < 0 || b.i0 < a.i0));
    if (bwins) { Top2 t = a; a = b; b = t; }
    a.d1 = fminf(a.d1, b.d0);
    return a;
}

__device__ __forceinline__ void top2_push(Top2& s, float d, int j) {
    if (d < s.d0) { s.d1 = s.d0; s.d0 = d; s.i0 = j; }
    else if (d < s.d1) { s.d1 = d; }
}

// The same for columns that arrive in no particular order (the re-check visits groups by approximate
// key): among columns at exactly the nearest distance the lowest index wins, as in the ordered scans.
// Only observable with ratio_threshold >= 1, where tied rows (ratio == 1) are emitted.
__device__ __forceinline__ void top2_push_any_order(Top2& s, float d, int j) {
    if (d < s.d0 || (d == s.d0 && j < s.i0)) { s.d1 = s.d0; s.d0 = d; s.i0 = j; }
    else if (d < s.d1) { s.d1 = d; }
}

// ------------------------------------------------------------------ re-check of tensor-core candidates

// Eight lanes per query row, four rows per warp.  Entries of the row's
// candidate lists are visited in increasing approximate key; each visit
// evaluates the 4 columns of the group exactly.  The gather is bound by L1TEX
// wavefronts (one per 128-byte line an instruction touches), so the team reads
// every column as four full lines (lane k takes the k-th 16 bytes), forms the
// squares (a - b)^2 where they land, and passes them through a padded
// shared-memory staging row to the summation layout: lane (c, h) owns column c
// of the group and numpy's accumulators r[4h .. 4h+3], adds its sixteen
// float4 in order, and the tree ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) closes
// with one shuffle.  The row is
// certified when every unvisited or untracked column is provably farther than
// the exact second-nearest found: approx_key + |a|^2 - E > d1^2, with E a
// rigorous bound on |approx - exact| (fp16 rounding of both operands by
// Cauchy-Schwarz on the residual norms, accumulation, index packing, and the
// float32 rounding of the exact sum itself).  Uncertified rows go to the exact
// scan.  NV = candidate entries per lane (ceil(n_lists * 4 / 8)).
constexpr int RC_ROWS = 256;           // most rows per CTA (phase 1 is thread per row); small batches use fewer, for more CTAs
constexpr int RC_TEAMS = 32;           // phase 2: 8-lane teams over the surviving rows
constexpr int RC_SQ_STRIDE = SFM_DESC_DIM + 8;   // staging row of squares: 544 bytes keeps the 4 columns on distinct banks

// Phase 1 (thread per row): the row's error bound and the ratio prune.  k0 <= k1 are the two
// smallest group keys, so the nearest exact distance is at least L0 = k0 + |a|^2 - E and the
// second-nearest (two distinct columns exist that close) at most U1 = k1 + |a|^2 + E.  When
// sqrt(L0 / U1) exceeds the threshold with room for the two float32 roundings of sqrt and divide,
// the reference's ratio test rejects the row whatever the exact values are: nothing is gathered
// (d1 = 0 makes k_match_emit skip the row).  Surviving rows are compacted so that phase 2 runs
// with full teams.
// Phase 2 (8 lanes per row): as described above.
template <int NV, int GRP>
__global__ void __launch_bounds__(256, 2) k_match_recheck(const __grid_constant__ MatchPlan P, int rows_per_cta) {
    using G = MtG<GRP>;
    extern __shared__ __align__(16) unsigned char rc_smem[];
    float (*s_sq)[MT_SUB][RC_SQ_STRIDE] = reinterpret_cast<float (*)[MT_SUB][RC_SQ_STRIDE]>(rc_smem);
    __shared__ double s_ebase[RC_ROWS], s_na[RC_ROWS];
    __shared__ int16_t s_rows[RC_ROWS];
    __shared__ int s_n, s_visited;
    const int p = P.p0 + blockIdx.y;
    const int qa = P.pairs[2 * p], qb = P.pairs[2 * p + 1];
    const int n1 = P.set_cnt[qa], n2 = P.set_cnt[qb];
    if (n2 < 2) return;                                    // the pair emits nothing
    const int E = P.n_lists * MT_TOPK;
    const double q_rel = 1.1 / 8192.0 * 2.0;               // packing drops 10 mantissa bits
    if (threadIdx.x == 0) { s_n = 0; s_visited = 0; }
    __syncthreads();
    {
        const int row = blockIdx.x * rows_per_cta + threadIdx.x;
        if (row < n1 && (int)threadIdx.x < rows_per_cta) {
            const uint4* l4 = reinterpret_cast<const uint4*>(P.cands + ((size_t)p * P.nmax_pad + row) * (size_t)E);
            float k0 = INFINITY, k1 = INFINITY;
            for (int e = 0; e < E / 4; ++e) {
                const uint4 q = l4[e];
                const uint32_t pk[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float v = __uint_as_float(pk[i] & ~MT_IDX_MASK);
                    if (v < MT_INVALID) {
                        if (v < k0) { k1 = k0; k0 = v; } else if (v < k1) k1 = v;
                    }
                }
            }
            const size_t arow = (size_t)qa * P.nmax_pad + row;
            const double na = (double)P.nb[arow];
            const double hat_a = (double)P.hatn[arow], res_a = (double)P.resn[arow];
            const double mh = (double)P.setmax[4 * qb + 0], mr = (double)P.setmax[4 * qb + 1], mnb = (double)P.setmax[4 * qb + 2];
            const double e_fp16 = 2.0 * (hat_a * mr + res_a * mh + res_a * mr);
            const double e_acc = 2.0 * hat_a * mh * (1.0 / 262144.0);                   // 2^-18
            const double e_ref = (na + mnb + 2.0 * sqrt(na * mnb)) * (1.0 / 524288.0);  // 2^-19 * dmax^2
            const double e_nrm = (na + mnb) * (1.0 / 131072.0);                         // 2^-17
            const double e_base = 1.1 * (e_fp16 + e_acc + e_ref + e_nrm) + 1e-12;
            bool pruned = false;
            if (k1 < INFINITY && P.thr >= 0.0f && !P.no_prune) {
                const double em = e_base + q_rel * fmax(fabs((double)k0), fabs((double)k1));
                const double L0 = (double)k0 + na - em, U1 = (double)k1 + na + em;
                pruned = L0 > 0.0 && L0 * (1.0 - 4e-6) > (double)P.thr * (double)P.thr * U1;
            }
            if (pruned) {
                const size_t o = (size_t)p * P.nmax + row;
                P.res_idx[o] = -1; P.res_d0[o] = 0.f; P.res_d1[o] = 0.f;
            } else {
                const int slot = atomicAdd(&s_n, 1);
                s_rows[slot] = (int16_t)threadIdx.x; s_ebase[slot] = e_base; s_na[slot] = na;
            }
        }
    }
    __syncthreads();
    const int n_live = s_n;
    const int team = threadIdx.x >> 3, j = threadIdx.x & 7;
    const int c = j & 3, h = j >> 2;
    const unsigned tmask = 0xffu << (threadIdx.x & 24);
    const float* B = P.set_ptr[qb];
    int visited = 0;
    for (int slot = team; slot < n_live; slot += RC_TEAMS) {
        const int row = blockIdx.x * rows_per_cta + (int)s_rows[slot];
        const double e_base = s_ebase[slot], na = s_na[slot];
        float4 a4[4];                                      // elements 32 i + 4 j .. + 3 of the query row
        {
            const float4* A4 = reinterpret_cast<const float4*>(P.set_ptr[qa] + (size_t)row * SFM_DESC_DIM);
#pragma unroll
            for (int i = 0; i < 4; ++i) a4[i] = A4[j + 8 * i];
        }
        const uint32_t* lists = P.cands + ((size_t)p * P.nmax_pad + row) * (size_t)E;
        float val[NV];
        uint32_t code[NV];
        float Lmin = INFINITY;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int e = j + 8 * i;
            val[i] = INFINITY; code[i] = 0;
            if (e < E) {
                const uint32_t pk = lists[e];
                const float v = __uint_as_float(pk & ~MT_IDX_MASK);
                if (v < MT_INVALID) {
                    val[i] = v;
                    code[i] = ((uint32_t)(e / MT_TOPK) << MT_IDX_BITS) | (pk & MT_IDX_MASK);
                    if ((e % MT_TOPK) == MT_TOPK - 1) Lmin = fminf(Lmin, v);
                }
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) Lmin = fminf(Lmin, __shfl_xor_sync(tmask, Lmin, o, 8));

        Top2 best = {INFINITY, -1, INFINITY};

        // smallest unvisited entry of the team: its value (INFINITY when none is left), owner lane and slot
        auto find_min = [&](float& wv, int& wl, int& bi, uint32_t& bc) {
            float bv = INFINITY; bc = 0; bi = -1;
#pragma unroll
            for (int i = 0; i < NV; ++i) if (val[i] < bv) { bv = val[i]; bc = code[i]; bi = i; }
            wv = bv;
            wl = (bi >= 0) ? j : 64;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(tmask, wv, o, 8);
                const int ol = __shfl_xor_sync(tmask, wl, o, 8);
                if (ov < wv || (ov == wv && ol < wl)) { wv = ov; wl = ol; }
            }
        };
        auto consume = [&](int wl, int bi, uint32_t bc) -> uint32_t {
            const uint32_t wc = __shfl_sync(tmask, bc, wl & 7, 8);
            if (j == wl) {
#pragma unroll
                for (int i = 0; i < NV; ++i) if (i == bi) val[i] = INFINITY;
            }
            return wc;
        };
        auto group_col = [&](uint32_t wc) -> int {
            const int list = (int)(wc >> MT_IDX_BITS);
            const int split = list >> 1, half = list & 1;
            const int tile = split * P.tiles_per_split + (int)((wc & MT_IDX_MASK) >> G::GROUP_BITS);
            return tile * MT_COLS + half * (MT_COLS / 2) + (int)(wc & (G::GROUPS_PER_HALF - 1)) * G::GROUP;
        };
        // 4 columns as full lines: v[4 cc + i] = elements 32 i + 4 j .. + 3 of column col0 + cc
        auto group_load = [&](int col0, float4 (&v)[16]) {
#pragma unroll
            for (int cc = 0; cc < MT_SUB; ++cc) {
                const float4* bp = reinterpret_cast<const float4*>(B + (size_t)min(col0 + cc, n2 - 1) * SFM_DESC_DIM) + j;
#pragma unroll
                for (int i = 0; i < 4; ++i) v[4 * cc + i] = bp[8 * i];
            }
        };
        // squares (a - b)^2 of the loaded columns into the team's staging rows, in the columns' own element order
        auto stage_squares = [&](const float4 (&v)[16]) {
#pragma unroll
            for (int cc = 0; cc < MT_SUB; ++cc)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 b = v[4 * cc + i], a = a4[i];
                    float4 q;
                    float t;
                    t = __fsub_rn(a.x, b.x); q.x = __fmul_rn(t, t);
                    t = __fsub_rn(a.y, b.y); q.y = __fmul_rn(t, t);
                    t = __fsub_rn(a.z, b.z); q.z = __fmul_rn(t, t);
                    t = __fsub_rn(a.w, b.w); q.w = __fmul_rn(t, t);
                    *reinterpret_cast<float4*>(&s_sq[team][cc][32 * i + 4 * j]) = q;
                }
            __syncwarp(tmask);
        };
        // numpy's summation of the staged squares of column c; every lane of the team gets its column's squared distance
        auto sum_staged = [&]() -> float {
            const float4* sp = reinterpret_cast<const float4*>(&s_sq[team][c][4 * h]);
            float4 x = sp[0];
#pragma unroll
            for (int m = 1; m < 16; ++m) {
                const float4 y = sp[2 * m];
                x.x = __fadd_rn(x.x, y.x); x.y = __fadd_rn(x.y, y.y);
                x.z = __fadd_rn(x.z, y.z); x.w = __fadd_rn(x.w, y.w);
            }
            __syncwarp(tmask);                                                  // staging rows are reused by the next sub-group
            const float s4 = __fadd_rn(__fadd_rn(x.x, x.y), __fadd_rn(x.z, x.w));   // (r0+r1)+(r2+r3) or (r4+r5)+(r6+r7)
            return __fadd_rn(s4, __shfl_xor_sync(tmask, s4, 4, 8));
        };
        auto push_sub = [&](float d2, int col0) {
#pragma unroll
            for (int tm = 0; tm < MT_SUB; ++tm) {
                const float dd = __shfl_sync(tmask, d2, tm, 8);
                if (col0 + tm < n2) top2_push_any_order(best, dd, col0 + tm);
            }
        };
        // NG candidate groups, MT_SUB columns at a time; the next sub-group's loads are issued once the
        // current one's registers are free and fly under its summation
        auto visit = [&](const int (&cols)[2], auto ng_tag) {
            constexpr int SPG = G::GROUP / MT_SUB;
            constexpr int NS = decltype(ng_tag)::value * SPG;
            float4 v[16];
            group_load(cols[0], v);
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                const int cur = cols[i / SPG] + (i % SPG) * MT_SUB;
                stage_squares(v);
                if (i + 1 < NS) group_load(cols[(i + 1) / SPG] + ((i + 1) % SPG) * MT_SUB, v);
                push_sub(sum_staged(), cur);
            }
            visited += decltype(ng_tag)::value;
        };

        // The two best entries are evaluated together: the second one is needed in practice anyway,
        // because one group rarely certifies a row.
        float wv; int wl, bi; uint32_t bc;
        find_min(wv, wl, bi, bc);
        if (wl < 8) {
            int cols[2];
            cols[0] = group_col(consume(wl, bi, bc));
            find_min(wv, wl, bi, bc);
            if (wl < 8) {
                cols[1] = group_col(consume(wl, bi, bc));
                visit(cols, std::integral_constant<int, 2>{});
            } else {
                cols[1] = cols[0];
                visit(cols, std::integral_constant<int, 1>{});
            }
            for (;;) {
                find_min(wv, wl, bi, bc);
                if (wl >= 8) break;                               // nothing left
                const double bound = (double)wv + na - (e_base + q_rel * fabs((double)wv));
                if (bound > (double)best.d1) break;               // the rest cannot matter
                cols[0] = cols[1] = group_col(consume(wl, bi, bc));
                visit(cols, std::integral_constant<int, 1>{});
            }
        }
        const bool certified =
            (Lmin == INFINITY) ||
            ((double)Lmin + na - (e_base + q_rel * fabs((double)Lmin)) > (double)best.d1);
        // An uncertified row whose ratio test is decided anyway needs no rescan.  Unseen columns are at
        // least lb = Lmin + |a|^2 - E away, so the true nearest is >= min(d0, lb) and the true
        // second-nearest <= d1: when even that ratio exceeds the threshold, or when two columns at
        // distance exactly 0 were found (d1 = 0 fails the reference's `> 0`), the row emits nothing.
        bool rejected = false;
        if (!certified && P.thr >= 0.0f && !P.no_prune) {
            const double lb = fmax((double)Lmin + na - (e_base + q_rel * fabs((double)Lmin)), 0.0);
            const double lo0 = fmin((double)best.d0, lb);
            rejected = best.d1 == 0.0f ||
                       (best.d1 < INFINITY && lo0 > 0.0 && lo0 * (1.0 - 4e-6) > (double)P.thr * (double)P.thr * (double)best.d1);
        }
        if (j == 0) {
            const size_t o = (size_t)p * P.nmax + row;
            if (rejected) {
                P.res_idx[o] = -1; P.res_d0[o] = 0.f; P.res_d1[o] = 0.f;
            } else if (certified) {
                P.res_idx[o] = best.i0; P.res_d0[o] = best.d0; P.res_d1[o] = best.d1;
            } else {
                const int pos = atomicAdd(&P.flag_cnt[p], 1);
                P.flag_rows[(size_t)p * P.nmax + pos] = row;
                P.res_d1[o] = best.d1;                            // upper bound of the true second-nearest: the rescan's filter
            }
        }
    }
    if (P.stats) {
        if (j == 0 && visited) atomicAdd(&s_visited, visited);
        __syncthreads();
        if (threadIdx.x == 0 && s_visited) atomicAdd(&P.stats[2 * p + 1], s_visited);
    }
}

static int launch_match_recheck(SfmCtx* ctx, cudaStream_t s, const MatchPlan& P) {
    // rows per CTA: 256 when that still gives a few waves of CTAs, fewer for small batches (the kernel is
    // a chain of dependent gathers per row: with one wave the slowest CTA sets the time)
    int rows = RC_ROWS;
    {
        constexpr int min_rows = 32, waves = 8;                 // measured on B200 (scripts/time_match.py)
        while (rows > min_rows && (long long)ceil_div(P.nmax, rows) * P.pn < (long long)waves * ctx->sm_count) rows >>= 1;
    }
    const dim3 grid(ceil_div(P.nmax, rows), P.pn);
    const int nv = ceil_div(P.n_lists * MT_TOPK, 8);
    constexpr int smem = RC_TEAMS * MT_SUB * RC_SQ_STRIDE * (int)sizeof(float);
#define RC_GO(NV)                                                                                          \
    do {                                                                                                   \
        if (P.group == 4) {                                                                                \
            SFM_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_match_recheck<NV, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            SFM_LAUNCH(ctx, s, "k_match_recheck", k_match_recheck<NV, 4><<<grid, 256, smem, s>>>(P, rows)); \
        } else {                                                                                           \
            SFM_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_match_recheck<NV, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            SFM_LAUNCH(ctx, s, "k_match_recheck", k_match_recheck<NV, 8><<<grid, 256, smem, s>>>(P, rows)); \
        }                                                                                                  \
    } while (0)
    if (nv <= 1) RC_GO(1);
    else if (nv <= 2) RC_GO(2);
    else if (nv <= 4) RC_GO(4);
    else if (nv <= 8) RC_GO(8);
    else RC_GO(16);
#undef RC_GO
    return SFM_OK;
}

// ------------------------------------------------------------------ exact tile scan

__global__ void k_flag_all(const __grid_constant__ MatchPlan P) {
    const int p = P.p0 + blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n1 = P.set_cnt[P.pairs[2 * p]];
    if (i < n1) P.flag_rows[(size_t)p * P.nmax + i] = i;
    if (i == 0) P.flag_cnt[p] = n1;
}

// Exclusive prefix of the pairs' work-item counts (one CTA of 256 threads, chunks of 256 pairs).
__global__ void __launch_bounds__(256) k_work_scan(const __grid_constant__ MatchPlan P) {
    __shared__ int s_w[8];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t* wo = P.work_off + P.woff;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < P.pn; base += 256) {
        const int q = base + threadIdx.x;
        const int mine = (q < P.pn) ? ((P.flag_cnt[P.p0 + q] + MX_ROWS - 1) / MX_ROWS) * P.n_xchunks : 0;
        int incl = mine;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        int before = s_carry;
        for (int w = 0; w < warp; ++w) before += s_w[w];
        if (q < P.pn) wo[q] = before + incl - mine;
        __syncthreads();
        if (threadIdx.x == 255) s_carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) wo[P.pn] = s_carry;
}

// Work item = (pair, 8 flagged rows, 1024-column chunk).  Each thread walks its
// columns with all 8 x 8 numpy accumulators in registers; the query rows are
// broadcast from shared memory.
__global__ void __launch_bounds__(256, 1) k_match_exact(const __grid_constant__ MatchPlan P) {
    __shared__ __align__(16) float s_a[MX_ROWS][SFM_DESC_DIM];
    __shared__ Top2 s_red[8][MX_ROWS];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int32_t* wo = P.work_off + P.woff;
    const int total = wo[P.pn];
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        int lo = 0, hi = P.pn - 1;                             // last q with wo[q] <= w
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (wo[mid] <= w) lo = mid; else hi = mid - 1;
        }
        const int p = P.p0 + lo;
        const int local = w - wo[lo];
        const int rg = local / P.n_xchunks, ch = local - rg * P.n_xchunks;
        const int qa = P.pairs[2 * p], qb = P.pairs[2 * p + 1];
        const int n2 = P.set_cnt[qb];
        const int nflag = P.flag_cnt[p];
        const int col0 = ch * MX_COLS;
        const int col1 = min(col0 + MX_COLS, n2);
        __syncthreads();                                       // previous item's smem is free
        for (int q = t; q < MX_ROWS * SFM_DESC_DIM; q += 256) {
            const int r = q >> 7, c = q & 127;
            const int slot = rg * MX_ROWS + r;
            float v = 0.f;
            if (slot < nflag) {
                const int row = P.flag_rows[(size_t)p * P.nmax + slot];
                v = P.set_ptr[qa][(size_t)row * SFM_DESC_DIM + c];
            }
            s_a[r][c] = v;
        }
        __syncthreads();
        Top2 st[MX_ROWS];
#pragma unroll
        for (int r = 0; r < MX_ROWS; ++r) st[r] = {INFINITY, -1, INFINITY};
        const float* B = P.set_ptr[qb];
        for (int jc = col0 + t; jc < col1; jc += 256) {
            const float4* bp = reinterpret_cast<const float4*>(B + (size_t)jc * SFM_DESC_DIM);
            float acc[MX_ROWS][8];
            {
                const float4 b0 = bp[0], b1 = bp[1];
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int r = 0; r < MX_ROWS; ++r)
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float tt = __fsub_rn(s_a[r][q], bb[q]);
                        acc[r][q] = __fmul_rn(tt, tt);
                    }
            }
            float4 nb0 = bp[2], nb1 = bp[3];                       // software-pipelined: next 32 bytes in flight
#pragma unroll 1
            for (int m = 1; m < 16; ++m) {
                const float4 b0 = nb0, b1 = nb1;
                if (m < 15) { nb0 = bp[2 * m + 2]; nb1 = bp[2 * m + 3]; }
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int r = 0; r < MX_ROWS; ++r) {
                    const float4 a0 = *reinterpret_cast<const float4*>(&s_a[r][8 * m]);
                    const float4 a1 = *reinterpret_cast<const float4*>(&s_a[r][8 * m + 4]);
                    const float aa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float tt = __fsub_rn(aa[q], bb[q]);
                        acc[r][q] = __fadd_rn(acc[r][q], __fmul_rn(tt, tt));
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < MX_ROWS; ++r) {
                float lo4 = __fadd_rn(__fadd_rn(acc[r][0], acc[r][1]), __fadd_rn(acc[r][2], acc[r][3]));
                float hi4 = __fadd_rn(__fadd_rn(acc[r][4], acc[r][5]), __fadd_rn(acc[r][6], acc[r][7]));
                top2_push(st[r], __fadd_rn(lo4, hi4), jc);
            }
        }
#pragma unroll
        for (int r = 0; r < MX_ROWS; ++r) {
            Top2 v = st[r];
            for (int o = 16; o > 0; o >>= 1) {
                Top2 u;
                u.d0 = __shfl_xor_sync(0xffffffffu, v.d0, o);
                u.i0 = __shfl_xor_sync(0xffffffffu, v.i0, o);
                u.d1 = __shfl_xor_sync(0xffffffffu, v.d1, o);
                v = top2_merge(v, u);
            }
            if (lane == 0) s_red[warp][r] = v;
        }
        __syncthreads();
        if (t < MX_ROWS) {
            Top2 v = s_red[0][t];
            for (int q = 1; q < 8; ++q) v = top2_merge(v, s_red[q][t]);
            const int slot = rg * MX_ROWS + t;
            if (slot < nflag)
                P.part[((size_t)p * P.nmax + slot) * P.n_xchunks + ch] =
                    make_float4(v.d0, __int_as_float(v.i0), v.d1, 0.f);
        }
    }
}

// Filtered rescan of the rows the certificate rejected (tensor-core mode).  The re-check leaves an
// exact upper bound d1 of each such row's second-nearest squared distance, so only columns that can
// be at most that far need the reference's arithmetic.  The scan itself is a plain float32 FMA dot
// product: approx = |a|^2 + |b|^2 - 2 a.b differs from the exact value by at most E' (dot-product
// rounding by Cauchy-Schwarz, the norms, the float32 rounding of the exact sum), and a column is
// evaluated exactly iff approx <= d1 + E'.
// Work item = (pair, 8 flagged rows, 1024-column chunk), as in k_match_exact.  A warp takes one column
// at a time as ONE coalesced 512-byte read (lane k holds elements 4k..4k+3 -- a thread-per-column
// walk costs 32 L1TEX wavefronts per load instruction and was 8x slower), multiplies it with the
// eight query rows it keeps in registers, and folds the 8 x 32 partial sums with a 9-shuffle
// transpose-reduction that leaves row r's dot product in the lanes whose bits 4..2 spell r.  Each
// lane keeps the rows in the order i -> row i ^ (its bits 4..2), which makes "the half I send" the
// same registers in every lane (no selects).
constexpr int RS_UNROLL = 6;

// One work item with R (2, 4 or 8) query-row slots: most items carry one or two flagged rows, and the
// dot products and the reduction scale with R.
template <int R>
__device__ __forceinline__ void rescan_item(const MatchPlan& P, int p, int rg, int ch, int qa, int qb, int n2, int nflag,
                                            Top2* s_top /* [MX_ROWS] of this warp */) {
    constexpr int LOG = (R == 8) ? 3 : (R == 4 ? 2 : 1);
    constexpr unsigned HITMASK = (R == 8) ? 0x11111111u : (R == 4 ? 0x01010101u : 0x00010001u);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int myrow = lane >> (5 - LOG);
    const int col0 = ch * MX_COLS;
    const int col1 = min(col0 + MX_COLS, n2);
    // the query rows (lane k: elements 4k..4k+3 of each; a[i] is row i ^ myrow), this lane's row norm and threshold
    float4 a[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int slot = rg * MX_ROWS + (r ^ myrow);
        a[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (slot < nflag) {
            const int row = P.flag_rows[(size_t)p * P.nmax + slot];
            a[r] = reinterpret_cast<const float4*>(P.set_ptr[qa] + (size_t)row * SFM_DESC_DIM)[lane];
        }
    }
    float my_na = 0.f, my_T = -INFINITY;                   // empty slot: nothing passes the filter
    {
        const int slot = rg * MX_ROWS + myrow;
        if (slot < nflag) {
            const int row = P.flag_rows[(size_t)p * P.nmax + slot];
            my_na = P.nb[(size_t)qa * P.nmax_pad + row];
            const double mnb = (double)P.setmax[4 * qb + 2];
            const double dn = (double)my_na, ab = sqrt(dn * mnb);
            const double e = 1.1 * (ab * (1.0 / 65536.0) + (dn + mnb) * (1.0 / 262144.0) +
                                    (dn + mnb + 2.0 * ab) * (1.0 / 524288.0)) + 1e-12;
            // rounded up: the comparison below is in float32
            my_T = __double2float_ru((double)P.res_d1[(size_t)p * P.nmax + row] + e);
        }
    }
    const float* B = P.set_ptr[qb];
    const float* nbq = P.nb + (size_t)qb * P.nmax_pad;

    // rotating prefetch: RS_UNROLL columns in flight per warp, each slot refilled as soon as it is read
    float4 buf[RS_UNROLL];
    float buf_nb[RS_UNROLL];
#pragma unroll
    for (int u = 0; u < RS_UNROLL; ++u) {
        const int jc = min(col0 + warp + 8 * u, n2 - 1);
        buf[u] = reinterpret_cast<const float4*>(B + (size_t)jc * SFM_DESC_DIM)[lane];
        buf_nb[u] = nbq[jc];
    }
    for (int jb = col0 + warp; jb < col1; jb += 8 * RS_UNROLL) {
#pragma unroll
        for (int u = 0; u < RS_UNROLL; ++u) {
            const int jc = jb + 8 * u;
            if (jc >= col1) break;                         // warp-uniform
            const float4 b = buf[u];
            const float b_nb = buf_nb[u];
            {
                const int jn = jc + 8 * RS_UNROLL;
                if (jn < col1) {
                    buf[u] = reinterpret_cast<const float4*>(B + (size_t)jn * SFM_DESC_DIM)[lane];
                    buf_nb[u] = nbq[jn];
                }
            }
            float d[R];
#pragma unroll
            for (int r = 0; r < R; ++r)
                d[r] = fmaf(a[r].w, b.w, fmaf(a[r].z, b.z, fmaf(a[r].y, b.y, a[r].x * b.x)));
            // transpose-reduce: R rows x 32 lanes -> row (top LOG bits of the lane), summed over all lanes
#pragma unroll
            for (int half = R / 2, o = 16; half >= 1; half >>= 1, o >>= 1) {
#pragma unroll
                for (int r = 0; r < half; ++r) d[r] += __shfl_xor_sync(0xffffffffu, d[r + half], o);
            }
            float e1 = d[0];
#pragma unroll
            for (int o = 16 >> LOG; o > 0; o >>= 1) e1 += __shfl_xor_sync(0xffffffffu, e1, o);
            const float approx = fmaf(-2.0f, e1, my_na + b_nb);
            unsigned hits = __ballot_sync(0xffffffffu, approx <= my_T) & HITMASK;
            while (hits) {                                 // rare: the reference's arithmetic for (row r, column jc)
                const int src = __ffs(hits) - 1;
                hits &= hits - 1;
                const int r = src >> (5 - LOG);
                float4 av = a[0];                          // a[r ^ myrow] is row r
#pragma unroll
                for (int rr = 1; rr < R; ++rr) if (rr == (r ^ myrow)) av = a[rr];
                // element 4k+q belongs to numpy's accumulator 4(k&1)+q at step k>>1: the running sums
                // walk down the lanes two at a time
                float tt;
                float4 run;
                tt = __fsub_rn(av.x, b.x); run.x = __fmul_rn(tt, tt);
                tt = __fsub_rn(av.y, b.y); run.y = __fmul_rn(tt, tt);
                tt = __fsub_rn(av.z, b.z); run.z = __fmul_rn(tt, tt);
                tt = __fsub_rn(av.w, b.w); run.w = __fmul_rn(tt, tt);
                const float4 sq = run;
#pragma unroll 1
                for (int m = 1; m < 16; ++m) {
                    const float px = __shfl_up_sync(0xffffffffu, run.x, 2);
                    const float py = __shfl_up_sync(0xffffffffu, run.y, 2);
                    const float pz = __shfl_up_sync(0xffffffffu, run.z, 2);
                    const float pw = __shfl_up_sync(0xffffffffu, run.w, 2);
                    if ((lane >> 1) == m) {
                        run.x = __fadd_rn(px, sq.x); run.y = __fadd_rn(py, sq.y);
                        run.z = __fadd_rn(pz, sq.z); run.w = __fadd_rn(pw, sq.w);
                    }
                }
                const float s4 = __fadd_rn(__fadd_rn(run.x, run.y), __fadd_rn(run.z, run.w));  // lane 30: (r0+r1)+(r2+r3), lane 31: (r4+r5)+(r6+r7)
                const float d2 = __fadd_rn(__shfl_sync(0xffffffffu, s4, 30), __shfl_sync(0xffffffffu, s4, 31));
                if (lane == 0) top2_push(s_top[r], d2, jc);
            }
        }
    }
}

__global__ void __launch_bounds__(256, 2) k_match_rescan(const __grid_constant__ MatchPlan P) {
    __shared__ Top2 s_red[8][MX_ROWS];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int32_t* wo = P.work_off + P.woff;
    const int total = wo[P.pn];
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        int lo = 0, hi = P.pn - 1;                             // last q with wo[q] <= w
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (wo[mid] <= w) lo = mid; else hi = mid - 1;
        }
        const int p = P.p0 + lo;
        const int local = w - wo[lo];
        const int rg = local / P.n_xchunks, ch = local - rg * P.n_xchunks;
        const int qa = P.pairs[2 * p], qb = P.pairs[2 * p + 1];
        const int n2 = P.set_cnt[qb];
        const int nflag = P.flag_cnt[p];
        __syncthreads();                                       // previous item's s_red has been read
        if (lane < MX_ROWS) s_red[warp][lane] = {INFINITY, -1, INFINITY};   // this warp's running top-2 per row (lane 0 updates)
        __syncwarp();
        const int live = min(MX_ROWS, nflag - rg * MX_ROWS);
        if (live <= 2) rescan_item<2>(P, p, rg, ch, qa, qb, n2, nflag, s_red[warp]);
        else if (live <= 4) rescan_item<4>(P, p, rg, ch, qa, qb, n2, nflag, s_red[warp]);
        else rescan_item<8>(P, p, rg, ch, qa, qb, n2, nflag, s_red[warp]);
        __syncthreads();
        if (t < MX_ROWS) {
            Top2 v = s_red[0][t];
            for (int q = 1; q < 8; ++q) v = top2_merge(v, s_red[q][t]);
            const int slot = rg * MX_ROWS + t;
            if (slot < nflag)
                P.part[((size_t)p * P.nmax + slot) * P.n_xchunks + ch] =
                    make_float4(v.d0, __int_as_float(v.i0), v.d1, 0.f);
        }
    }
}

__global__ void k_match_merge(const __grid_constant__ MatchPlan P) {
    const int p = P.p0 + blockIdx.y;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= P.flag_cnt[p]) return;
    const int n2 = P.set_cnt[P.pairs[2 * p + 1]];
    const int nch = (n2 + MX_COLS - 1) / MX_COLS;
    Top2 v = {INFINITY, -1, INFINITY};
    for (int c = 0; c < nch; ++c) {
        const float4 q = P.part[((size_t)p * P.nmax + slot) * P.n_xchunks + c];
        Top2 u = {q.x, __float_as_int(q.y), q.z};
        v = top2_merge(v, u);
    }
    const int row = P.flag_rows[(size_t)p * P.nmax + slot];
    const size_t o = (size_t)p * P.nmax + row;
    P.res_idx[o] = v.i0; P.res_d0[o] = v.d0; P.res_d1[o] = v.d1;
}

// ------------------------------------------------------------------ ratio test, ordering

__global__ void k_match_emit(const __grid_constant__ MatchPlan P) {
    const int p = P.p0 + blockIdx.y;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int n1 = P.set_cnt[P.pairs[2 * p]], n2 = P.set_cnt[P.pairs[2 * p + 1]];
    if (row >= n1 || n2 < 2) return;
    const size_t o = (size_t)p * P.nmax + row;
    const float d0 = __fsqrt_rn(P.res_d0[o]), d1 = __fsqrt_rn(P.res_d1[o]);   // :34
    if (d1 > 0.0f) {                                                         // :46
        const float nndr = __fdiv_rn(d0, d1);                                // :47
        if (nndr <= P.thr) {                                                 // :49 (float32 compare)
            const int pos = atomicAdd(&P.mcount[p], 1);
            P.mkeys[(size_t)p * P.nmax + pos] =
                ((unsigned long long)__float_as_uint(nndr) << 32) | (unsigned long long)(uint32_t)row;
            P.midx[(size_t)p * P.nmax + pos] = P.res_idx[o];
        }
    }
}

// Order by (confidence, query row): NNRatioFeatureMatcher.py:56-58 with the tie order made
// canonical.  One CTA per pair sorts the pair's 64-bit keys (confidence bits << 32 | row) in shared
// memory with a bitonic network; the matched index is looked up again through the row.  Pairs with
// more matches than fit (SORT_SMEM_MAX) are left to the rank sort below.
constexpr int SORT_SMEM_MAX = 8192;

__global__ void __launch_bounds__(1024) k_match_sort(const __grid_constant__ MatchPlan P, int n2pow_max,
                                                     int32_t* __restrict__ match_out, float* __restrict__ conf_out,
                                                     int32_t* __restrict__ count_out, int32_t* __restrict__ stats_out) {
    extern __shared__ __align__(16) unsigned char sort_smem[];
    unsigned long long* s_k = reinterpret_cast<unsigned long long*>(sort_smem);
    const int p = P.p0 + blockIdx.x;
    const int n = P.mcount[p];
    if (threadIdx.x == 0) {
        count_out[p] = n < P.cap ? n : P.cap;
        if (stats_out) { stats_out[2 * p] = P.flag_cnt[p]; stats_out[2 * p + 1] = P.stats[2 * p + 1]; }
    }
    if (n == 0 || n > n2pow_max) return;
    int N2 = 32;
    while (N2 < n) N2 <<= 1;
    const unsigned long long* keys = P.mkeys + (size_t)p * P.nmax;
    for (int i = threadIdx.x; i < N2; i += blockDim.x) s_k[i] = (i < n) ? keys[i] : ~0ull;
    // Compare-exchange i of a stage with distance jj <= 32 touches only elements of the 64-aligned block
    // i / 32, i.e. of the warp's own blocks: those stages need a warp barrier, not a CTA barrier (51 of
    // the 66 stages at 2048 keys).
    int prev_jj = 64;
    for (int k = 2; k <= N2; k <<= 1) {
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            if (jj > 32 || prev_jj > 32) __syncthreads(); else __syncwarp();
            prev_jj = jj;
            for (int i = threadIdx.x; i < (N2 >> 1); i += blockDim.x) {
                const int lo = ((i & ~(jj - 1)) << 1) | (i & (jj - 1));
                const int hi = lo | jj;
                const unsigned long long x = s_k[lo], y = s_k[hi];
                const bool up = (lo & k) == 0;
                if ((x > y) == up) { s_k[lo] = y; s_k[hi] = x; }
            }
        }
    }
    __syncthreads();
    const int lim = n < P.cap ? n : P.cap;
    for (int r = threadIdx.x; r < lim; r += blockDim.x) {
        const unsigned long long key = s_k[r];
        const uint32_t row = (uint32_t)key;
        const size_t o = (size_t)p * P.cap + r;
        match_out[2 * o] = (int32_t)row;
        match_out[2 * o + 1] = P.res_idx[(size_t)p * P.nmax + row];
        conf_out[o] = __uint_as_float((uint32_t)(key >> 32));
    }
}

// Rank sort for the pairs the shared-memory sort skipped (more than SORT_SMEM_MAX matches).
__global__ void __launch_bounds__(256) k_match_sort_big(const __grid_constant__ MatchPlan P, int32_t* __restrict__ match_out,
                                                        float* __restrict__ conf_out) {
    __shared__ unsigned long long s_k[256];
    const int p = P.p0 + blockIdx.y;
    const int n = P.mcount[p];
    if (n <= SORT_SMEM_MAX || (int)blockIdx.x * 256 >= n) return;
    const unsigned long long* keys = P.mkeys + (size_t)p * P.nmax;
    const int e = blockIdx.x * 256 + threadIdx.x;
    const unsigned long long mine = (e < n) ? keys[e] : 0ull;
    int rank = 0;
    for (int base = 0; base < n; base += 256) {
        const int jx = base + threadIdx.x;
        s_k[threadIdx.x] = (jx < n) ? keys[jx] : ~0ull;
        __syncthreads();
        if (e < n) {
            const int m = min(256, n - base);
            for (int q = 0; q < m; ++q) rank += (s_k[q] < mine) ? 1 : 0;
        }
        __syncthreads();
    }
    if (e >= n || rank >= P.cap) return;
    const size_t o = (size_t)p * P.cap + rank;
    match_out[2 * o] = (int32_t)(uint32_t)mine;
    match_out[2 * o + 1] = P.midx[(size_t)p * P.nmax + e];
    conf_out[o] = __uint_as_float((uint32_t)(mine >> 32));
}

// ------------------------------------------------------------------ host side

static int choose_splits(int n_pairs, int nmax_pad) {
    const int n_tiles = nmax_pad / MT_COLS;
    const int rowblocks = nmax_pad / MT_ROWS;
    const int max_tiles = mt_max_tiles(mt_group_for(nmax_pad));
    const int min_s = (n_tiles + max_tiles - 1) / max_tiles;
    const long long base = (long long)n_pairs * rowblocks;
    // B200: 148 SMs.  A constant on purpose -- the split count sizes the workspace, and
    // sfm_match_workspace_bytes has no context to ask (sfm_ctx_create accepts sm_100 devices only)
    const int sms = 148;
    if (base >= 4LL * sms) return min_s;
    int best = min_s;
    double best_eff = -1.0;
    for (int s = min_s; s <= MT_MAX_SPLITS && s <= n_tiles; ++s) {
        const long long units = base * s;
        const long long waves = (units + sms - 1) / sms;
        double eff = (double)units / (double)(waves * sms);
        // a split below 8 tiles pays too much prologue per unit
        if ((n_tiles + s - 1) / s < 4 && s > min_s) break;
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    }
    return best;
}

struct MatchWs {
    size_t set_ptr, set_cnt, setmax, h16, nb, hatn, resn, pair_begin;
    size_t pairs, zero_begin, flag_cnt, mcount, stats, work_off, zero_end;
    size_t cands, res_idx, res_d0, res_d1, flag_rows, part, mkeys, midx, total;
};

// Set-side arrays first: their offsets depend on (n_sets, nmax) only, so a workspace prepared once
// (k_match_prep) serves any number of later calls with SFM_MATCH_PREPARED and other pair lists.
static void match_layout(int n_sets, int nmax, int n_pairs, MatchPlan& P, MatchWs& ws) {
    P.n_sets = n_sets; P.nmax = nmax; P.n_pairs = n_pairs;
    P.nmax_pad = (int)align_up((size_t)std::max(nmax, 1), MT_ROWS);
    P.n_tiles = P.nmax_pad / MT_COLS;
    P.group = mt_group_for(P.nmax_pad);
    P.n_splits = choose_splits(n_pairs, P.nmax_pad);
    P.tiles_per_split = (P.n_tiles + P.n_splits - 1) / P.n_splits;
    P.n_lists = 2 * P.n_splits;
    P.n_xchunks = (std::max(nmax, 1) + MX_COLS - 1) / MX_COLS;
    const size_t rows = (size_t)n_sets * P.nmax_pad;
    const size_t pr = (size_t)n_pairs * std::max(nmax, 1);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 256); return at; };
    ws.set_ptr = take(sizeof(void*) * (n_sets + 1));          // + the empty set that out-of-range pair ids are redirected to
    ws.set_cnt = take(sizeof(int32_t) * (n_sets + 1));
    ws.setmax = take(sizeof(float) * 4 * n_sets);
    ws.h16 = take(sizeof(__half) * rows * SFM_DESC_DIM);
    ws.nb = take(sizeof(float) * rows);
    ws.hatn = take(sizeof(float) * rows);
    ws.resn = take(sizeof(float) * rows);
    ws.pair_begin = o;
    ws.pairs = take(sizeof(int32_t) * 2 * n_pairs);
    ws.zero_begin = o;
    ws.flag_cnt = take(sizeof(int32_t) * n_pairs);
    ws.mcount = take(sizeof(int32_t) * n_pairs);
    ws.stats = take(sizeof(int32_t) * 2 * n_pairs);
    ws.work_off = take(sizeof(int32_t) * (n_pairs + 1));
    ws.zero_end = o;
    ws.cands = take(sizeof(uint32_t) * (size_t)n_pairs * P.nmax_pad * P.n_lists * MT_TOPK);
    ws.res_idx = take(sizeof(int32_t) * pr);
    ws.res_d0 = take(sizeof(float) * pr);
    ws.res_d1 = take(sizeof(float) * pr);
    ws.flag_rows = take(sizeof(int32_t) * pr);
    ws.part = take(sizeof(float4) * pr * P.n_xchunks);
    ws.mkeys = take(sizeof(unsigned long long) * pr);
    ws.midx = take(sizeof(int32_t) * pr);
    ws.total = o;
}

static void match_bind(MatchPlan& P, const MatchWs& ws, void* base) {
    char* c = (char*)base;
    P.set_ptr = (const float**)(c + ws.set_ptr);
    P.set_cnt = (int32_t*)(c + ws.set_cnt);
    P.pairs = (int32_t*)(c + ws.pairs);
    P.setmax = (float*)(c + ws.setmax);
    P.flag_cnt = (int32_t*)(c + ws.flag_cnt);
    P.mcount = (int32_t*)(c + ws.mcount);
    P.stats = (int32_t*)(c + ws.stats);
    P.work_off = (int32_t*)(c + ws.work_off);
    P.h16 = (__half*)(c + ws.h16);
    P.nb = (float*)(c + ws.nb);
    P.hatn = (float*)(c + ws.hatn);
    P.resn = (float*)(c + ws.resn);
    P.cands = (uint32_t*)(c + ws.cands);
    P.res_idx = (int32_t*)(c + ws.res_idx);
    P.res_d0 = (float*)(c + ws.res_d0);
    P.res_d1 = (float*)(c + ws.res_d1);
    P.flag_rows = (int32_t*)(c + ws.flag_rows);
    P.part = (float4*)(c + ws.part);
    P.mkeys = (unsigned long long*)(c + ws.mkeys);
    P.midx = (int32_t*)(c + ws.midx);
}

// Everything after the (global) prep for the pair chunk [p0, p0 + pn) on stream `s`.
static int run_match_chunk(SfmCtx* ctx, cudaStream_t s, MatchPlan P, int p0, int pn,
                           int32_t* match_out, float* conf_out, int32_t* count_out, int32_t* stats_out) {
    P.p0 = p0; P.pn = pn; P.woff = 0;             // chunks are stream-ordered: the work-prefix table is reused
    const dim3 rowgrid(ceil_div(P.nmax, 256), pn);
    if (P.mode == SFM_MATCH_AUTO) {
        int rc = launch_match_tc(ctx, s, P);
        if (rc) return rc;
        rc = launch_match_recheck(ctx, s, P);
        if (rc) return rc;
    } else {
        SFM_LAUNCH(ctx, s, "k_flag_all", k_flag_all<<<rowgrid, 256, 0, s>>>(P));
    }
    SFM_LAUNCH(ctx, s, "k_work_scan", k_work_scan<<<1, 256, 0, s>>>(P));
    // persistent grids: exactly the CTAs that are resident at once (a second wave would start on items
    // the first one has already walked past)
    int per_sm = 1;
    if (P.mode == SFM_MATCH_AUTO) {
        SFM_CUDA_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_match_rescan, 256, 0));
        SFM_LAUNCH(ctx, s, "k_match_rescan", k_match_rescan<<<std::max(per_sm, 1) * ctx->sm_count, 256, 0, s>>>(P));
    } else {
        SFM_CUDA_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_match_exact, 256, 0));
        SFM_LAUNCH(ctx, s, "k_match_exact", k_match_exact<<<std::max(per_sm, 1) * ctx->sm_count, 256, 0, s>>>(P));
    }
    SFM_LAUNCH(ctx, s, "k_match_merge", k_match_merge<<<rowgrid, 256, 0, s>>>(P));
    SFM_LAUNCH(ctx, s, "k_match_emit", k_match_emit<<<rowgrid, 256, 0, s>>>(P));
    {
        int n2pow = 32;
        while (n2pow < P.nmax && n2pow < SORT_SMEM_MAX) n2pow <<= 1;
        const int smem = n2pow * (int)sizeof(unsigned long long);
        SFM_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_match_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        SFM_LAUNCH(ctx, s, "k_match_sort", k_match_sort<<<pn, n2pow >= 2048 ? 1024 : 256, smem, s>>>(
                                               P, n2pow, match_out, conf_out, count_out, stats_out));
        if (P.nmax > SORT_SMEM_MAX)
            SFM_LAUNCH(ctx, s, "k_match_sort_big", k_match_sort_big<<<rowgrid, 256, 0, s>>>(P, match_out, conf_out));
    }
    return SFM_OK;
}

static int run_match(SfmCtx* ctx, cudaStream_t st, MatchPlan& P, const MatchWs& ws, void* workspace,
                     int32_t* match_out, float* conf_out, int32_t* count_out, int32_t* stats_out, bool prepared) {
    SFM_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    SFM_CUDA_CHECK(ctx, cudaMemsetAsync((char*)workspace + ws.zero_begin, 0, ws.zero_end - ws.zero_begin, st));
    if (!prepared) SFM_CUDA_CHECK(ctx, cudaMemsetAsync((char*)workspace + ws.setmax, 0, sizeof(float) * 4 * P.n_sets, st));
    const int nt = std::max(P.n_sets, P.n_pairs);
    P.p0 = 0; P.pn = P.n_pairs; P.woff = 0;
    // (with prepared sets k_match_setup rewrites the same set table and copies this call's pair list)
    SFM_LAUNCH(ctx, st, "k_match_setup", k_match_setup<<<ceil_div(nt, 256), 256, 0, st>>>(P));
    if (P.mode == SFM_MATCH_AUTO && !prepared)
        SFM_LAUNCH(ctx, st, "k_match_prep", k_match_prep<<<dim3(P.nmax_pad / 8, P.n_sets), 256, 0, st>>>(P));
    // Pair chunks of at most 65 535 (gridDim.y of the per-row kernels), one after the other on the caller's stream;
    // a batch below that is one chunk.  (Measured on B200: overlapping the re-check of one chunk with the
    // tensor-core pass of the next on two streams buys nothing -- both are bound by the same L2 -> SM bandwidth.)
    constexpr int MAX_CHUNK = 65535;
    for (int p0 = 0; p0 < P.n_pairs; p0 += MAX_CHUNK) {
        const int rc = run_match_chunk(ctx, st, P, p0, std::min(MAX_CHUNK, P.n_pairs - p0), match_out, conf_out, count_out, stats_out);
        if (rc) return rc;
    }
    return SFM_OK;
}

extern "C" {

size_t sfm_match_workspace_bytes(int n_sets, int nmax, int n_pairs) {
    if (n_sets < 1 || nmax < 1 || n_pairs < 1) return 0;
    MatchPlan P;
    MatchWs ws;
    memset(&P, 0, sizeof(P));
    match_layout(n_sets, nmax, n_pairs, P, ws);
    return ws.total;
}

size_t sfm_match_prepared_bytes(int n_sets, int nmax) {
    if (n_sets < 1 || nmax < 1) return 0;
    MatchPlan P;
    MatchWs ws;
    memset(&P, 0, sizeof(P));
    match_layout(n_sets, nmax, 1, P, ws);
    return ws.pair_begin;
}

int sfm_match_ratio(SfmCtx* ctx, void* stream, const float* f1_dev, int n1, const float* f2_dev,
                    int n2, int dim, float ratio_threshold, int mode, void* workspace_dev,
                    size_t workspace_bytes, int32_t* match_out, float* conf_out,
                    int32_t* count_out, int cap) {
    if (!ctx) return SFM_ERR_BAD_ARG;
    if (dim != SFM_DESC_DIM) return sfm_set_error(ctx, SFM_ERR_UNSUPPORTED, "descriptor dim %d != 128", dim);
    if (!f1_dev || !f2_dev || !workspace_dev || !match_out || !conf_out || !count_out)
        return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "NULL pointer");
    if (n1 < 1 || n2 < 2)
        return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "need n1 >= 1 and n2 >= 2 (got %d, %d): the reference indexes the second neighbour", n1, n2);
    if (((uintptr_t)f1_dev | (uintptr_t)f2_dev) & 15)
        return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "descriptor pointers must be 16-byte aligned");
    if (mode != SFM_MATCH_AUTO && mode != SFM_MATCH_EXACT) return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "bad mode");
    MatchPlan P;
    MatchWs ws;
    memset(&P, 0, sizeof(P));
    match_layout(2, std::max(n1, n2), 1, P, ws);
    if (workspace_bytes < ws.total)
        return sfm_set_error(ctx, SFM_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, ws.total);
    if (cap < 1) return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "cap < 1");
    match_bind(P, ws, workspace_dev);
    P.f1 = f1_dev; P.f2 = f2_dev; P.n1 = n1; P.n2 = n2;
    P.thr = ratio_threshold; P.mode = mode; P.cap = cap; P.no_prune = 0;
    return run_match(ctx, (cudaStream_t)stream, P, ws, workspace_dev, match_out, conf_out, count_out, nullptr, false);
}

int sfm_match_ratio_batch(SfmCtx* ctx, void* stream, const float* desc_dev, const int32_t* counts_dev,
                          int n_sets, int nmax, const int32_t* pairs_dev, int n_pairs,
                          float ratio_threshold, int mode, void* workspace_dev, size_t workspace_bytes,
                          int32_t* match_out, float* conf_out, int32_t* count_out, int32_t* stats_out,
                          int cap) {
    if (!ctx) return SFM_ERR_BAD_ARG;
    if (!desc_dev || !counts_dev || !pairs_dev || !workspace_dev || !match_out || !conf_out || !count_out)
        return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "NULL pointer");
    if (n_sets < 1 || nmax < 2 || n_pairs < 1) return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "bad sizes");
    if ((uintptr_t)desc_dev & 15) return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "descriptor pointer must be 16-byte aligned");
    const bool prepared = (mode & SFM_MATCH_PREPARED) != 0;
    const int no_prune = (mode & SFM_MATCH_NO_PRUNE) ? 1 : 0;
    mode &= ~(SFM_MATCH_PREPARED | SFM_MATCH_NO_PRUNE);
    if (mode != SFM_MATCH_AUTO && mode != SFM_MATCH_EXACT) return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "bad mode");
    MatchPlan P;
    MatchWs ws;
    memset(&P, 0, sizeof(P));
    match_layout(n_sets, nmax, n_pairs, P, ws);
    if (workspace_bytes < ws.total)
        return sfm_set_error(ctx, SFM_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, ws.total);
    if (cap < 1) return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "cap < 1");
    match_bind(P, ws, workspace_dev);
    P.desc = desc_dev; P.counts_in = counts_dev; P.pairs_in = pairs_dev;
    P.thr = ratio_threshold; P.mode = mode; P.cap = cap; P.no_prune = no_prune;
    return run_match(ctx, (cudaStream_t)stream, P, ws, workspace_dev, match_out, conf_out, count_out, stats_out, prepared);
}

}  // extern "C"
