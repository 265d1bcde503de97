// assoc.cu -- brute-force nearest-point association loops that follow the two-view stage
// (SURVEY.md section 8f row 3).
//
// Reference (paths relative to the reference root):
//   Runner.py:241-247  for every 2-D point of the previous frame's match list: the nearest point of
//                      the already triangulated set (CameraPose.compute_euclidean_distance,
//                      SFM.py:376-382, + np.argmin), kept when closer than dist_threshold
//   Runner.py:361-385  add_points / is_new_point / find_existing_point: 3-D points are appended to
//                      the global store unless one at distance < 1e-6 is already there
//
// Both are O(N*M) scans of np.linalg.norm rows in the reference.  Here: one warp per query point,
// float64, the distance evaluated exactly as numpy does (squares summed left to right, IEEE sqrt,
// comparisons on the rooted value, first minimum wins), so indices are bit-exact.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace {

// (distance, index) minimum with numpy's first-occurrence tie-break
__device__ __forceinline__ void arg_min_step(double d, int i, double& bd, int& bi) {
    if (d < bd || (d == bd && i < bi)) { bd = d; bi = i; }
}

__device__ __forceinline__ void warp_arg_min(double& bd, int& bi) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double d = __shfl_xor_sync(0xffffffffu, bd, o);
        const int i = __shfl_xor_sync(0xffffffffu, bi, o);
        arg_min_step(d, i, bd, bi);
    }
}

// Runner.py:241-247.  ref [m][2], query [q][2].
__global__ void __launch_bounds__(256) k_assoc_nearest(const double* __restrict__ ref, int m,
                                                       const double* __restrict__ query, int q, double thr,
                                                       int32_t* __restrict__ nearest, double* __restrict__ dist,
                                                       int32_t* __restrict__ flag) {
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= q) return;
    const double2 b = reinterpret_cast<const double2*>(query)[w];
    double bd = INFINITY;
    int bi = 0x7fffffff;
    for (int i = lane; i < m; i += 32) {
        const double2 a = reinterpret_cast<const double2*>(ref)[i];
        const double dx = __dsub_rn(a.x, b.x), dy = __dsub_rn(a.y, b.y);
        arg_min_step(sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))), i, bd, bi);
    }
    warp_arg_min(bd, bi);
    if (lane == 0) {
        nearest[w] = bi;
        if (dist) dist[w] = bd;
        flag[w] = (bd < thr) ? 1 : 0;
    }
}

// Stable compaction of the rows whose flag is set (one CTA; q is a few thousand).
__global__ void __launch_bounds__(1024) k_assoc_compact(const int32_t* __restrict__ flag, int q,
                                                        int32_t* __restrict__ kept, int32_t* __restrict__ count) {
    __shared__ int s_scan[32];
    __shared__ int s_off, s_total;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_off = 0;
    __syncthreads();
    for (int base = 0; base < q; base += 1024) {
        const int i = base + t;
        const bool in = (i < q) && flag[i] != 0;
        const unsigned ball = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_scan[warp] = __popc(ball);
        __syncthreads();
        if (warp == 0) {
            const int v = s_scan[lane];
            int incl = v;
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            s_scan[lane] = incl - v;
            if (lane == 31) s_total = incl;
        }
        __syncthreads();
        if (in) kept[s_off + s_scan[warp] + __popc(ball & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (t == 0) s_off += s_total;
        __syncthreads();
    }
    if (t == 0) *count = s_off;
}

struct DedupPlan {
    const double* pts;     // [n][3]
    const double* store;   // [e][3]
    int n, e, pair_cap;
    double thr;
    double* e_dist;        // [n] nearest existing store point closer than thr (INF: none)
    int32_t* e_idx;        // [n]
    int32_t* cnt;          // [n] earlier batch points closer than thr
    int32_t* off;          // [n + 1]
    int32_t* pair_j;       // [pair_cap]
    double* pair_d;        // [pair_cap]
    int32_t* flag;         // [1] overflow
    int32_t* rank;         // [n] rank among the new points (-1: not new)
    int32_t* index_out;    // [n]
    int32_t* is_new_out;   // [n]
    int32_t* n_new_out;    // [1]
};

__device__ __forceinline__ double dist3(const double* a, const double* b) {
    const double dx = __dsub_rn(a[0], b[0]), dy = __dsub_rn(a[1], b[1]), dz = __dsub_rn(a[2], b[2]);
    return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

// Pass A (fill == 0): nearest existing point within thr, number of earlier batch points within thr.
// Pass B (fill == 1): the earlier batch points within thr, ascending j, at off[i].
template <int FILL>
__global__ void __launch_bounds__(256) k_dedup_scan(const __grid_constant__ DedupPlan P) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= P.n) return;
    if (FILL && *P.flag) return;
    const double p[3] = {P.pts[3 * (size_t)i], P.pts[3 * (size_t)i + 1], P.pts[3 * (size_t)i + 2]};
    if (!FILL) {
        double bd = INFINITY;
        int bi = 0x7fffffff;
        for (int k = lane; k < P.e; k += 32) {
            const double d = dist3(P.store + 3 * (size_t)k, p);
            if (d < P.thr) arg_min_step(d, k, bd, bi);
        }
        warp_arg_min(bd, bi);
        if (lane == 0) { P.e_dist[i] = bd; P.e_idx[i] = bi; }
    }
    int total = 0;
    const int base_off = FILL ? P.off[i] : 0;
    for (int j0 = 0; j0 < i; j0 += 32) {
        const int j = j0 + lane;
        double d = INFINITY;
        if (j < i) d = dist3(P.pts + 3 * (size_t)j, p);
        const bool near = d < P.thr;
        const unsigned ball = __ballot_sync(0xffffffffu, near);
        if (FILL && near) {
            const int pos = base_off + total + __popc(ball & ((1u << lane) - 1u));
            P.pair_j[pos] = j;
            P.pair_d[pos] = d;
        }
        total += __popc(ball);
    }
    if (!FILL && lane == 0) P.cnt[i] = total;
}

// Exclusive scan of cnt into off (one CTA), overflow flag.
__global__ void __launch_bounds__(1024) k_dedup_offsets(const __grid_constant__ DedupPlan P) {
    __shared__ long long s_scan[32];
    __shared__ long long s_off, s_total;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_off = 0;
    __syncthreads();
    for (int base = 0; base < P.n; base += 1024) {
        const int i = base + t;
        const long long v = (i < P.n) ? P.cnt[i] : 0;
        long long incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const long long w = s_scan[lane];
            long long wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                const long long u = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += u;
            }
            s_scan[lane] = wi - w;
            if (lane == 31) s_total = wi;
        }
        __syncthreads();
        const long long excl = s_off + s_scan[warp] + incl - v;
        if (i < P.n) P.off[i] = (int32_t)(excl > 0x7fffffffLL ? 0x7fffffffLL : excl);
        __syncthreads();
        if (t == 0) s_off += s_total;
        __syncthreads();
    }
    if (t == 0) {
        P.off[P.n] = (int32_t)(s_off > 0x7fffffffLL ? 0x7fffffffLL : s_off);
        *P.flag = (s_off > (long long)P.pair_cap) ? 1 : 0;
    }
}

// Runner.py:361-371 in order: a point is new iff no stored point (existing, or an earlier point
// of this batch that was itself new) lies closer than thr; otherwise it maps to the first nearest
// stored point.  One warp walks the batch; the near lists are almost always empty.
__global__ void __launch_bounds__(32) k_dedup_resolve(const __grid_constant__ DedupPlan P) {
    const int lane = threadIdx.x;
    if (*P.flag) { if (lane == 0) { *P.n_new_out = -1; P.index_out[0] = P.off[P.n]; } return; }
    int n_new = 0;
    for (int i = 0; i < P.n; ++i) {
        double bd = P.e_dist[i];
        int bi = (bd < P.thr) ? P.e_idx[i] : 0x7fffffff;
        if (!(bd < P.thr)) bd = INFINITY;
        const int o0 = P.off[i], o1 = P.off[i + 1];
        for (int k = o0 + lane; k < o1; k += 32) {
            const int r = P.rank[P.pair_j[k]];
            if (r >= 0) arg_min_step(P.pair_d[k], P.e + r, bd, bi);
        }
        if (o1 > o0) warp_arg_min(bd, bi);
        const bool is_new = !(bd < P.thr);
        if (lane == 0) {
            P.rank[i] = is_new ? n_new : -1;
            P.index_out[i] = is_new ? P.e + n_new : bi;
            P.is_new_out[i] = is_new ? 1 : 0;
        }
        n_new += is_new ? 1 : 0;
        __syncwarp();
    }
    if (lane == 0) *P.n_new_out = n_new;
}

struct DedupWs { size_t e_dist, e_idx, cnt, off, pair_j, pair_d, flag, rank, total; };

DedupWs dedup_layout(int n, int pair_cap) {
    DedupWs w;
    size_t o = 0;
    w.e_dist = o; o = align_up(o + (size_t)n * 8, 256);
    w.pair_d = o; o = align_up(o + (size_t)pair_cap * 8, 256);
    w.e_idx = o;  o = align_up(o + (size_t)n * 4, 256);
    w.cnt = o;    o = align_up(o + (size_t)n * 4, 256);
    w.off = o;    o = align_up(o + (size_t)(n + 1) * 4, 256);
    w.pair_j = o; o = align_up(o + (size_t)pair_cap * 4, 256);
    w.rank = o;   o = align_up(o + (size_t)n * 4, 256);
    w.flag = o;   o = align_up(o + 4, 256);
    w.total = o;
    return w;
}

}  // namespace

extern "C" {

SFM_EXPORT int sfm_associate_nearest(SfmCtx* ctx, void* stream, const double* ref_dev, int m, const double* query_dev,
                                     int q, double dist_threshold, int32_t* nearest_out, double* dist_out,
                                     int32_t* flag_out, int32_t* kept_out, int32_t* count_out) {
    if (!ctx) return SFM_ERR_BAD_ARG;
    if (!ref_dev || !query_dev || !nearest_out || !flag_out || !kept_out || !count_out || m < 1 || q < 1)
        return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "associate_nearest: null pointer or empty point set");
    cudaStream_t st = (cudaStream_t)stream;
    SFM_LAUNCH(ctx, st, "k_assoc_nearest",
               k_assoc_nearest<<<ceil_div(q, 8), 256, 0, st>>>(ref_dev, m, query_dev, q, dist_threshold, nearest_out, dist_out, flag_out));
    SFM_LAUNCH(ctx, st, "k_assoc_compact", k_assoc_compact<<<1, 1024, 0, st>>>(flag_out, q, kept_out, count_out));
    return SFM_OK;
}

SFM_EXPORT size_t sfm_dedup_workspace_bytes(int n, int pair_cap) {
    return (n > 0 && pair_cap >= 0) ? dedup_layout(n, pair_cap).total : 0;
}

SFM_EXPORT int sfm_dedup_points(SfmCtx* ctx, void* stream, const double* pts_dev, int n, const double* store_dev, int e,
                                double threshold, int pair_cap, void* workspace_dev, size_t workspace_bytes,
                                int32_t* index_out, int32_t* is_new_out, int32_t* n_new_out) {
    if (!ctx) return SFM_ERR_BAD_ARG;
    if (!pts_dev || (e > 0 && !store_dev) || !workspace_dev || !index_out || !is_new_out || !n_new_out || n < 1 || e < 0 || pair_cap < 0)
        return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "dedup_points: null pointer or bad size");
    const DedupWs w = dedup_layout(n, pair_cap);
    if (workspace_bytes < w.total)
        return sfm_set_error(ctx, SFM_ERR_WORKSPACE, "dedup workspace: %zu bytes given, %zu needed", workspace_bytes, w.total);
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)workspace_dev;
    DedupPlan P;
    memset(&P, 0, sizeof(P));
    P.pts = pts_dev; P.store = store_dev; P.n = n; P.e = e; P.pair_cap = pair_cap; P.thr = threshold;
    P.e_dist = (double*)(base + w.e_dist); P.pair_d = (double*)(base + w.pair_d);
    P.e_idx = (int32_t*)(base + w.e_idx); P.cnt = (int32_t*)(base + w.cnt); P.off = (int32_t*)(base + w.off);
    P.pair_j = (int32_t*)(base + w.pair_j); P.rank = (int32_t*)(base + w.rank); P.flag = (int32_t*)(base + w.flag);
    P.index_out = index_out; P.is_new_out = is_new_out; P.n_new_out = n_new_out;
    SFM_LAUNCH(ctx, st, "k_dedup_scan", k_dedup_scan<0><<<ceil_div(n, 8), 256, 0, st>>>(P));
    SFM_LAUNCH(ctx, st, "k_dedup_offsets", k_dedup_offsets<<<1, 1024, 0, st>>>(P));
    SFM_LAUNCH(ctx, st, "k_dedup_fill", k_dedup_scan<1><<<ceil_div(n, 8), 256, 0, st>>>(P));
    SFM_LAUNCH(ctx, st, "k_dedup_resolve", k_dedup_resolve<<<1, 32, 0, st>>>(P));
    return SFM_OK;
}

}  // extern "C"
