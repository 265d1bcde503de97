// ransac.cu -- two-view 8-point RANSAC on the matcher's output.
//
// Reference (paths relative to the reference root):
//   SFM.py:126-160  CameraPose.find_inliers          (every consecutive pair, Runner.py:351)
//   SFM.py:38-102   CameraPose.ransac_camera_motion  (the initial pair, Runner.py:203)
//   SFM.py:104-124  _check_valid_pose, :239-253 triangulate_point, :163-178 normalize_points,
//   SFM.py:189-236  _compute_fundamental_matrix
//
// The reference evaluates its hypotheses one after the other (5 967 iterations per pair at
// Runner.py:170), each with two LAPACK SVDs and a pass over all correspondences.  Here every
// hypothesis is independent work:
//   k_ransac_fit     one thread per hypothesis: Hartley normalisation, null vector of the 8x9
//                    system by Householder QR of its transpose, rank-2 projection by a one-sided
//                    Jacobi SVD, un-normalisation; in pose mode also E = K2^T F K1 and its four
//                    (R, T) decompositions.  All float64.
//   k_ransac_valid   pose mode: one warp per (hypothesis, candidate); lanes triangulate
//                    correspondences (DLT, 4x4 one-sided Jacobi SVD) 32 at a time and the warp
//                    stops at the first one behind a camera, as the reference's loop does.
//   k_ransac_score   one warp per hypothesis: epipolar distances of all correspondences,
//                    inlier count.
//   k_ransac_select  one CTA: first hypothesis with the largest count (the reference replaces its
//                    best only on a strictly larger count), then the winner's inlier indices in
//                    ascending order (the reference's boolean-mask gather).
// The 8-subsets come from numpy's legacy global generator (np.random.seed(5) followed by
// np.random.choice(n, 8, replace=False) per iteration): an inherently sequential MT19937 stream
// with rejection sampling, reproduced on the host by sfm_ransac_sample_indices.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <memory>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "ransac_math.cuh"

namespace {

struct RansacPlan {
    const double* p1;         // [n][2]
    const double* p2;         // [n][2]
    const int32_t* samples;   // [iters][8]
    int n, iters, pose;
    double thr;
    double* F;                // [iters][9]
    int32_t* counts;          // [iters]
    uint32_t* valid;          // [iters], bit c = candidate c passes the cheirality test (pose mode); bit 31 = degenerate sample
    double* cand;             // [iters][4][12]: R row-major, T (pose mode)
    double K1[9], K2[9], Rb[9], Tb[3], P1[12];
    int32_t* result;          // [4]: winner (-1: none), its inlier count, its valid bits, number of degenerate samples
    int32_t* inliers;         // [n]
    double* best;             // [9 + 48]: winner's F and candidates (may be null)
};

__global__ void __launch_bounds__(64) k_ransac_fit(const __grid_constant__ RansacPlan P) {
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= P.iters) return;
    double x1[8], y1[8], x2[8], y2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int idx = P.samples[it * 8 + j];
        const double2 a = reinterpret_cast<const double2*>(P.p1)[idx], b = reinterpret_cast<const double2*>(P.p2)[idx];
        x1[j] = a.x; y1[j] = a.y; x2[j] = b.x; y2[j] = b.y;
    }
    double F[9];
    bool degenerate;
    fundamental_8pt(x1, y1, x2, y2, F, &degenerate);
#pragma unroll
    for (int i = 0; i < 9; ++i) P.F[(size_t)it * 9 + i] = F[i];
    P.valid[it] = degenerate ? 0x80000000u : 0u;
    if (P.pose) {
        double cand[48];
        pose_candidates(F, P.K1, P.K2, cand);
#pragma unroll
        for (int i = 0; i < 48; ++i) P.cand[(size_t)it * 48 + i] = cand[i];
    }
}

// SFM.py:104-124 for one (hypothesis, candidate) per warp.
__global__ void __launch_bounds__(128) k_ransac_valid(const __grid_constant__ RansacPlan P) {
    const int it = blockIdx.x, c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* cd = P.cand + (size_t)it * 48 + c * 12;
    double Rc[9], Tc[3], P2[12];
#pragma unroll
    for (int i = 0; i < 9; ++i) Rc[i] = cd[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) Tc[i] = cd[9 + i];
    projection3x4(P.K2, Rc, Tc, P2);
    bool ok = true;
    for (int base = 0; base < P.n; base += 32) {
        const int i = base + lane;
        bool bad = false;
        if (i < P.n)
            bad = !point_in_front(P.P1, P2, P.Rb, P.Tb, Rc, Tc, reinterpret_cast<const double2*>(P.p1)[i],
                                  reinterpret_cast<const double2*>(P.p2)[i]);
        if (__any_sync(0xffffffffu, bad)) { ok = false; break; }
    }
    if (lane == 0 && ok) atomicOr(&P.valid[it], 1u << c);
}

__global__ void __launch_bounds__(256) k_ransac_score(const __grid_constant__ RansacPlan P) {
    const int it = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (it >= P.iters) return;
    if (P.pose && (P.valid[it] & 0xfu) == 0) { if (lane == 0) P.counts[it] = 0; return; }
    double F[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] = P.F[(size_t)it * 9 + i];
    int cnt = 0;
    for (int i = lane; i < P.n; i += 32)
        cnt += is_inlier(F, reinterpret_cast<const double2*>(P.p1)[i], reinterpret_cast<const double2*>(P.p2)[i], P.thr) ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) P.counts[it] = cnt;
}

__global__ void __launch_bounds__(1024) k_ransac_select(const __grid_constant__ RansacPlan P) {
    __shared__ unsigned long long s_best[32];
    __shared__ int s_scan[32];
    __shared__ int s_off, s_total;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // key: count descending, then iteration ascending; a hypothesis needs at least one inlier
    // (and, in pose mode, one valid candidate) to replace the reference's empty initial best
    unsigned long long best = 0ull;
    int n_deg = 0;
    for (int it = t; it < P.iters; it += 1024) {
        const int c = P.counts[it];
        const uint32_t v = P.valid[it];
        n_deg += (int)(v >> 31);
        if (c > 0 && (!P.pose || (v & 0xfu) != 0)) {
            const unsigned long long key = ((unsigned long long)(uint32_t)c << 32) | (uint32_t)(0x7fffffff - it);
            best = key > best ? key : best;
        }
    }
    for (int o = 16; o; o >>= 1) {
        const unsigned long long v = __shfl_xor_sync(0xffffffffu, best, o);
        best = v > best ? v : best;
    }
    n_deg = __reduce_add_sync(0xffffffffu, n_deg);
    if (t == 0) s_total = 0;
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
    if (lane == 0 && n_deg) atomicAdd(&s_total, n_deg);
    if (warp == 0) {
        best = s_best[lane];
        for (int o = 16; o; o >>= 1) {
            const unsigned long long v = __shfl_xor_sync(0xffffffffu, best, o);
            best = v > best ? v : best;
        }
        if (lane == 0) s_best[0] = best;
    }
    __syncthreads();
    best = s_best[0];
    n_deg = s_total;
    if (best == 0ull) {
        if (t < 4) P.result[t] = (t == 0) ? -1 : (t == 3 ? n_deg : 0);
        return;
    }
    const int win = 0x7fffffff - (int)(uint32_t)(best & 0xffffffffull);
    double F[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] = P.F[(size_t)win * 9 + i];
    if (t == 0) {
        P.result[0] = win;
        P.result[1] = (int)(best >> 32);
        P.result[2] = P.pose ? (int)(P.valid[win] & 0xfu) : 0;
        P.result[3] = n_deg;
        s_off = 0;
    }
    if (P.best) {
        if (t < 9) P.best[t] = F[t];
        if (P.pose && t < 48) P.best[9 + t] = P.cand[(size_t)win * 48 + t];
    }
    __syncthreads();
    for (int base = 0; base < P.n; base += 1024) {
        const int i = base + t;
        const bool in = (i < P.n) && is_inlier(F, reinterpret_cast<const double2*>(P.p1)[i],
                                               reinterpret_cast<const double2*>(P.p2)[i], P.thr);
        const unsigned ball = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_scan[warp] = __popc(ball);
        __syncthreads();
        if (warp == 0) {
            int v = s_scan[lane], incl = v;
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            s_scan[lane] = incl - v;
            if (lane == 31) s_total = incl;
        }
        __syncthreads();
        const int off = s_off + s_scan[warp] + __popc(ball & ((1u << lane) - 1u));
        if (in) P.inliers[off] = i;
        __syncthreads();
        if (t == 0) s_off += s_total;
        __syncthreads();
    }
}

// Runner.py:423-434.
__global__ void k_matches_to_coords(const int32_t* __restrict__ m, const int32_t* __restrict__ cnt,
                                    const int32_t* __restrict__ x1, const int32_t* __restrict__ y1,
                                    const int32_t* __restrict__ x2, const int32_t* __restrict__ y2, int num,
                                    double* __restrict__ p1, double* __restrict__ p2, int32_t* __restrict__ n_out) {
    const int n = min(*cnt, num);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *n_out = n;
    if (i >= n) return;
    const int2 ab = reinterpret_cast<const int2*>(m)[i];
    reinterpret_cast<double2*>(p1)[i] = make_double2((double)x1[ab.x], (double)y1[ab.x]);
    reinterpret_cast<double2*>(p2)[i] = make_double2((double)x2[ab.y], (double)y2[ab.y]);
}

struct WsLayout { size_t F, counts, valid, cand, total; };

WsLayout ws_layout(int iters) {
    WsLayout w;
    size_t o = 0;
    w.F = o;      o = align_up(o + (size_t)iters * 9 * sizeof(double), 256);
    w.cand = o;   o = align_up(o + (size_t)iters * 48 * sizeof(double), 256);
    w.counts = o; o = align_up(o + (size_t)iters * sizeof(int32_t), 256);
    w.valid = o;  o = align_up(o + (size_t)iters * sizeof(uint32_t), 256);
    w.total = o;
    return w;
}

int run_ransac(SfmCtx* ctx, void* stream, const double* p1, const double* p2, int n, const int32_t* samples, int iters,
               double thr, int pose, const double* K1, const double* K2, const double* Rb, const double* Tb,
               void* ws, size_t ws_bytes, int32_t* inliers, int32_t* result, double* best_out) {
    if (!ctx) return SFM_ERR_BAD_ARG;
    if (!p1 || !p2 || !samples || !ws || !inliers || !result || n < 8 || iters < 1)
        return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "ransac: null pointer, fewer than 8 correspondences or no iterations");
    if (pose && (!K1 || !K2 || !Rb || !Tb || !best_out))
        return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "ransac_camera_motion: K1, K2, R_base, T_base and pose_out are required");
    const WsLayout w = ws_layout(iters);
    if (ws_bytes < w.total) return sfm_set_error(ctx, SFM_ERR_WORKSPACE, "ransac workspace: %zu bytes given, %zu needed", ws_bytes, w.total);
    cudaStream_t st = (cudaStream_t)stream;
    RansacPlan P;
    memset(&P, 0, sizeof(P));
    P.p1 = p1; P.p2 = p2; P.samples = samples; P.n = n; P.iters = iters; P.pose = pose; P.thr = thr;
    char* base = (char*)ws;
    P.F = (double*)(base + w.F); P.cand = (double*)(base + w.cand);
    P.counts = (int32_t*)(base + w.counts); P.valid = (uint32_t*)(base + w.valid);
    P.result = result; P.inliers = inliers; P.best = best_out;
    if (pose) {
        memcpy(P.K1, K1, sizeof(P.K1)); memcpy(P.K2, K2, sizeof(P.K2));
        memcpy(P.Rb, Rb, sizeof(P.Rb)); memcpy(P.Tb, Tb, sizeof(P.Tb));
        projection3x4(K1, Rb, Tb, P.P1);                    // SFM.py:308-309
    }
    SFM_LAUNCH(ctx, st, "k_ransac_fit", k_ransac_fit<<<ceil_div(iters, 64), 64, 0, st>>>(P));
    if (pose) {
        SFM_LAUNCH(ctx, st, "k_ransac_valid", k_ransac_valid<<<iters, 128, 0, st>>>(P));
    }
    SFM_LAUNCH(ctx, st, "k_ransac_score", k_ransac_score<<<ceil_div(iters, 8), 256, 0, st>>>(P));
    SFM_LAUNCH(ctx, st, "k_ransac_select", k_ransac_select<<<1, 1024, 0, st>>>(P));
    return SFM_OK;
}

// numpy's legacy MT19937 stream (numpy/random/src/mt19937/mt19937.c): state + block generation.
struct Mt19937 {
    uint32_t key[624];
    explicit Mt19937(uint32_t seed) {
        for (int i = 0; i < 624; ++i) { key[i] = seed; seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u; }
    }
    // the next 624 tempered outputs (both loops vectorise: the recurrence reaches 227 elements back)
    void block(uint32_t* out) {
        constexpr uint32_t UP = 0x80000000u, LO = 0x7fffffffu, MAT = 0x9908b0dfu;
        int i = 0;
        for (; i < 624 - 397; ++i) { const uint32_t y = (key[i] & UP) | (key[i + 1] & LO); key[i] = key[i + 397] ^ (y >> 1) ^ (MAT & (0u - (y & 1u))); }
        for (; i < 623; ++i) { const uint32_t y = (key[i] & UP) | (key[i + 1] & LO); key[i] = key[i - 227] ^ (y >> 1) ^ (MAT & (0u - (y & 1u))); }
        const uint32_t yl = (key[623] & UP) | (key[0] & LO);
        key[623] = key[396] ^ (yl >> 1) ^ (MAT & (0u - (yl & 1u)));
        for (i = 0; i < 624; ++i) {
            uint32_t y = key[i];
            y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
            out[i] = y;
        }
    }
};

// Source of stream words for the shuffle loop: a window [pos, end) of w, refilled by more().
struct WordSource {
    const uint32_t* w = nullptr;
    size_t pos = 0, end = 0;
    virtual void more() = 0;
    virtual ~WordSource() {}
};

// The stream depends on the seed only, and the reference always seeds with 5 (SFM.py:45,133): the
// tempered outputs are generated once per process and shared by every call and host thread.  The
// buffer's address space is reserved once (pages are committed as they are written), so growing it
// never moves the words concurrent readers are walking over.
struct StreamCache {
    explicit StreamCache(uint32_t seed) : gen(seed) {}
    std::mutex grow_mu;
    Mt19937 gen;
    std::vector<uint32_t> words;
    std::atomic<size_t> ready{0};
    void grow(size_t want) {
        std::lock_guard<std::mutex> g(grow_mu);
        want = std::min(((want + 623) / 624) * 624, words.capacity() / 624 * 624);
        const size_t old = words.size();
        if (want <= old) return;
        words.resize(want);                                 // within the reserved capacity: no reallocation
        for (size_t o = old; o < want; o += 624) gen.block(words.data() + o);
        ready.store(want, std::memory_order_release);
    }
};
constexpr size_t kCacheMaxWords = (size_t)1 << 28;         // 1 GiB of address space; longer draws stream locally
std::mutex g_caches_mu;
std::vector<std::pair<uint32_t, std::shared_ptr<StreamCache>>> g_caches;

std::shared_ptr<StreamCache> stream_cache(uint32_t seed) {
    std::lock_guard<std::mutex> g(g_caches_mu);
    for (auto& c : g_caches) if (c.first == seed) return c.second;
    auto c = std::make_shared<StreamCache>(seed);
    c->words.reserve(kCacheMaxWords);                       // may throw bad_alloc: caller streams locally instead
    if (g_caches.size() >= 4) g_caches.erase(g_caches.begin());
    g_caches.emplace_back(seed, c);
    return c;
}

struct CachedSource : WordSource {
    std::shared_ptr<StreamCache> c;
    CachedSource(uint32_t seed, size_t expect) : c(stream_cache(seed)) {
        c->grow(expect);
        w = c->words.data();
        end = c->ready.load(std::memory_order_acquire);
    }
    void more() override {
        const size_t have = c->ready.load(std::memory_order_acquire);
        if (have <= end) {
            c->grow(std::max(end + end / 4, end + (size_t)(1 << 20)));
            if (c->ready.load(std::memory_order_acquire) <= end) throw std::bad_alloc();   // reservation exhausted
        }
        end = c->ready.load(std::memory_order_acquire);
    }
};

struct LocalSource : WordSource {
    Mt19937 gen;
    std::vector<uint32_t> buf;
    explicit LocalSource(uint32_t seed) : gen(seed), buf(624 * 128) { w = buf.data(); }
    void more() override {
        for (size_t o = 0; o < buf.size(); o += 624) gen.block(buf.data() + o);
        pos = 0; end = buf.size();
    }
};

// RandomState.choice(n, 8, replace=False) == RandomState.permutation(n)[:8]: a Fisher-Yates shuffle
// of arange(n), i = n-1 .. 1, with j = random_interval(i) drawn by masked rejection (numpy/random/
// src/distributions/distributions.c).
void legacy_choice8(WordSource& src, int n, int iterations, int32_t* out) {
    std::vector<int32_t> ident((size_t)n), perm((size_t)n);
    for (int i = 0; i < n; ++i) ident[i] = i;
    int32_t* __restrict__ pm = perm.data();
    for (int it = 0; it < iterations; ++it) {
        memcpy(pm, ident.data(), (size_t)n * sizeof(int32_t));
        uint32_t i = (uint32_t)n - 1;
        while (i >= 1) {
            uint32_t mask = i;
            mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
            const uint32_t lo = mask >> 1;                  // this level: lo < i <= mask
            while (i > lo) {
                if (src.pos == src.end) src.more();
                const uint32_t* __restrict__ w = src.w + src.pos;
                // Two phases per chunk (keeps the stream scan free of the permutation's loads and
                // stores).  Phase 1: i - lo words can never take i below lo (each accepts at most one
                // step), so that many are scanned without testing the level boundary; word k is
                // accepted iff its value is <= i - (accepted so far).  Phase 2: the swaps.
                const uint32_t cnt = (uint32_t)std::min<size_t>(std::min<size_t>(src.end - src.pos, (size_t)(i - lo)), 512);
                uint32_t js[512];
                uint32_t c = 0;
                for (uint32_t k = 0; k < cnt; ++k) {
                    const uint32_t j = w[k] & mask;
                    js[c] = j;
                    c += (j <= i - c) ? 1u : 0u;
                }
                for (uint32_t q = 0; q < c; ++q) {
                    const uint32_t jj = js[q];
                    const int32_t a = pm[i], b = pm[jj];
                    pm[i] = b; pm[jj] = a;
                    --i;
                }
                src.pos += cnt;
            }
        }
        memcpy(out + (size_t)it * 8, pm, 8 * sizeof(int32_t));
    }
}

}  // namespace

extern "C" {

SFM_EXPORT int sfm_matches_to_coords(SfmCtx* ctx, void* stream, const int32_t* match_dev, const int32_t* count_dev,
                                     const int32_t* x1_dev, const int32_t* y1_dev, const int32_t* x2_dev,
                                     const int32_t* y2_dev, int num_matches, double* p1_out, double* p2_out,
                                     int32_t* n_out) {
    if (!ctx) return SFM_ERR_BAD_ARG;
    if (!match_dev || !count_dev || !x1_dev || !y1_dev || !x2_dev || !y2_dev || !p1_out || !p2_out || !n_out || num_matches < 1)
        return sfm_set_error(ctx, SFM_ERR_BAD_ARG, "matches_to_coords: null pointer or num_matches < 1");
    cudaStream_t st = (cudaStream_t)stream;
    SFM_LAUNCH(ctx, st, "k_matches_to_coords",
               k_matches_to_coords<<<ceil_div(num_matches, 256), 256, 0, st>>>(match_dev, count_dev, x1_dev, y1_dev, x2_dev,
                                                                               y2_dev, num_matches, p1_out, p2_out, n_out));
    return SFM_OK;
}

SFM_EXPORT int sfm_ransac_sample_indices(uint32_t seed, int n, int iterations, int32_t* out_host) {
    if (!out_host || n < 8 || iterations < 0) return SFM_ERR_BAD_ARG;
    // expected stream words: sum over the steps of (mask + 1) / (i + 1)
    double per = 0.0;
    for (uint32_t i = 1; i < (uint32_t)n; ++i) {
        uint32_t mask = i;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        per += (double)((uint64_t)mask + 1) / (double)(i + 1);
    }
    const double expect = per * iterations * 1.01 + 65536.0;
    bool done = false;
    if (expect < (double)kCacheMaxWords * 0.9) {
        try {
            CachedSource src(seed, (size_t)expect);
            legacy_choice8(src, n, iterations, out_host);
            done = true;
        } catch (const std::bad_alloc&) {}                  // no address space for the shared cache: stream locally
    }
    if (!done) {
        try {
            LocalSource src(seed);
            legacy_choice8(src, n, iterations, out_host);
        } catch (const std::bad_alloc&) {
            return SFM_ERR_UNSUPPORTED;
        }
    }
    return SFM_OK;
}

SFM_EXPORT size_t sfm_ransac_workspace_bytes(int iterations) { return iterations > 0 ? ws_layout(iterations).total : 0; }

SFM_EXPORT int sfm_find_inliers(SfmCtx* ctx, void* stream, const double* p1_dev, const double* p2_dev, int n,
                                const int32_t* samples_dev, int iterations, double threshold, void* workspace_dev,
                                size_t workspace_bytes, int32_t* inlier_idx_out, int32_t* result_out, double* f_out) {
    return run_ransac(ctx, stream, p1_dev, p2_dev, n, samples_dev, iterations, threshold, 0, nullptr, nullptr, nullptr,
                      nullptr, workspace_dev, workspace_bytes, inlier_idx_out, result_out, f_out);
}

SFM_EXPORT int sfm_ransac_camera_motion(SfmCtx* ctx, void* stream, const double* p1_dev, const double* p2_dev, int n,
                                        const double* K1, const double* K2, const double* R_base, const double* T_base,
                                        const int32_t* samples_dev, int iterations, double threshold,
                                        void* workspace_dev, size_t workspace_bytes, int32_t* inlier_idx_out,
                                        int32_t* result_out, double* pose_out) {
    return run_ransac(ctx, stream, p1_dev, p2_dev, n, samples_dev, iterations, threshold, 1, K1, K2, R_base, T_base,
                      workspace_dev, workspace_bytes, inlier_idx_out, result_out, pose_out);
}

// Per-hypothesis data of the last call on this workspace (parity tests): device pointers into it.
SFM_EXPORT int sfm_ransac_debug_views(void* workspace_dev, int iterations, double** f_dev, int32_t** counts_dev,
                                      uint32_t** valid_dev, double** cand_dev) {
    if (!workspace_dev || iterations < 1) return SFM_ERR_BAD_ARG;
    const WsLayout w = ws_layout(iterations);
    char* base = (char*)workspace_dev;
    if (f_dev) *f_dev = (double*)(base + w.F);
    if (counts_dev) *counts_dev = (int32_t*)(base + w.counts);
    if (valid_dev) *valid_dev = (uint32_t*)(base + w.valid);
    if (cand_dev) *cand_dev = (double*)(base + w.cand);
    return SFM_OK;
}

}  // extern "C"
