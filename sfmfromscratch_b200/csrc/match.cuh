// match.cuh -- plan / workspace layout of the NN-ratio matcher.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

constexpr int MT_ROWS = 256;          // query rows per work unit (two M=128 MMA row tiles)
constexpr int MT_COLS = 256;          // train columns per MMA tile (UMMA N)
constexpr int MT_SUB = 4;             // columns the re-check evaluates at a time (a group is a multiple)
constexpr int MT_TOPK = 4;            // candidate groups kept per list
constexpr int MT_IDX_BITS = 10;       // packed into the low mantissa bits of a key: tile_local << GROUP_BITS | group within the list's half tile
// Columns per candidate group: the tensor-core epilogue keeps the minimum of every GROUP columns, the re-check gathers
// whole groups.  8 makes the epilogue cheap enough to hide behind the MMAs of large sets (8192 x 8192: k_match_tc
// 0.82 ms against 0.97 with 4); 4 halves what the re-check gathers, which is what counts for sets of a few thousand
// rows (31 pairs of ~1.7 k x 1.7 k: 0.229 -> 0.189 ms).  Chosen per call from nmax (mt_group_for).
template <int GRP> struct MtG {
    static constexpr int GROUP = GRP;
    static constexpr int GROUPS_PER_HALF = (MT_COLS / 2) / GRP;
    static constexpr int GROUP_BITS = GROUPS_PER_HALF == 32 ? 5 : (GROUPS_PER_HALF == 16 ? 4 : 3);
    static constexpr int MAX_TILES = 1 << (MT_IDX_BITS - GROUP_BITS);       // column tiles per split
    static_assert(GRP == 4 || GRP == 8 || GRP == 16, "group size");
};
inline int mt_group_for(int nmax_pad) { return nmax_pad <= 4096 ? 4 : 8; }
inline int mt_max_tiles(int group) { return group == 4 ? MtG<4>::MAX_TILES : MtG<8>::MAX_TILES; }
constexpr uint32_t MT_IDX_MASK = (1u << MT_IDX_BITS) - 1u;
constexpr float MT_SENTINEL = 1.0e30f;   // norm term of padding rows: never a candidate
constexpr float MT_INVALID = 1.0e29f;    // packed values >= this are empty slots
constexpr int MX_ROWS = 8;            // flagged rows per exact-scan work item
constexpr int MX_COLS = 1024;         // columns per exact-scan work item
constexpr int MT_MAX_SPLITS = 16;

struct MatchPlan {
    int n_sets, nmax, nmax_pad, n_pairs;
    int p0, pn, woff, no_prune;       // pair chunk [p0, p0+pn) this launch covers; offset into work_off; validation flag
    int n_splits, tiles_per_split, n_tiles, n_lists;
    int n_xchunks;                    // exact-scan column chunks
    int group;                        // columns per candidate group of this call (4 or 8)
    int mode, cap;
    float thr;
    // inputs
    const float* desc;                // batch: [n_sets][nmax][128]; single pair: unused
    const float* f1; const float* f2; // single pair
    const int32_t* counts_in;         // batch
    int n1, n2;                       // single pair
    const int32_t* pairs_in;          // batch
    // workspace
    const float** set_ptr;            // [n_sets]
    int32_t* set_cnt;                 // [n_sets]
    int32_t* pairs;                   // [n_pairs][2]
    __half* h16;                      // [n_sets][nmax_pad][128]
    float* nb;                        // [n_sets][nmax_pad]   |b|^2 (float32), sentinel on padding
    float* hatn;                      // [n_sets][nmax_pad]   |fp16(b)|
    float* resn;                      // [n_sets][nmax_pad]   |b - fp16(b)|
    float* setmax;                    // [n_sets][4]  max hatn, max resn, max nb
    uint32_t* cands;                  // [n_pairs][nmax_pad][n_lists][4]
    int32_t* res_idx;                 // [n_pairs][nmax]
    float* res_d0;                    // squared distances
    float* res_d1;
    int32_t* flag_cnt;                // [n_pairs]
    int32_t* flag_rows;               // [n_pairs][nmax]
    int32_t* work_off;                // [n_pairs + 1] exact-scan work prefix
    float4* part;                     // [n_pairs][nmax][n_xchunks]  (d0, idx0, d1, -)
    unsigned long long* mkeys;        // [n_pairs][nmax]
    int32_t* midx;                    // [n_pairs][nmax]
    int32_t* mcount;                  // [n_pairs]
    int32_t* stats;                   // [n_pairs][2] groups re-checked (internal)
};
