// extract.cuh -- plan / workspace layout shared by the extraction kernels.
#pragma once
#include "common.cuh"

#define SFM_HIST1_BINS 4096      // radix-select pass 1 (fused into k_harris): key >> 20
#define SFM_MAX_FW 32            // feature_width / window side upper bound

struct SegState {                // one per (image, level)
    uint32_t prefix[2];          // 12-bit bucket (key >> 20) holding each of the two median ranks
    uint32_t rank[2];            // remaining rank inside that bucket
    float    median;
    uint32_t n_cand;             // candidates appended by NMS (may exceed cap)
    uint32_t n_sel;              // selected, border-valid keypoints
    uint32_t med_cnt;            // keys of bucket prefix[0] compacted by k_nms (may exceed cap)
    uint32_t min1;               // smallest key of bucket prefix[1] when it differs from prefix[0]
    uint32_t pad[7];
};

struct LevelInfo {
    int H, W;
    int fw, hw;                  // feature width at this level and fw / 2
    int k;                       // per-level top-k
    int cand_cap;                // candidate slots
    int resize_mode;             // 0: level 0, 1: exact 2x2 mean, 2: bilinear
    int sel_off;                 // offset of this level inside the per-image sel block
    int med_cap;                 // slots of the median bucket list
    int pad1;
    long long med_off;           // u32 offset inside the per-image median-list block
    long long img_off;           // float offset inside the per-image pyramid block (level >= 1)
    long long r_off;             // float offset inside the per-image R block
    long long cand_off;          // u64 offset inside the per-image candidate block
    double scale;                // pyramid_scale_factor ** level
    double inv_x, inv_y;         // src/dst size ratios for the bilinear resize
};

struct ExtractPlan {
    int B, L, H0, W0;
    int nms_half, G, rot, pad0;
    float alpha;
    int sel_stride;              // sum of per-level k
    long long pyr_stride, r_stride, cand_stride, med_stride;
    LevelInfo lv[SFM_MAX_LEVELS];
    // workspace
    const float* images;
    float* pyr;
    float* R;
    uint32_t* hist1;             // [S][4096]
    uint32_t* med;               // [B][med_stride] keys of the median bucket
    SegState* seg;               // [S]
    unsigned long long* cand;    // [B][cand_stride]
    unsigned long long* sel;     // [B][sel_stride]
    int* flags;                  // [0]: candidate overflow
    double e9[9];                // np.linspace(-pi, pi, 9)
    double e37[37];              // np.linspace(-pi, pi, 37)
    float ef37[37];              // smallest float32 >= e37[i]: (double)o >= e37[i]  <=>  o >= ef37[i] for a float32 o
    float ef37_top;              // largest float32 <= e37[36]
    // slot_thr[b][k], k < 8: the smallest float32 o with (double)o - dom_b >= e9[k] (dom_b = centre of orientation bin b;
    // row 36: no rotation, dom = 0); [b][8]: the largest float32 o with (double)o - dom_b <= e9[8].  The 8-bin
    // histogram's float64 edge tests on a float32 orientation, decided in float32.
    float slot_thr[37][10];
    double atan_poly[19];        // coefficients of atan2_f32 (constant-bank operands of its DFMAs)
};

// Window weights, one row of the G x G kernel per 16 floats (64-byte aligned rows: the rolled
// tap-row loop of k_harris fetches a row with vector constant loads).
#define SFM_GW_PITCH 16
struct __align__(16) GaussWeights {
    float w[SFM_MAX_GAUSS * SFM_GW_PITCH];
    // wp[jj][dx] = (w[jj][dx], w[jj-1][dx]) for jj = 1..G-1: the packed (upper row, lower row) weight
    // pair of product row jj, consumed as one 64-bit uniform operand by the FFMA2 variant
    float2 wp[SFM_MAX_GAUSS * SFM_GW_PITCH];
};
// Window weights paired for the two output rows a thread owns: entry [jj][dx] =
// (w[jj][dx] or 0 when jj == G,  w[jj-1][dx] or 0 when jj == 0), jj = 0..G.
struct GaussPairs { float2 w[(SFM_MAX_GAUSS + 1) * SFM_MAX_GAUSS]; };

struct ExtractOut {
    int32_t *x, *y, *lx, *ly, *level;
    float *conf, *desc;
    int32_t* count;
    int cap;
};
