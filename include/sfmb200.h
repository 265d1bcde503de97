/*
 * sfmb200.h -- C ABI of libsfmb200.so, the B200 (sm_100a) implementation of
 * the SfmFromScratch feature hot path.
 *
 * The reference is pure Python and has no FFI; the interfaces this boundary
 * replaces are the Python methods a binding would route here (paths relative
 * to the reference root):
 *
 *   sfm_extract_batch        FeatureExtractor/SIFT/ScaleRotInvSIFT.py:9-16,89-115
 *                            (ScaleRotInvSIFT.__init__ -> _build_image_pyramid +
 *                            compute) and, with rotation_invariant = 0 and
 *                            pyramid_level = 1, FeatureExtractor/SIFT/
 *                            NaiveSIFT.py:42-52 (detect_keypoints +
 *                            extract_descriptors)
 *   sfm_harris_response      FeatureExtractor/SIFT/NaiveSIFT.py:60-74 (R map only;
 *                            exposed for parity tests)
 *   sfm_ingest_rgb8          Runner.py:33-46 with :467-493,:551-563 (_load_image ->
 *                            _PIL_resize -> _rgb2gray), from the decoded 8-bit image
 *   sfm_match_ratio          FeatureMatcher/NNRatioFeatureMatcher.py:8-60
 *                            (match_features_ratio_test)
 *   sfm_match_ratio_batch    the same method over a list of image pairs (the
 *                            reference loops over consecutive pairs at
 *                            Runner.py:183-191,344-347)
 *   sfm_matches_to_coords    Runner.py:423-434 (_convert_matches_to_coords)
 *   sfm_ransac_sample_indices  the np.random.seed(5) / np.random.choice(n, 8,
 *                            replace=False) draws of SFM.py:45-49,133-137
 *   sfm_associate_nearest    Runner.py:241-247 (nearest already-triangulated 2-D point of
 *                            every previous-frame match, SFM.py:376-382 distances)
 *   sfm_dedup_points         Runner.py:361-385 (add_points / is_new_point /
 *                            find_existing_point)
 *   sfm_find_inliers         SFM.py:126-160 (CameraPose.find_inliers)
 *   sfm_ransac_camera_motion SFM.py:38-124 (CameraPose.ransac_camera_motion with
 *                            _check_valid_pose)
 *
 * Conventions
 *   - Every pointer marked "dev" is CUDA device memory owned by the caller (the
 *     Python host side allocates it as torch tensors); the library never frees
 *     caller memory and keeps no reference after the call is enqueued, except
 *     for `workspace`, which must stay alive until the stream has drained.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it
 *     and the calls do not synchronise unless stated.
 *   - Return value: 0 on success, a negative SfmStatus otherwise;
 *     sfm_last_error(ctx) returns a description for the calling thread.
 *   - A context may be used from several host threads concurrently (the
 *     reference calls the extractor from an 8-thread pool, Runner.py:186-191):
 *     calls carry their own workspace and stream, the context holds only
 *     immutable device properties and a mutex-protected error slot.
 *   - There is no CPU fallback: without a CUDA device every entry point other
 *     than sfm_version / sfm_last_error fails with SFM_ERR_CUDA.
 */
#ifndef SFMB200_H_
#define SFMB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SFM_EXPORT __attribute__((visibility("default")))
#else
#define SFM_EXPORT
#endif

#define SFM_API_VERSION 1
#define SFM_DESC_DIM 128      /* descriptor length (4 x 4 cells x 8 bins) */
#define SFM_MAX_LEVELS 8      /* pyramid_level upper bound */
#define SFM_MAX_GAUSS 11      /* gaussian_size upper bound (odd); cv2.filter2D
                                 switches to a DFT path at 11x11 and above in
                                 area >= 130, which the reference's defaults
                                 never reach */

typedef enum SfmStatus {
    SFM_OK = 0,
    SFM_ERR_BAD_ARG = -1,
    SFM_ERR_CUDA = -2,
    SFM_ERR_WORKSPACE = -3,   /* workspace_bytes smaller than the query says */
    SFM_ERR_CAPACITY = -4,    /* reported by sfm_extract_status: a per-level
                                 candidate buffer overflowed (plateau image);
                                 call again with cand_full = 1 */
    SFM_ERR_UNSUPPORTED = -5
} SfmStatus;

typedef struct SfmCtx SfmCtx;

/* Extractor parameters: the keys of the reference's extractor_params dict
 * (main.py:19-28; FeatureExtractor.py:11; NaiveSIFT.py:35-39;
 * ScaleRotInvSIFT.py:12-13) with the same defaults. */
typedef struct SfmExtractParams {
    int32_t num_interest_points;   /* 2500 */
    int32_t ksize;                 /* 7: NMS window = 2*(ksize/2)+1 */
    int32_t gaussian_size;         /* 7, odd, <= SFM_MAX_GAUSS */
    double  sigma;                 /* 5 */
    double  alpha;                 /* 0.05 (applied as float32) */
    int32_t feature_width;         /* 16 */
    int32_t pyramid_level;         /* 4; 1 for NaiveSIFT */
    double  pyramid_scale_factor;  /* 2 */
    int32_t rotation_invariant;    /* 1: ScaleRotInvSIFT descriptors, 0: NaiveSIFT */
    int32_t split_k_by_level;      /* 1: per-level k = int(k / levels)
                                      (ScaleRotInvSIFT.py:90); 0: k */
    int32_t cand_full;             /* 0: candidate buffers sized by the NMS
                                      density bound; 1: one slot per pixel */
    /* Optional HOST pointer to gaussian_size^2 float32 window weights in
     * row-major order.  NULL: the library evaluates NaiveSIFT.py:175-199 in
     * double precision with libm exp().  The Python host side passes the
     * numpy-computed kernel so the weights are the reference's to the bit. */
    const float* gauss_weights;
} SfmExtractParams;

SFM_EXPORT int  sfm_version(void);
SFM_EXPORT int  sfm_ctx_create(int device, SfmCtx** out);
SFM_EXPORT void sfm_ctx_destroy(SfmCtx* ctx);
SFM_EXPORT const char* sfm_last_error(SfmCtx* ctx);
/* Number of SMs of the context's device (grid sizing for callers/tests). */
SFM_EXPORT int  sfm_ctx_sm_count(SfmCtx* ctx);

/* Kernel launches issued through this context so far (bench.py's gpu_launches). */
SFM_EXPORT unsigned long long sfm_ctx_launch_count(SfmCtx* ctx);

/* Per-context tuning.  Results never depend on an option; only which kernel computes them does.
 *   SFM_OPT_HARRIS_STREAM_MIN_BANDS  a pyramid level goes to the persistent Harris stream (k_harris_stream) when it
 *       has at least this many 64x16 bands per SM, otherwise to the one-tile-per-CTA kernel (default 24; 0 sends
 *       every level the stream can take -- the tests use it to run the stream on small and odd shapes). */
#define SFM_OPT_HARRIS_STREAM_MIN_BANDS 1
SFM_EXPORT int sfm_ctx_set_option(SfmCtx* ctx, int option, int value);

/* Optional per-kernel timing: while enabled, every kernel the library launches
 * is bracketed by two CUDA events on the launching stream.  sfm_profile_collect
 * waits for the recorded events, aggregates them by kernel name into out[0..cap)
 * and returns the number of distinct kernels (records are consumed). */
typedef struct SfmKernelStat {
    char    name[48];
    int32_t launches;
    float   total_ms;
} SfmKernelStat;
SFM_EXPORT int sfm_profile_enable(SfmCtx* ctx, int on);
SFM_EXPORT int sfm_profile_collect(SfmCtx* ctx, SfmKernelStat* out, int cap);

/* Fill p with the reference defaults. */
SFM_EXPORT void sfm_extract_default_params(SfmExtractParams* p);

/* Upper bound on keypoints per image for these parameters
 * (levels * per-level k): the row capacity the output arrays need. */
SFM_EXPORT int  sfm_extract_max_keypoints(const SfmExtractParams* p);

/*
 * HOST function (no device work): the float32 decision tables of the descriptor stage, for inspection and tests.
 * The reference bins float32 orientations against float64 edges (ScaleRotInvSIFT.py:66-87: np.histogram with
 * np.linspace(-pi, pi, 37) and, after the float64 shift by the dominant bin's centre, np.linspace(-pi, pi, 9));
 * the kernel decides the same tests on float32 thresholds:
 *   ef37[i], i < 37:      the smallest float32 o with (double)o >= linspace37[i];  ef37[37]: the largest float32 o
 *                         with (double)o <= linspace37[36]
 *   slot_thr[b][k], k < 8: the smallest float32 o with (double)o - centre_b >= linspace9[k]  (b < 36: dominant
 *                         bin b; b = 36: no rotation, centre 0);  [b][8]: the largest float32 o with
 *                         (double)o - centre_b <= linspace9[8];  [b][9]: +inf (padding)
 * ef37 has 38 floats, slot_thr 37 * 10.
 */
SFM_EXPORT int sfm_describe_tables(float* ef37, float* slot_thr);

/* Workspace size in bytes for a batch of B images of H x W. */
SFM_EXPORT size_t sfm_extract_workspace_bytes(int B, int H, int W, const SfmExtractParams* p);

/*
 * Extract keypoints and descriptors for B grayscale float32 images
 * (images_dev: [B][H][W], row stride W).  Outputs, all dev, capacity `cap`
 * rows per image (cap >= sfm_extract_max_keypoints):
 *   x_out, y_out   [B][cap] int32  level-0 coordinates, (x * scale).astype(int)
 *   lx_out, ly_out [B][cap] int32  coordinates inside the pyramid level (may be NULL)
 *   level_out      [B][cap] int32  pyramid level (may be NULL)
 *   conf_out       [B][cap] float  Harris response (may be NULL)
 *   desc_out       [B][cap][128] float
 *   count_out      [B] int32       keypoints written for each image
 * Keypoints are ordered as the reference orders them: level by level, inside
 * a level by response descending (ties: row-major pixel index ascending).
 */
SFM_EXPORT int sfm_extract_batch(SfmCtx* ctx, void* stream, const float* images_dev, int B, int H, int W,
                      const SfmExtractParams* p, void* workspace_dev, size_t workspace_bytes,
                      int32_t* x_out, int32_t* y_out, int32_t* lx_out, int32_t* ly_out,
                      int32_t* level_out, float* conf_out, float* desc_out, int32_t* count_out,
                      int cap);

/* After the stream has been synchronised: 0, or SFM_ERR_CAPACITY when a
 * candidate buffer of the last sfm_extract_batch call on this workspace
 * overflowed (reads one flag word back from the workspace). */
SFM_EXPORT int sfm_extract_status(SfmCtx* ctx, void* stream, const void* workspace_dev);

/* Harris response map of one H x W image (NaiveSIFT.py:60-74): r_out [H][W]. */
SFM_EXPORT int sfm_harris_response(SfmCtx* ctx, void* stream, const float* image_dev, int H, int W,
                        const SfmExtractParams* p, float* r_out);

/* ---- ingest -------------------------------- */

/*
 * Runner.py:33-46: _load_image (:551-563) -> _PIL_resize (:481-493, PIL's default BICUBIC on
 * RGB) -> _rgb2gray (:467-478), starting from the decoded 8-bit image.
 * rgb_dev [B][H][W][3] uint8 (dev) -> gray_out [B][out_h][out_w] float32 (dev), bit-identical
 * to the array the reference hands to its extractor.  (out_w, out_h) is the reference's
 * (int(W * scale_factor), int(H * scale_factor)).
 */
SFM_EXPORT size_t sfm_ingest_workspace_bytes(int B, int H, int W, int out_h, int out_w);
SFM_EXPORT int sfm_ingest_rgb8(SfmCtx* ctx, void* stream, const uint8_t* rgb_dev, int B, int H, int W,
                               int out_h, int out_w, void* workspace_dev, size_t workspace_bytes,
                               float* gray_out);

/* ---- matching ---------------------------------------------------------- */

typedef enum SfmMatchMode {
    SFM_MATCH_AUTO = 0,     /* tcgen05 fp16 candidate pass + exact float32
                               re-check (+ exact scan of rows whose error bound
                               cannot certify the candidates) */
    SFM_MATCH_EXACT = 1,    /* exact float32 scan of every row (validation) */
    SFM_MATCH_PREPARED = 16,/* flag, OR-ed into the mode of sfm_match_ratio_batch: the
                               workspace already holds the per-set preparation (fp16
                               copy, norms) of an earlier call with the same desc_dev,
                               counts_dev, n_sets and nmax -- only the pair list is new.
                               All-pairs matching in chunks (configs[4]) prepares the
                               512 sets once instead of once per chunk. */
    SFM_MATCH_NO_PRUNE = 32 /* flag (sfm_match_ratio_batch, validation): re-check every row
                               exactly instead of skipping the rows whose ratio test the
                               approximate keys already decide.  Same results, more work:
                               scripts/check_config5.py measures what the prune saves. */
} SfmMatchMode;

/* Workspace bytes for matching n_sets descriptor sets of at most nmax rows over
 * n_pairs pairs. */
SFM_EXPORT size_t sfm_match_workspace_bytes(int n_sets, int nmax, int n_pairs);

/* Leading bytes of the workspace that hold the per-set preparation: they depend on
 * (n_sets, nmax) only, so one workspace sized for the largest pair chunk serves every
 * later SFM_MATCH_PREPARED call over the same sets. */
SFM_EXPORT size_t sfm_match_prepared_bytes(int n_sets, int nmax);

/*
 * One pair.  f1_dev [n1][128], f2_dev [n2][128] float32 row-major, n2 >= 2.
 * Outputs (dev): match_out [cap][2] int32 (index into f1, index into f2),
 * conf_out [cap] float32 (d_nearest / d_second), count_out [1] int32, ordered
 * by confidence ascending (ties: f1 index ascending).  cap >= n1 always fits.
 */
SFM_EXPORT int sfm_match_ratio(SfmCtx* ctx, void* stream, const float* f1_dev, int n1, const float* f2_dev,
                    int n2, int dim, float ratio_threshold, int mode, void* workspace_dev,
                    size_t workspace_bytes, int32_t* match_out, float* conf_out,
                    int32_t* count_out, int cap);

/*
 * Many pairs over a table of descriptor sets.  desc_dev [n_sets][nmax][128]
 * float32, counts_dev [n_sets] int32 (rows used per set, each >= 2 when the set
 * is used as a train set), pairs_dev [n_pairs][2] int32 (query set, train set).  Outputs as above with a leading
 * pair dimension: match_out [n_pairs][cap][2], conf_out [n_pairs][cap],
 * count_out [n_pairs].  stats_out (dev, may be NULL) [n_pairs][2] int32:
 * rows re-scanned exactly, candidate groups re-checked.  Any n_pairs is accepted
 * (batches above 65 535 pairs run as consecutive chunks on the stream).  A pair
 * whose set id lies outside [0, n_sets) is matched as a pair of empty sets:
 * count_out is 0 for it and nothing is read out of bounds.
 */
SFM_EXPORT int sfm_match_ratio_batch(SfmCtx* ctx, void* stream, const float* desc_dev,
                          const int32_t* counts_dev, int n_sets, int nmax,
                          const int32_t* pairs_dev, int n_pairs, float ratio_threshold, int mode,
                          void* workspace_dev, size_t workspace_bytes, int32_t* match_out,
                          float* conf_out, int32_t* count_out, int32_t* stats_out, int cap);

/* ---- two-view RANSAC ------------------------ */

/*
 * Runner.py:423-434 (_convert_matches_to_coords): the first min(count, num_matches) rows of a
 * match list as two coordinate arrays.  match_dev [*][2] int32 and count_dev [1] int32 are the
 * matcher's outputs; x1/y1/x2/y2 the extractor's level-0 coordinates of the two images (dev,
 * int32).  p1_out, p2_out [num_matches][2] float64 (dev); n_out [1] int32 (dev) receives the
 * number of rows written.
 */
SFM_EXPORT int sfm_matches_to_coords(SfmCtx* ctx, void* stream, const int32_t* match_dev, const int32_t* count_dev,
                                     const int32_t* x1_dev, const int32_t* y1_dev, const int32_t* x2_dev,
                                     const int32_t* y2_dev, int num_matches, double* p1_out, double* p2_out,
                                     int32_t* n_out);

/*
 * HOST function (no device work): the `iterations` 8-subsets that
 *     np.random.seed(seed); [np.random.choice(n, 8, replace=False) for _ in range(iterations)]
 * draws (SFM.py:45-49,133-137 with seed 5) -- numpy's legacy MT19937 stream, one Fisher-Yates
 * permutation of n per draw with masked rejection sampling.  out_host [iterations][8] int32.
 */
SFM_EXPORT int sfm_ransac_sample_indices(uint32_t seed, int n, int iterations, int32_t* out_host);

SFM_EXPORT size_t sfm_ransac_workspace_bytes(int iterations);

/*
 * SFM.py:126-160.  p1_dev, p2_dev [n][2] float64 (dev, n >= 8), samples_dev [iterations][8] int32
 * (dev).  Outputs (dev): result_out [4] int32 = {winning hypothesis or -1, its inlier count, 0, 0};
 * inlier_idx_out [n] int32, the winner's inlier rows in ascending order (the reference returns
 * p1[mask], p2[mask]); f_out [9] float64 (may be NULL), the winner's fundamental matrix.
 * The winner is the first hypothesis with the largest inlier count, as in the reference.
 * A sample is degenerate when its 8x9 design matrix is numerically rank-deficient (|R| diagonal of its QR
 * spanning more than 1e10): the null space then has several dimensions and the vector LAPACK returns --
 * hence the reference's F, inlier count and possibly its winner -- is decided by rounding noise.  With
 * result_out[3] == 0 the outputs are the reference's; otherwise they are one valid outcome of the same
 * procedure (typical cause: a pair without relative motion, p1 == p2 up to a shift).
 */
SFM_EXPORT int sfm_find_inliers(SfmCtx* ctx, void* stream, const double* p1_dev, const double* p2_dev, int n,
                                const int32_t* samples_dev, int iterations, double threshold, void* workspace_dev,
                                size_t workspace_bytes, int32_t* inlier_idx_out, int32_t* result_out, double* f_out);

/*
 * SFM.py:38-102 with _check_valid_pose (:104-124).  K1, K2, R_base [9] row-major and T_base [3]
 * are HOST pointers (copied into the launch).  Outputs as sfm_find_inliers, with result_out[2] =
 * bit mask of the winner's pose candidates that pass the cheirality test, and pose_out [9 + 48]
 * float64 (dev) = the winner's F followed by its four candidates (R row-major, T) in the
 * canonical order (Ra,T), (Ra,-T), (Rb,T), (Rb,-T).  The reference tries the same four in the order
 * LAPACK's sign choices for svd(E) imply and keeps the first valid one; the Python host layer
 * re-derives that order from the returned F.
 */
SFM_EXPORT int sfm_ransac_camera_motion(SfmCtx* ctx, void* stream, const double* p1_dev, const double* p2_dev, int n,
                                        const double* K1, const double* K2, const double* R_base, const double* T_base,
                                        const int32_t* samples_dev, int iterations, double threshold,
                                        void* workspace_dev, size_t workspace_bytes, int32_t* inlier_idx_out,
                                        int32_t* result_out, double* pose_out);

/* Device pointers to the per-hypothesis data the last call left in the workspace (parity tests):
 * F [iterations][9], counts [iterations], valid [iterations], candidates [iterations][48]. */
SFM_EXPORT int sfm_ransac_debug_views(void* workspace_dev, int iterations, double** f_dev, int32_t** counts_dev,
                                      uint32_t** valid_dev, double** cand_dev);

/* ---- point association ----------------------- */

/*
 * Runner.py:241-247.  ref_dev [m][2], query_dev [q][2] float64 (dev).  For every query row:
 * nearest_out [q] int32 = np.argmin of its distances to the ref rows (first minimum), dist_out [q]
 * float64 (may be NULL) that distance, flag_out [q] int32 = distance < dist_threshold; kept_out [q]
 * int32 = the flagged query rows in ascending order, count_out [1] their number.
 */
SFM_EXPORT int sfm_associate_nearest(SfmCtx* ctx, void* stream, const double* ref_dev, int m, const double* query_dev,
                                     int q, double dist_threshold, int32_t* nearest_out, double* dist_out,
                                     int32_t* flag_out, int32_t* kept_out, int32_t* count_out);

/*
 * Runner.py:361-385 for one batch: pts_dev [n][3] float64 are visited in order against the global
 * store store_dev [e][3] (e may be 0).  A point with no stored point closer than `threshold`
 * (existing, or an earlier new point of the batch) is new and takes the next store slot; any other
 * maps to the first nearest stored point.  Outputs (dev): index_out [n] int32 store index per point
 * (new points: e, e+1, ... in batch order), is_new_out [n] int32, n_new_out [1] int32.
 * pair_cap bounds the number of (earlier point, point) pairs closer than the threshold that the
 * workspace can list (duplicates inside a batch are rare: n is a safe default); if it is exceeded
 * n_new_out is set to -1 and index_out[0] to the capacity needed, and the caller repeats the call.
 */
SFM_EXPORT size_t sfm_dedup_workspace_bytes(int n, int pair_cap);
SFM_EXPORT int sfm_dedup_points(SfmCtx* ctx, void* stream, const double* pts_dev, int n, const double* store_dev, int e,
                                double threshold, int pair_cap, void* workspace_dev, size_t workspace_bytes,
                                int32_t* index_out, int32_t* is_new_out, int32_t* n_new_out);

#ifdef __cplusplus
}
#endif
#endif /* SFMB200_H_ */
