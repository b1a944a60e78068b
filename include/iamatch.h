/*
 * iamatch.h — C-ABI of libiamatch.so, the B200-native (sm_100a) pairwise
 * feature-matching engine that sits behind the `matcher` module of
 * NorthStarUAS/ImageAnalysis.
 *
 * The reference has no FFI layer of its own: its numeric seam is the Python
 * call `the_matcher.knnMatch(des1, des2, k)` inside `raw_matches`
 * (reference scripts/lib/matcher.py:203-216, matcher object built at :62-79)
 * and the Python reductions that follow it.  Every entry point below names
 * the reference code it stands in for.  Plain C types only; the caller owns
 * all host buffers; the library owns device memory until iam_destroy().
 *
 * Error convention: every function returns 0 on success or a negative
 * IAM_E_* code; iam_last_error() returns a human-readable string for the
 * last failure on the calling thread.  There is no CPU fallback: if no
 * CUDA device / kernel image is usable the call fails with IAM_E_CUDA.
 */
#ifndef IAMATCH_H
#define IAMATCH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums ----------------------------------------------------------- */

/* Norm selected by matcher.configure() (matcher.py:49-56):
 * SIFT/SURF -> cv2.NORM_L2, ORB/Star -> cv2.NORM_HAMMING. */
#define IAM_NORM_L2       0
#define IAM_NORM_HAMMING  1

/* Host descriptor element types accepted by iam_upload_descriptors():
 * OpenCV SIFT hands back float32 [N,128] (integer valued, 0..255),
 * ORB uint8 [N,32] (image.py:160-180 loads exactly those). */
#define IAM_DTYPE_U8   0
#define IAM_DTYPE_F32  1

/* kNN engine selection (debug / validation aid; IAM_ENGINE_AUTO in
 * production).  UMMA = tcgen05 tensor-core kernel, SIMT = exact CUDA-core
 * kernel used as the on-device cross-check.  For L2 the tensor-core kernel
 * has two operand kinds: integer-valued descriptors (what cv2.SIFT produces,
 * image.py:324) run on byte operands (kind::i8, s32 accumulate, exact);
 * anything else (SURF, RootSIFT, norms beyond the byte layout's capacity)
 * on fp16 operands (kind::f16).  UMMA_F16 forces the latter. */
#define IAM_ENGINE_AUTO      0
#define IAM_ENGINE_UMMA      1
#define IAM_ENGINE_SIMT      2
#define IAM_ENGINE_UMMA_F16  3

/* tensor-core operand kind reported in iam_timing.mma_kind */
#define IAM_KIND_F16  0
#define IAM_KIND_F8   1
#define IAM_KIND_I8   2

/* Reduction applied to the k=2 neighbour lists of one direction. */
#define IAM_REDUCE_LOWE        0  /* keep d0 <= d1*ratio          (find_obj.py:50-60, matcher.py:227) */
#define IAM_REDUCE_REF_METRIC  1  /* metric=d0*(d0/d1) sorted, < max_distance*ratio, best `cap`  (matcher.py:253-269) */

/* RANSAC models of filter_by_transform (matcher.py:121-128). */
#define IAM_MODEL_ESSENTIAL    0   /* cv2.findEssentialMat  (matcher.py:126): 5-point, Sampson error          */
#define IAM_MODEL_HOMOGRAPHY   1   /* cv2.findHomography    (matcher.py:122): 4-point, transfer error         */
#define IAM_MODEL_FUNDAMENTAL  2   /* cv2.findFundamentalMat (matcher.py:124): 7-point, epipolar-line distance */
#define IAM_MODEL_AFFINE_PARTIAL 3 /* cv2.estimateAffinePartial2D (smart.py:88-89): 2-point similarity, transfer error,
                                      least-squares polish on the inliers; rows 0-1 of the 3 x 3 output are the 2 x 3 matrix */

#define IAM_OK           0
#define IAM_E_ARG       -1
#define IAM_E_CUDA      -2
#define IAM_E_NOMEM     -3
#define IAM_E_STATE     -4
#define IAM_E_UNSUPPORTED -5

typedef struct iam_ctx iam_ctx;

/* ---- lifetime -------------------------------------------------------- */

/* Replaces the matcher construction in matcher.configure() (matcher.py:62-79).
 * `desc_bytes` is the descriptor row size in bytes as OpenCV stores it for
 * u8 data (128 for SIFT-as-u8, 32 for ORB); for float32 SIFT pass 128 too
 * (the element count).  Only 128 (L2) and 32 (Hamming) take the tensor-core
 * path; other sizes up to 128 use the SIMT engine. */
int iam_create(int device, int norm, int desc_bytes, iam_ctx** out);
int iam_destroy(iam_ctx* ctx);
const char* iam_last_error(void);
/* Library/ABI version, bumped on any signature change. */
int iam_abi_version(void);

/* Use an externally owned CUDA stream (cudaStream_t passed as void*), e.g.
 * torch.cuda.current_stream().cuda_stream, so that callers can bracket work
 * with their own events.  NULL restores the context's private stream. */
int iam_set_stream(iam_ctx* ctx, void* cuda_stream);
int iam_set_engine(iam_ctx* ctx, int engine);
int iam_synchronize(iam_ctx* ctx);

/* ---- descriptors ----------------------------------------------------- */

/* Stands in for `np.array(i1.des_list)` being handed to knnMatch
 * (matcher.py:212-213).  Copies `n` descriptors from HOST memory `ptr`
 * (row-major, `desc_bytes` elements per row of `dtype`) to the device and
 * converts them into the tiled tensor-core operand layout.  `image_id` is a
 * small non-negative integer chosen by the caller (index into
 * proj.image_list).  Re-uploading an id replaces it.  `pinned` != 0 promises
 * that `ptr` is page-locked so the copy can be asynchronous. */
int iam_upload_descriptors(iam_ctx* ctx, int image_id, const void* ptr,
                           int n, int dtype, int pinned);
/* Same, but `dptr` already lives in device memory (no H2D copy). */
int iam_upload_descriptors_device(iam_ctx* ctx, int image_id, const void* dptr,
                                  int n, int dtype);
/* Keypoint position ids for filter_duplicates (matcher.py:157-182): two
 * keypoints of one image get the same id iff their '%.2f-%.2f' pixel keys
 * (matcher.py:165-166) are equal.  `keys` is a HOST array [n] with values in
 * [0, n); the image's descriptors must already be uploaded. */
int iam_upload_keypoint_keys(iam_ctx* ctx, int image_id, const int32_t* keys, int n);
/* Mirrors the LRU flush that sets des_list=None (matcher.py:1022-1026). */
int iam_release_descriptors(iam_ctx* ctx, int image_id);
/* Number of descriptors held for image_id, or <0 if absent. */
int iam_num_descriptors(iam_ctx* ctx, int image_id);
/* 1 if the image's descriptors were integer valued in [0,255] (exact path,
 * SURVEY D8), 0 if they were rounded to fp16 (tolerance path). */
int iam_descriptors_exact(iam_ctx* ctx, int image_id);

/* ---- kNN: replaces cv2 DescriptorMatcher.knnMatch -------------------- */

/* For every pair p = (pairs[2p], pairs[2p+1]) = (i, j) computes
 *   forward : for each descriptor of image i its k nearest in image j
 *   reverse : for each descriptor of image j its k nearest in image i
 * exactly as cv2.BFMatcher(norm).knnMatch(des_i, des_j, k) would
 * (ascending distance, ties -> lowest trainIdx).  Outputs are HOST buffers
 * laid out [P][n_stride][k]; rows >= n(image) are left untouched.  Distances
 * are float32: sqrt(sum of squared differences) for L2, the bit count for
 * Hamming.  `out_*_rev` may be NULL to skip the reverse direction.
 * Reference: matcher.py:203-216 (raw_matches), called with k=2 (:219,:600)
 * and k=3 (:465,:707). */
int iam_knn_pairs(iam_ctx* ctx, const int32_t* pairs, int n_pairs, int k,
                  int n_stride,
                  int32_t* out_idx_fwd, float* out_dist_fwd,
                  int32_t* out_idx_rev, float* out_dist_rev);

/* Keypoint pixel coordinates (cv2.KeyPoint.pt, float32 xy interleaved, HOST) of an image whose
 * descriptors are resident; consumed by the GMS filter.  They stay until replaced or released. */
int iam_upload_keypoints(iam_ctx* ctx, int image_id, const float* xy, int n);
/* The same for many images with one synchronisation at the end: xy[i] = HOST float [counts[i]][2] of image ids[i]. */
int iam_upload_keypoints_batch(iam_ctx* ctx, int n_images, const int32_t* ids, const float* const* xy, const int32_t* counts);

/* Stand-alone GMS grid filter on one HOST match list: what
 * cv2.xfeatures2d.matchGMS((w,h), (w,h), kp1, kp2, matches, withRotation, withScale, thresholdFactor)
 * returns at matcher.py:285, as a mask over `matches` ([n][2] = queryIdx, trainIdx; at most 4096).
 * Duplicate (queryIdx, trainIdx) rows are not supported.  with_scale: bit 0 = withScale; bit 1 selects the
 * last-half-cell rule of the archive Python restatement (see iam_match_params.gms = 2). */
int iam_gms_filter(iam_ctx* ctx, const float* xy1, int n1, const float* xy2, int n2,
                   const int32_t* matches, int n_matches, int width_px, int height_px,
                   int with_rotation, int with_scale, double threshold_factor, uint8_t* out_mask);

/* ---- full per-pair match: kNN (both ways) + reduction + cross-check --- */

typedef struct iam_match_params {
  double match_ratio;   /* /config/matcher/match_ratio (3a-matching.py:40), 0.75; a Python float, hence double */
  double max_distance;  /* matcher.max_distance: 270.0 L2 / 64 Hamming (matcher.py:53,56) */
  int    reduce_mode;   /* IAM_REDUCE_*                                            */
  int    cap;           /* `mymax` = 2000 (matcher.py:265)                         */
  int    min_pairs;     /* /config/matcher/min_pairs (matcher.py:80,271,312)       */
  int    cross_check;   /* !=0: filter_cross_check (matcher.py:187-200)            */
  int    dedupe;        /* !=0: filter_duplicates (matcher.py:157-182, :294) + its min_pairs gate (:296-298);
                           uses the keys given to iam_upload_keypoint_keys (identity keys if none)  */
  int    gms;           /* 1 or 2: GMS grid filter between the metric reduction and filter_duplicates, as
                           cv2.xfeatures2d.matchGMS(size, size, kp1, kp2, matches, withRotation, withScale,
                           thresholdFactor) at matcher.py:285; needs iam_upload_keypoints for both images.
                           1 = OpenCV's C++ rule for key points in the last half cell of the image (the match is
                           skipped for the half-cell-shifted grids); 2 = the rule of the reference's archive Python
                           restatement (scripts/lib/archive/gms_matcher.py:205: wrap-around to the last cell) */
  int    gms_rotation;  /* withRotation (matcher.py:285: True)                     */
  int    gms_scale;     /* withScale    (matcher.py:285: False)                    */
  double gms_threshold; /* thresholdFactor (matcher.py:285: 5.0)                   */
  int    width_px;      /* camera.get_image_params() (matcher.py:275-283)          */
  int    height_px;
} iam_match_params;

/* Device pipeline for the 'traditional' strategy
 * (bidirectional_pair_matches, matcher.py:304-318, with basic_pair_matches
 * :218-273 for each direction).  Writes per pair a table of [queryIdx,
 * trainIdx] rows in the order the reference would produce (ascending
 * metric, stable) into HOST buffers:
 *   out_table : [P][cap][2] int32,  out_count : [P] int32.
 * The reverse table is by construction the column swap of the forward one
 * after cross-check (matcher.py:195-196); without cross-check pass
 * out_table_rev/out_count_rev to receive the independent reverse table. */
int iam_match_pairs(iam_ctx* ctx, const int32_t* pairs, int n_pairs,
                    const iam_match_params* prm,
                    int32_t* out_table, int32_t* out_count,
                    int32_t* out_table_rev, int32_t* out_count_rev);

/* One-call form for a whole project step (what find_matches() does per call,
 * matcher.py:852-1031): descriptors are still in HOST memory.  Image
 * image_ids[i] has counts[i] rows of `dtype` at host_ptrs[i] (page-locked
 * memory makes the copies asynchronous); key_ptrs (may be NULL, entries may
 * be NULL) are the per-keypoint position ids of iam_upload_keypoint_keys.
 * The pair list is cut into waves; each image's H2D copy + layout conversion
 * is enqueued on an upload stream right before the first wave that needs it,
 * so PCIe transfers overlap the matching of earlier waves.  Outputs as
 * iam_match_pairs.  The images stay resident afterwards.
 * float32 L2 descriptors (the reference's SIFT arrays, image.py:160-180, are
 * integers in 0..255): worker threads of the library narrow them to bytes in
 * a page-locked arena while earlier waves upload and match, so a quarter of
 * the bytes crosses PCIe; an image with a component that is not an integer in
 * 0..255 is sent as float32 and the call repeats itself on fp16 operands.
 * Transport only: every distance is computed on the GPU.  Environment:
 * IAM_HOST_NARROW=0 switches it off, IAM_HOST_THREADS=n sets the worker count
 * (default: hardware threads / LOCAL_WORLD_SIZE, at most 16, minus one). */
int iam_match_images(iam_ctx* ctx, int n_images, const int32_t* image_ids,
                     const void* const* host_ptrs, const int32_t* counts, int dtype,
                     const int32_t* const* key_ptrs,
                     const int32_t* pairs, int n_pairs, const iam_match_params* prm,
                     int32_t* out_table, int32_t* out_count);

/* Same work, but results stay in device memory owned by the context (the
 * "inputs and outputs resident in HBM" form used for kernel-level timing and
 * by the multi-GPU gather).  Pointers returned are DEVICE pointers valid
 * until the next call on this context. */
int iam_match_pairs_device(iam_ctx* ctx, const int32_t* pairs, int n_pairs,
                           const iam_match_params* prm,
                           void** d_table, void** d_count);
/* Copy the device tables of the last iam_match_pairs_device() to the host. */
int iam_fetch_tables(iam_ctx* ctx, int32_t* out_table, int32_t* out_count);

/* Compact (CSR) form of the tables of the last iam_match_pairs_device() / iam_match_images() call, left in device
 * memory: d_offsets [n_pairs + 1] int32 = exclusive prefix sums of the counts, d_rows [total][2] int32 = the valid
 * [queryIdx, trainIdx] rows of every pair back to back, in pair order; *total = d_offsets[n_pairs] (the call
 * synchronises to return it).  This is what the pair-sharded multi-GPU job all-gathers instead of the padded
 * tables: the per-pair results the reference stores with `i1.match_list[i2.name] = ...` (matcher.py:979-980)
 * are lists of very different lengths.  Pointers stay valid until the next call on this context. */
int iam_pack_tables_device(iam_ctx* ctx, void** d_rows, void** d_offsets, long long* total);
/* The same compact form on the HOST: out_rows [cap_rows][2] receives the *total valid rows back to back, out_offsets
 * [n_pairs + 1] the prefix offsets (pair p owns rows offsets[p] .. offsets[p + 1]) -- the ragged `match_list` of
 * matcher.py:979-980 without the padding of iam_fetch_tables (a few per cent to ~40 % of the padded bytes).
 * IAM_E_UNSUPPORTED when cap_rows is too small (*total then holds the size needed). */
int iam_fetch_packed_tables(iam_ctx* ctx, int32_t* out_rows, long long cap_rows, int32_t* out_offsets, long long* total);

/* ---- RANSAC: replaces cv2.findEssentialMat(..., RANSAC, threshold) ---- */

/* Batched robust model fit for filter_by_transform (matcher.py:90-142;
 * the call being replaced is :126 / :122).  Pair p owns points
 * pts1[off[p] .. off[p+1]) / pts2[...] (float32 xy, pixel coordinates, HOST).
 * K is the row-major 3x3 camera matrix (double).  Outputs (HOST):
 * out_mask[off[P]] (1 = inlier), out_model[P][9] double (E, H or F, row-major),
 * out_inliers[P].  `seed` makes the sampler reproducible. */
int iam_ransac_pairs(iam_ctx* ctx, int model, const float* pts1, const float* pts2,
                     const int32_t* off, int n_pairs, const double* K,
                     double threshold_px, double prob, int max_iters, uint32_t seed,
                     uint8_t* out_mask, double* out_model, int32_t* out_inliers);

/* The same fit on the DEVICE tables of the last iam_match_pairs_device() / iam_match_images() call: the batched form
 * of find_matches' per-pair filter_by_transform(K, i1, i2, transform) (matcher.py:90-142).  The correspondences of
 * pair p are its table rows looked up in the resident key-point coordinates (iam_upload_keypoints; give it the images'
 * uv_list, matcher.py:112-113).  Pairs with fewer than min_pairs rows are emptied without a fit (matcher.py:99-101).
 * compact != 0 removes the outliers from the tables in place, order kept (matcher.py:134-141): iam_fetch_tables /
 * iam_pack_tables_device then return the filtered tables.  Outputs are HOST buffers and any may be NULL (with all NULL
 * the call only enqueues): out_mask [P][cap] (1 = inlier, rows beyond a pair's count untouched), out_model [P][9]
 * double, out_inliers [P]. */
int iam_ransac_tables(iam_ctx* ctx, int model, const double* K, double threshold_px, double prob, int max_iters,
                      uint32_t seed, int min_pairs, int compact, uint8_t* out_mask, double* out_model,
                      int32_t* out_inliers);

/* ---- pair-wise side estimators of the 'smart' strategy (scripts/lib/smart.py) ---- */

/* Two-view triangulation of the matches of n_pairs image pairs in one launch: what smart.triangulate_features
 * (smart.py:26-63) computes per pair with cv2.triangulatePoints(PROJ1, PROJ2, pts1, pts2) followed by `points /=
 * points[3]`, and the statistics estimate_surface_elevation (smart.py:116-131) takes of it (np.average / np.std of the
 * down coordinate).  proj1, proj2: HOST double [n_pairs][12], the row-major 3 x 4 matrices [R | t] of Image.get_proj;
 * off: HOST int32 [n_pairs + 1] prefix offsets into the point lists; x1, x2: HOST double [off[n_pairs]][2] normalised
 * image coordinates K^-1 (u, v, 1) (rows 0-1 of smart.py:58-59's pts1 / pts2, transposed).  Outputs (HOST): out_points
 * [off[n_pairs]][3] = X, Y, Z after the division by the homogeneous coordinate (may be NULL); out_stats [n_pairs][2] =
 * mean and population standard deviation of Z (NaN for a pair without points).  float64 throughout; agreement with
 * cv2: 1e-9 relative on well-conditioned geometry (tests/test_smart.py).  The partial-affine fit of find_affine
 * (smart.py:66-90, cv2.estimateAffinePartial2D) is iam_ransac_pairs with IAM_MODEL_AFFINE_PARTIAL. */
int iam_triangulate_pairs(iam_ctx* ctx, int n_pairs, const double* proj1, const double* proj2, const int32_t* off,
                          const double* x1, const double* x2, double* out_points, double* out_stats);

/* ---- feature detection: replaces cv2.ORB_create(n).detectAndCompute ---- */

/* ORB detect + describe with OpenCV's defaults (8 levels, scale 1.2, edge threshold 31, patch 31, FAST threshold
 * 20, Harris score, WTA_K 2): what `detector = cv2.ORB_create(max_features)` / `detector.detectAndCompute(scaled,
 * None)` return in Image.detect_features (image.py:243-245, :324).  gray: HOST uint8 [height][width], row-major (a
 * colour image must be converted first, as cv2 does internally).  Outputs (HOST): out_kp [max_out][6] float =
 * pt.x, pt.y, size, angle (degrees), response, octave; out_des [max_out][32] uint8; *out_n = number of key points
 * (listed level by level, raster order inside a level; cv2's order inside a level is unspecified).  max_out should
 * leave room for ties (retainBest keeps every key point that ties with the last one): nfeatures + 256 is ample;
 * IAM_E_UNSUPPORTED if it does not fit. */
int iam_orb_detect(iam_ctx* ctx, const uint8_t* gray, int width, int height, int nfeatures, int max_out,
                   float* out_kp, uint8_t* out_des, int* out_n);

/* SIFT detect + describe with OpenCV's defaults (3 layers per octave, sigma 1.6, contrast threshold 0.04, edge
 * threshold 10, doubled base image): what `detector = cv2.SIFT_create()` / `detector.detectAndCompute(scaled, None)`
 * return in Image.detect_features (image.py:236-237, :324) — the pipeline's default detector.  gray: HOST uint8
 * [height][width].  Outputs (HOST): out_kp [max_out][5] float = pt.x, pt.y, size, angle (degrees), response;
 * out_octave [max_out] = cv2's packed octave field; out_des [max_out][128] uint8 (cv2 hands the same integers out as
 * float32; the matcher consumes uint8 directly); *out_n = number of key points, in cv2's order.  Float arithmetic:
 * key points agree with cv2 to >= 99 % (position 1e-2 px), descriptors to +-1 per byte on those (tests/test_sift.py);
 * the remainder are extrema on a decision boundary of the pyramid's float rounding.  IAM_E_UNSUPPORTED if more than
 * max_out key points are found. */
int iam_sift_detect(iam_ctx* ctx, const uint8_t* gray, int width, int height, int max_out, float* out_kp,
                    int32_t* out_octave, uint8_t* out_des, int* out_n);

/* Debug aid: the FAST-9/16 corner score of every pixel of a grey image (threshold 20, no non-maximum suppression),
 * HOST in, HOST out [height][width]; what cv2.FastFeatureDetector reports as `response` at its key points. */
int iam_debug_orb_fast(iam_ctx* ctx, const uint8_t* gray, int width, int height, uint8_t* out_score);

/* ---- bundle adjustment: replaces Optimizer.fun (optimizer.py:174-279) -- */

/* The reprojection residual of every observation in one launch, and its
 * analytic Jacobian, for cam_method 'ned_quat' (optimizer.py:84-85): a camera
 * is [ned(3), quat(4) = (w, x, y, z), normalised as quaternion_matrix does],
 * the projection is cv2.projectPoints' pinhole + (k1, k2, p1, p2, k3) model.
 *
 * iam_ba_setup makes the problem STRUCTURE resident: observation i is pixel
 * obs_uv[i] = (u, v) of 3-D point pt_idx[i] seen by camera cam_idx[i];
 * observations are in the order Optimizer.setup() lays them out (by camera,
 * then list order, optimizer.py:396-404).  HOST pointers. */
int iam_ba_setup(iam_ctx* ctx, int n_cam, int n_pts, int n_obs, const int32_t* cam_idx,
                 const int32_t* pt_idx, const double* obs_uv);
/* One evaluation.  params (HOST) = [n_cam][7] cameras then [n_pts][3] points, the layout of x0
 * (optimizer.py:422-423); K4 = (fx, fy, cx, cy), dist5 = distCoeffs (either the fixed calibration, :190-191,
 * or the tail of the parameter vector in global-calibration mode, :182-188).
 * out_residual[2*n_obs] = observed - projected, (u, v) interleaved: the vector fun() returns.
 * out_jac (may be NULL) [n_obs][2][10]: per residual row d/d(ned, quat) of its camera, then d/d(point) --
 * the non-zeros of the row in the column order of bundle_adjustment_sparsity (:142-169). */
int iam_ba_eval(iam_ctx* ctx, const double* params, const double* K4, const double* dist5,
                double* out_residual, double* out_jac);
/* The same in two steps with the results left in device memory (pointers valid until the next iam_ba_* call). */
int iam_ba_upload_params(iam_ctx* ctx, const double* params);
int iam_ba_eval_device(iam_ctx* ctx, const double* K4, const double* dist5, int want_jac,
                       void** d_residual, void** d_jac);
/* Global-calibration mode (optimizer.py:146-147, :160-166, :181-189: K and distCoeffs are the last eight entries of the
 * parameter vector, fx = fy): the eight dense Jacobian columns of every residual row, at the parameters last uploaded
 * (iam_ba_eval / iam_ba_upload_params).  out_jac_calib (HOST) [n_obs][2][8] = d residual / d (f, cu, cv, k1, k2, p1,
 * p2, k3), the column order of bundle_adjustment_sparsity's calibration block. */
int iam_ba_calib_jacobian(iam_ctx* ctx, const double* K4, const double* dist5, double* out_jac_calib);
/* Debug aid: the per-observation function of the kernel (same source) evaluated on the HOST.  Needs no GPU. */
int iam_debug_ba_host(const double* cam7, const double* pt3, const double* uv, const double* K4,
                      const double* dist5, double* out_res2, double* out_jac20);

/* ---- instrumentation ------------------------------------------------- */

/* When enabled, the context brackets its kernels with CUDA events on the
 * launching stream. */
int iam_set_profiling(iam_ctx* ctx, int enable);
typedef struct iam_timing {
  float knn_ms;        /* distance+top-k kernel(s) of the last call          */
  float reduce_ms;     /* ratio/metric reduction + cross-check kernels       */
  float convert_ms;    /* descriptor layout conversion (last upload)         */
  int   knn_launches;  /* kernels launched by the last match/knn call        */
  int   total_launches;/* all kernel launches since context creation         */
  int   engine_used;   /* IAM_ENGINE_UMMA or IAM_ENGINE_SIMT                 */
  int   waves;         /* pair-list chunks of the last match call            */
  int   mma_kind;      /* IAM_KIND_* of the last tensor-core launch, -1: SIMT */
  /* last iam_match_images call (device timeline from CUDA events, host time from a monotonic clock) */
  float host_enqueue_ms;  /* host time spent enqueueing uploads + kernels          */
  float upload_span_ms;   /* first H2D copy start -> last conversion end           */
  float compute_span_ms;  /* first kernel start -> last kernel end                 */
  float total_span_ms;    /* first H2D copy start -> last kernel end               */
  unsigned long long h2d_bytes; /* descriptor bytes copied host -> device (float32 images that worker threads
                                   narrowed to bytes on the host count as bytes)       */
  int   narrowed_images;  /* images uploaded as host-narrowed bytes                */
} iam_timing;
int iam_get_timing(iam_ctx* ctx, iam_timing* out);

/* Debug aid (no reference counterpart): raw fp32 accumulators of one 128x128
 * tensor-core distance tile (query tile `q_tile` of image q_id x train tile
 * `t_tile` of image t_id) with explicit UMMA descriptor strides, into
 * out_host[128*128].  Production values: lbo=128, sbo=2304, kstep_bytes=256,
 * ksteps=9 (negative: A operand through tensor memory).  lbo=0 selects the
 * byte layout (kind::i8) of an L2 context with its own strides; the tile is
 * then s32 accumulators (converted to float) of RANK-ordered rows. */
int iam_debug_tile(iam_ctx* ctx, int q_id, int t_id, int q_tile, int t_tile,
                   uint32_t lbo, uint32_t sbo, uint32_t kstep_bytes, int ksteps,
                   float* out_host);

/* Debug aid: run the minimal solver of iam_ransac_pairs (5-point essential /
 * 4-point homography; same source as the device code) on the HOST for the
 * first 5 / 4 normalised correspondences.  Returns the number of models
 * written to out_models[10][9] (row-major), or <0.  Needs no GPU. */
int iam_debug_minimal_solver(int model, const float* x1, const float* y1,
                             const float* x2, const float* y2, float* out_models);

/* Debug aid: the host-side float32 -> uint8 narrowing iam_match_images applies to integer-valued L2 descriptors
 * before they cross PCIe (transport only; the reference holds SIFT descriptors as float32, image.py:160-180).
 * Returns 0 when every src[i] is an integer in 0..255 (dst[i] = that value), 1 otherwise.  Needs no GPU. */
int iam_debug_narrow(const float* src, uint8_t* dst, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* IAMATCH_H */
