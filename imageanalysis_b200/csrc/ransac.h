// ransac.h — batched robust two-view model fitting (essential matrix via the
// 5-point minimal solver, homography via 4-point DLT, fundamental matrix via the
// 7-point solver), replacing the cv2.findEssentialMat / cv2.findHomography /
// cv2.findFundamentalMat calls of filter_by_transform (reference
// scripts/lib/matcher.py:121-126) and the per-bin homography fits of the
// strategies (:532, :637, :803).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

namespace iam {

struct RedJob;
struct ImgDev;

// Device block kept between calls (owned by the context): no allocation on the hot path.
struct RansacScratch {
  void* buf = nullptr;
  size_t cap = 0;
  RansacScratch() = default;
  RansacScratch(const RansacScratch&) = delete;
  RansacScratch& operator=(const RansacScratch&) = delete;
  ~RansacScratch();
};

// Host pointers in, host pointers out; work is enqueued on `stream` and the
// call returns after the results have been copied back.  Returns 0 or a
// negative IAM_E_* code with `err` filled in.
int ransac_pairs(int model, const float* pts1, const float* pts2, const int32_t* off, int n_pairs, const double* K,
                 double threshold_px, double prob, int max_iters, uint32_t seed, uint8_t* out_mask, double* out_model,
                 int32_t* out_inliers, RansacScratch* scratch, cudaStream_t stream, std::string* err);

// Everything on the device: the correspondences of pair p are rows [0, d_count[p]) of d_table[p] ([cap][2] =
// queryIdx, trainIdx) looked up in the key points (ImgDev::kp_xy) of the pair's images (d_jobs[2p].q_slot / t_slot).
// compact: drop the outliers from the tables in place, order preserved, and rewrite d_count
// (what filter_by_transform does to match_list, matcher.py:134-141); pairs with fewer than min_pairs rows are
// emptied without a fit (:99-101).  d_mask ([P][cap]) may be null.  Enqueue only.
int ransac_tables(int model, int* d_table, int* d_count, int cap, int n_pairs, const RedJob* d_jobs, const ImgDev* d_imgs,
                  const double* K, double threshold_px, double prob, int max_iters, uint32_t seed, int min_pairs,
                  bool compact, uint8_t* d_mask, float* d_model, int* d_inliers, cudaStream_t stream, std::string* err);

// Host-side run of the minimal solvers (same source as the device code) on the
// first 5 (essential) / 4 (homography) / 7 (fundamental) normalised correspondences; returns the
// number of models written to out[10][9].  Used by the CPU tests.
int debug_minimal_solver(int model, const float* x1, const float* y1, const float* x2, const float* y2, float* out);

}  // namespace iam
