// ransac.h — batched robust two-view model fitting (essential matrix via the
// 5-point minimal solver, homography via 4-point DLT), replacing the
// cv2.findEssentialMat / cv2.findHomography calls of filter_by_transform
// (reference scripts/lib/matcher.py:121-126).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

namespace iam {

// Host pointers in, host pointers out; work is enqueued on `stream` and the
// call returns after the results have been copied back.  Returns 0 or a
// negative IAM_E_* code with `err` filled in.
int ransac_pairs(int model, const float* pts1, const float* pts2, const int32_t* off, int n_pairs, const double* K,
                 double threshold_px, double prob, int max_iters, uint32_t seed, uint8_t* out_mask, double* out_model,
                 int32_t* out_inliers, cudaStream_t stream, std::string* err);

// Host-side run of the minimal solvers (same source as the device code) on the
// first 5 (essential) / 4 (homography) normalised correspondences; returns the
// number of models written to out[10][9].  Used by the CPU tests.
int debug_minimal_solver(int model, const float* x1, const float* y1, const float* x2, const float* y2, float* out);

}  // namespace iam
