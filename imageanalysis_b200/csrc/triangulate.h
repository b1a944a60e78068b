// triangulate.h — two-view triangulation of matched features (triangulate.cu), replacing
// cv2.triangulatePoints(PROJ1, PROJ2, pts1, pts2) + the mean / standard deviation of the heights in
// smart.triangulate_features / estimate_surface_elevation (reference scripts/lib/smart.py:26-63, :116-131).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

namespace iam {

// All pointers are DEVICE pointers.  proj1 / proj2: [n_pairs][12] row-major 3 x 4 projection matrices [R | t];
// off: [n_pairs + 1] prefix offsets into the point lists; x1 / x2: [total][2] normalised image coordinates
// (K^-1 (u, v, 1)); out_points: [total][3] = X / W, Y / W, Z / W of the homogeneous solution; out_stats: [n_pairs][2] =
// mean and population standard deviation of Z / W over the pair's points (NaN for an empty pair).
cudaError_t triangulate_pairs(int n_pairs, const double* proj1, const double* proj2, const int32_t* off, const double* x1,
                              const double* x2, double* out_points, double* out_stats, cudaStream_t stream);

}  // namespace iam
