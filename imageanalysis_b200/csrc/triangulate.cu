// triangulate.cu — batched two-view triangulation for the pair-wise surface estimate of the 'smart' strategy
// (reference scripts/lib/smart.py:26-63 triangulate_features -> cv2.triangulatePoints, :116-131
// estimate_surface_elevation: -mean and std of the triangulated down coordinates).
//
// cv2.triangulatePoints solves, per correspondence, the homogeneous 4 x 4 system
//     [ x1 P1[2] - P1[0] ;  y1 P1[2] - P1[1] ;  x2 P2[2] - P2[0] ;  y2 P2[2] - P2[1] ] X = 0
// by SVD and returns the right singular vector of the smallest singular value.  Here: one thread per correspondence,
// one-sided (Hestenes) Jacobi rotations on the columns of that matrix in float64 -- the rotations are accumulated in V,
// the column that ends with the smallest norm is the answer.  One block per image pair; the block then reduces the mean
// and the population standard deviation of Z / W in two passes (numpy's np.average / np.std), fixed reduction order.
#include <cuda_runtime.h>

#include <cstdint>

#include "triangulate.h"

namespace iam {
namespace {

constexpr int kThreads = 128;

__device__ void smallest_right_singular_vector(double A[4][4], double X[4]) {
  double V[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 40; ++sweep) {
    bool changed = false;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          alpha += A[i][p] * A[i][p];
          beta += A[i][q] * A[i][q];
          gamma += A[i][p] * A[i][q];
        }
        if (fabs(gamma) <= 1e-16 * sqrt(alpha * beta) || gamma == 0.0) continue;
        changed = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double ap = A[i][p], aq = A[i][q];
          A[i][p] = c * ap - s * aq;
          A[i][q] = s * ap + c * aq;
          const double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - s * vq;
          V[i][q] = s * vp + c * vq;
        }
      }
    if (!changed) break;
  }
  double best = -1.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double nrm = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) nrm += A[i][j] * A[i][j];
    if (best < 0 || nrm < best) {
      best = nrm;
#pragma unroll
      for (int i = 0; i < 4; ++i) X[i] = V[i][j];
    }
  }
}

__device__ double block_sum(double v, double* s_red) {
  const int t = threadIdx.x;
  s_red[t] = v;
  __syncthreads();
  for (int d = kThreads / 2; d > 0; d >>= 1) {
    if (t < d) s_red[t] += s_red[t + d];
    __syncthreads();
  }
  const double r = s_red[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kThreads) triangulate_kernel(const double* __restrict__ proj1, const double* __restrict__ proj2,
                                                               const int32_t* __restrict__ off, const double* __restrict__ x1,
                                                               const double* __restrict__ x2, double* __restrict__ out_points,
                                                               double* __restrict__ out_stats) {
  __shared__ double s_p[24];
  __shared__ double s_red[kThreads];
  const int p = blockIdx.x, t = threadIdx.x;
  if (t < 12) s_p[t] = proj1[(size_t)p * 12 + t];
  else if (t < 24) s_p[t] = proj2[(size_t)p * 12 + t - 12];
  __syncthreads();
  const int o = off[p], n = off[p + 1] - o;
  double zsum = 0;
  for (int i = t; i < n; i += kThreads) {
    const double xa = x1[2 * (size_t)(o + i)], ya = x1[2 * (size_t)(o + i) + 1];
    const double xb = x2[2 * (size_t)(o + i)], yb = x2[2 * (size_t)(o + i) + 1];
    double A[4][4], X[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      A[0][k] = xa * s_p[8 + k] - s_p[k];
      A[1][k] = ya * s_p[8 + k] - s_p[4 + k];
      A[2][k] = xb * s_p[12 + 8 + k] - s_p[12 + k];
      A[3][k] = yb * s_p[12 + 8 + k] - s_p[12 + 4 + k];
    }
    smallest_right_singular_vector(A, X);
    const double z = X[2] / X[3];
    out_points[3 * (size_t)(o + i)] = X[0] / X[3];
    out_points[3 * (size_t)(o + i) + 1] = X[1] / X[3];
    out_points[3 * (size_t)(o + i) + 2] = z;
    zsum += z;
  }
  const double mean = block_sum(zsum, s_red) / n;
  double dev = 0;
  for (int i = t; i < n; i += kThreads) {   // each thread re-reads what it wrote
    const double d = out_points[3 * (size_t)(o + i) + 2] - mean;
    dev += d * d;
  }
  const double var = block_sum(dev, s_red) / n;
  if (t == 0) {
    out_stats[2 * p] = mean;          // n == 0: 0 / 0 = NaN, as np.average of an empty array
    out_stats[2 * p + 1] = sqrt(var);
  }
}

}  // namespace

cudaError_t triangulate_pairs(int n_pairs, const double* proj1, const double* proj2, const int32_t* off, const double* x1,
                              const double* x2, double* out_points, double* out_stats, cudaStream_t stream) {
  if (n_pairs <= 0) return cudaSuccess;
  triangulate_kernel<<<n_pairs, kThreads, 0, stream>>>(proj1, proj2, off, x1, x2, out_points, out_stats);
  return cudaGetLastError();
}

}  // namespace iam
