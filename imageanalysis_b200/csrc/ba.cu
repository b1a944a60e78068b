// ba.cu — reprojection residual and analytic Jacobian for every observation of a
// bundle-adjustment problem in one launch, replacing the per-camera
// cv2.projectPoints loop of Optimizer.fun (reference scripts/lib/optimizer.py:198-228)
// and the finite differences SciPy takes of it (least_squares(..., jac_sparsity=A), :491-501).
//
// One thread per observation, float64 throughout like the reference.  HBM-bound:
// per observation ~32 B of streamed input (indices, observed pixel) plus gathered camera / point
// parameters (cache hits: observations are grouped by camera), 16 B of residual and 160 B of
// Jacobian out.  The 2 x 10 blocks are staged through shared memory so a warp writes its
// 5120 contiguous bytes with full 128-byte lines.
#include <cuda_runtime.h>

#include "ba.h"

namespace iam {

__host__ __device__ void ba_observation(const double* cam, const double* X, double u_obs, double v_obs,
                                        const BaCalib& c, double* res, double* jac) {
  // body -> ned rotation of the (unnormalised) quaternion, as transformations.quaternion_matrix normalises it
  // (optimizer.py:121-127): b2n = M / n, M quadratic in q
  const double w = cam[3], x = cam[4], y = cam[5], z = cam[6];
  const double n = w * w + x * x + y * y + z * z;
  const double inv_n = 1.0 / n;
  const double M[3][3] = {{w * w + x * x - y * y - z * z, 2.0 * (x * y - z * w), 2.0 * (x * z + y * w)},
                          {2.0 * (x * y + z * w), w * w - x * x + y * y - z * z, 2.0 * (y * z - x * w)},
                          {2.0 * (x * z - y * w), 2.0 * (y * z + x * w), w * w - x * x - y * y + z * z}};
  const double d[3] = {X[0] - cam[0], X[1] - cam[1], X[2] - cam[2]};
  // body frame: Xb = b2n^T d ; camera frame: Xc = body2cam Xb = (Xb.y, Xb.z, Xb.x)   (cam2body :91-94)
  double Xb[3];
  for (int i = 0; i < 3; ++i) Xb[i] = (M[0][i] * d[0] + M[1][i] * d[1] + M[2][i] * d[2]) * inv_n;
  const double Xc[3] = {Xb[1], Xb[2], Xb[0]};
  // cv2.projectPoints: pinhole + (k1, k2, p1, p2, k3)
  const double iz = 1.0 / Xc[2];
  const double xn = Xc[0] * iz, yn = Xc[1] * iz;
  const double r2 = xn * xn + yn * yn;
  const double rad = 1.0 + r2 * (c.k1 + r2 * (c.k2 + r2 * c.k3));
  const double xd = xn * rad + 2.0 * c.p1 * xn * yn + c.p2 * (r2 + 2.0 * xn * xn);
  const double yd = yn * rad + c.p1 * (r2 + 2.0 * yn * yn) + 2.0 * c.p2 * xn * yn;
  res[0] = u_obs - (c.fx * xd + c.cx);
  res[1] = v_obs - (c.fy * yd + c.cy);
  if (!jac) return;

  // d(u, v) / d Xc
  const double drad = c.k1 + r2 * (2.0 * c.k2 + 3.0 * c.k3 * r2);
  const double xd_x = rad + 2.0 * xn * xn * drad + 2.0 * c.p1 * yn + 6.0 * c.p2 * xn;
  const double xd_y = 2.0 * xn * yn * drad + 2.0 * c.p1 * xn + 2.0 * c.p2 * yn;
  const double yd_x = xd_y;
  const double yd_y = rad + 2.0 * yn * yn * drad + 6.0 * c.p1 * yn + 2.0 * c.p2 * xn;
  double A[2][3];  // rows u, v; columns Xc
  A[0][0] = c.fx * xd_x * iz;
  A[0][1] = c.fx * xd_y * iz;
  A[0][2] = -c.fx * (xd_x * xn + xd_y * yn) * iz;
  A[1][0] = c.fy * yd_x * iz;
  A[1][1] = c.fy * yd_y * iz;
  A[1][2] = -c.fy * (yd_x * xn + yd_y * yn) * iz;
  // in body coordinates: Xc = (Xb1, Xb2, Xb0)  ->  Ab[.][0] = A[.][2], Ab[.][1] = A[.][0], Ab[.][2] = A[.][1]
  double Ab[2][3];
  for (int r = 0; r < 2; ++r) {
    Ab[r][0] = A[r][2];
    Ab[r][1] = A[r][0];
    Ab[r][2] = A[r][1];
  }
  // d Xb / d X = b2n^T, d Xb / d ned = -b2n^T
  for (int r = 0; r < 2; ++r)
    for (int j = 0; j < 3; ++j) {
      const double g = (Ab[r][0] * M[j][0] + Ab[r][1] * M[j][1] + Ab[r][2] * M[j][2]) * inv_n;  // d proj / d X_j
      jac[r * kBaJacCols + 7 + j] = -g;   // residual = observed - projected
      jac[r * kBaJacCols + j] = g;        // ned enters as X - ned
    }
  // d Xb / d q_k = ((dM/dq_k)^T d - 2 q_k Xb) / n
  const double dM[4][3][3] = {
      {{w, -z, y}, {z, w, -x}, {-y, x, w}},     // dM/dw / 2
      {{x, y, z}, {y, -x, -w}, {z, w, -x}},     // dM/dx / 2
      {{-y, x, w}, {x, y, z}, {-w, z, -y}},     // dM/dy / 2
      {{-z, -w, x}, {w, -z, y}, {x, y, z}}};    // dM/dz / 2
  const double qk[4] = {w, x, y, z};
  for (int k = 0; k < 4; ++k) {
    double t[3];
    for (int i = 0; i < 3; ++i)
      t[i] = 2.0 * ((dM[k][0][i] * d[0] + dM[k][1][i] * d[1] + dM[k][2][i] * d[2]) - qk[k] * Xb[i]) * inv_n;
    for (int r = 0; r < 2; ++r) jac[r * kBaJacCols + 3 + k] = -(Ab[r][0] * t[0] + Ab[r][1] * t[1] + Ab[r][2] * t[2]);
  }
}

namespace {

constexpr int kBaThreads = 256;
constexpr int kBaPad = 2 * kBaJacCols + 1;  // 21 doubles per observation in shared memory: no 8-way bank conflicts

template <bool kJac>
__global__ void __launch_bounds__(kBaThreads)
ba_kernel(const double* __restrict__ cams, const double* __restrict__ points, const int* __restrict__ cam_idx,
          const int* __restrict__ pt_idx, const double2* __restrict__ obs_uv, int n_obs, BaCalib calib,
          double2* __restrict__ residual, double* __restrict__ jac) {
  __shared__ double s_jac[kJac ? kBaThreads * kBaPad : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = blockIdx.x * kBaThreads; base < n_obs; base += gridDim.x * kBaThreads) {
    const int i = base + threadIdx.x;
    double j20[2 * kBaJacCols];
    if (i < n_obs) {
      const int ci = cam_idx[i], pi = pt_idx[i];
      double cam[7], X[3];
#pragma unroll
      for (int k = 0; k < 7; ++k) cam[k] = __ldg(cams + static_cast<size_t>(ci) * 7 + k);
#pragma unroll
      for (int k = 0; k < 3; ++k) X[k] = __ldg(points + static_cast<size_t>(pi) * 3 + k);
      const double2 uv = obs_uv[i];
      double res[2];
      ba_observation(cam, X, uv.x, uv.y, calib, res, kJac ? j20 : nullptr);
      residual[i] = make_double2(res[0], res[1]);
    }
    if (kJac) {
      // stage the warp's 32 x 20 block, then write it out as 20 fully coalesced rows of 32 doubles
      double* mine = s_jac + static_cast<size_t>(threadIdx.x) * kBaPad;
      if (i < n_obs) {
#pragma unroll
        for (int k = 0; k < 2 * kBaJacCols; ++k) mine[k] = j20[k];
      }
      __syncwarp();
      const int wbase = base + warp * 32;                       // first observation of this warp
      const int wcount = min(32, n_obs - wbase);                // observations it really has
      const double* ws = s_jac + static_cast<size_t>(warp) * 32 * kBaPad;
      for (int e = lane; e < wcount * 2 * kBaJacCols; e += 32) {
        const int o = e / (2 * kBaJacCols), k = e - o * (2 * kBaJacCols);
        jac[static_cast<size_t>(wbase) * 2 * kBaJacCols + e] = ws[o * kBaPad + k];
      }
      __syncwarp();
    }
  }
}

// Global-calibration mode (optimizer.py:146-147, :160-166, :181-189): the eight dense columns every residual row also
// has -- d/d(f, cu, cv, k1, k2, p1, p2, k3) with fx = fy = f.  One thread per observation; the 2 x 8 block is staged
// through shared memory like the main Jacobian so that a warp writes 4096 contiguous bytes.
constexpr int kCalCols = 8, kCalPad = 2 * kCalCols + 1;
__global__ void __launch_bounds__(kBaThreads)
ba_calib_kernel(const double* __restrict__ cams, const double* __restrict__ points, const int* __restrict__ cam_idx,
                const int* __restrict__ pt_idx, int n_obs, BaCalib c, double* __restrict__ jac) {
  __shared__ double s_jac[kBaThreads * kCalPad];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = blockIdx.x * kBaThreads; base < n_obs; base += gridDim.x * kBaThreads) {
    const int i = base + threadIdx.x;
    double* mine = s_jac + static_cast<size_t>(threadIdx.x) * kCalPad;
    if (i < n_obs) {
      const double* cam = cams + static_cast<size_t>(cam_idx[i]) * 7;
      const double* X = points + static_cast<size_t>(pt_idx[i]) * 3;
      const double w = cam[3], x = cam[4], y = cam[5], z = cam[6];
      const double inv_n = 1.0 / (w * w + x * x + y * y + z * z);
      const double M[3][3] = {{w * w + x * x - y * y - z * z, 2.0 * (x * y - z * w), 2.0 * (x * z + y * w)},
                              {2.0 * (x * y + z * w), w * w - x * x + y * y - z * z, 2.0 * (y * z - x * w)},
                              {2.0 * (x * z - y * w), 2.0 * (y * z + x * w), w * w - x * x - y * y + z * z}};
      const double d[3] = {X[0] - cam[0], X[1] - cam[1], X[2] - cam[2]};
      double Xb[3];
      for (int k = 0; k < 3; ++k) Xb[k] = (M[0][k] * d[0] + M[1][k] * d[1] + M[2][k] * d[2]) * inv_n;
      const double iz = 1.0 / Xb[0];
      const double xn = Xb[1] * iz, yn = Xb[2] * iz;
      const double r2 = xn * xn + yn * yn;
      const double rad = 1.0 + r2 * (c.k1 + r2 * (c.k2 + r2 * c.k3));
      const double xd = xn * rad + 2.0 * c.p1 * xn * yn + c.p2 * (r2 + 2.0 * xn * xn);
      const double yd = yn * rad + c.p1 * (r2 + 2.0 * yn * yn) + 2.0 * c.p2 * xn * yn;
      // residual = observed - (f * distorted + centre): every derivative carries a minus sign
      const double du[kCalCols] = {-xd, -1.0, 0.0, -c.fx * xn * r2, -c.fx * xn * r2 * r2, -c.fx * 2.0 * xn * yn,
                                   -c.fx * (r2 + 2.0 * xn * xn), -c.fx * xn * r2 * r2 * r2};
      const double dv[kCalCols] = {-yd, 0.0, -1.0, -c.fy * yn * r2, -c.fy * yn * r2 * r2, -c.fy * (r2 + 2.0 * yn * yn),
                                   -c.fy * 2.0 * xn * yn, -c.fy * yn * r2 * r2 * r2};
#pragma unroll
      for (int k = 0; k < kCalCols; ++k) {
        mine[k] = du[k];
        mine[kCalCols + k] = dv[k];
      }
    }
    __syncwarp();
    const int wbase = base + warp * 32;
    const int wcount = min(32, n_obs - wbase);
    const double* ws = s_jac + static_cast<size_t>(warp) * 32 * kCalPad;
    for (int e = lane; e < wcount * 2 * kCalCols; e += 32) {
      const int o = e / (2 * kCalCols), k = e - o * (2 * kCalCols);
      jac[static_cast<size_t>(wbase) * 2 * kCalCols + e] = ws[o * kCalPad + k];
    }
    __syncwarp();
  }
}

}  // namespace

cudaError_t launch_ba_calib(const double* cams, const double* points, const int* cam_idx, const int* pt_idx, int n_obs,
                            const BaCalib& calib, double* jac_calib, cudaStream_t stream) {
  if (n_obs <= 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int want = (n_obs + kBaThreads - 1) / kBaThreads;
  ba_calib_kernel<<<want < sms * 4 ? want : sms * 4, kBaThreads, 0, stream>>>(cams, points, cam_idx, pt_idx, n_obs, calib, jac_calib);
  return cudaGetLastError();
}

cudaError_t launch_ba(const double* cams, const double* points, const int* cam_idx, const int* pt_idx,
                      const double* obs_uv, int n_obs, const BaCalib& calib, double* residual, double* jac,
                      cudaStream_t stream) {
  if (n_obs <= 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int want = (n_obs + kBaThreads - 1) / kBaThreads;
  const int grid = want < sms * 4 ? want : sms * 4;   // grid-stride: a whole number of waves of the SM count
  if (jac)
    ba_kernel<true><<<grid, kBaThreads, 0, stream>>>(cams, points, cam_idx, pt_idx, reinterpret_cast<const double2*>(obs_uv),
                                                     n_obs, calib, reinterpret_cast<double2*>(residual), jac);
  else
    ba_kernel<false><<<grid, kBaThreads, 0, stream>>>(cams, points, cam_idx, pt_idx, reinterpret_cast<const double2*>(obs_uv),
                                                      n_obs, calib, reinterpret_cast<double2*>(residual), nullptr);
  return cudaGetLastError();
}

}  // namespace iam
