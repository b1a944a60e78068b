// ba.h — reprojection residual + analytic Jacobian of the bundle-adjustment cost
// (reference scripts/lib/optimizer.py:174-279, Optimizer.fun, cam_method 'ned_quat').
#pragma once
#include <cuda_runtime.h>

namespace iam {

struct BaCalib {
  double fx, fy, cx, cy;        // K (optimizer.py:190-191 / :182-188 in global-calibration mode)
  double k1, k2, p1, p2, k3;    // distCoeffs in OpenCV order
};

constexpr int kBaJacCols = 10;  // per residual row: d/d ned (3), d/d quat (4), d/d point (3)

// One observation: residual (observed - projected) and, when `jac` is non-null, its 2 x 10 Jacobian block.
// Shared by the device kernel and the host-side debug entry point.
__host__ __device__ void ba_observation(const double* cam7, const double* pt3, double u_obs, double v_obs,
                                        const BaCalib& c, double* res2, double* jac20);

// residual[2*i..] and jac[20*i..] for observation i = 0..n_obs-1 (all device pointers; jac may be null).
cudaError_t launch_ba(const double* cams, const double* points, const int* cam_idx, const int* pt_idx,
                      const double* obs_uv, int n_obs, const BaCalib& calib, double* residual, double* jac,
                      cudaStream_t stream);

// Global-calibration mode: jac_calib[16*i..] = the 2 x 8 block d residual_i / d (f, cu, cv, k1, k2, p1, p2, k3), fx = fy = f.
cudaError_t launch_ba_calib(const double* cams, const double* points, const int* cam_idx, const int* pt_idx, int n_obs,
                            const BaCalib& calib, double* jac_calib, cudaStream_t stream);

}  // namespace iam
