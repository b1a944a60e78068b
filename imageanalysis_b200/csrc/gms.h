// gms.h — launcher declaration of the GMS grid filter (gms.cu).
#pragma once
#include <cuda_runtime.h>

namespace iam {

struct RedJob;
struct ImgDev;

constexpr int kGmsMaxMatches = 4096;  // per directed job (the reference clips to 2000, matcher.py:265)

struct GmsParams {
  double threshold_factor;  // thresholdFactor of cv2.xfeatures2d.matchGMS (matcher.py:285: 5.0)
  int width, height;        // image size in pixels, both images (camera.get_image_params(), matcher.py:275-283)
  int with_rotation;        // matcher.py:285: True
  int with_scale;           // matcher.py:285: False
  int gate_min_pairs;       // > 0: fewer survivors than this empty the table (used when no filter_duplicates gate follows)
  int archive_wrap;         // != 0: a key point in the last half cell is looked up in cell 399 when inliers are marked (the
                            // wrap-around of the reference's archive Python, gms_matcher.py:205); 0: skipped (OpenCV's C++)
};

// In place on the per-job tables (order preserved).  Needs ImgDev::kp_xy of both images of every non-empty job.
cudaError_t launch_gms(const RedJob* jobs, int n_jobs, const ImgDev* imgs, const GmsParams& prm, int cap,
                       int* job_table, int* job_count, cudaStream_t stream);

}  // namespace iam
