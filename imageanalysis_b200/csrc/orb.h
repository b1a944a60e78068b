// orb.h — ORB detect + describe (orb.cu), replacing
// cv2.ORB_create(max_features).detectAndCompute(image, None) (reference scripts/lib/image.py:243-245, :324).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

namespace iam {

// Device block kept between calls (owned by the context): no allocation on the hot path.
struct OrbScratch {
  void* buf = nullptr;
  size_t cap = 0;
  OrbScratch() = default;
  OrbScratch(const OrbScratch&) = delete;
  OrbScratch& operator=(const OrbScratch&) = delete;
  ~OrbScratch();
};

// gray: HOST uint8 [h][w].  pattern256x4: the BRIEF tests (x0, y0, x1, y1) as int8.  Outputs (HOST): out_kp6
// [max_out][6] = x, y, size, angle (degrees), response, octave; out_des [max_out][32]; *out_n key points, listed
// level by level and in raster order inside a level.  Returns the number of kernels launched (>= 0) or a negative
// IAM_E_* code with `err` filled in.
int orb_detect(const uint8_t* gray, int w, int h, int nfeatures, const int8_t* pattern256x4, int max_out, float* out_kp6,
               uint8_t* out_des, int* out_n, OrbScratch* scratch, cudaStream_t stream, std::string* err);

// Debug aid: FAST-9/16 score map (no suppression) of the grey image, HOST in / HOST out [h][w].
int orb_debug_fast_scores(const uint8_t* gray, int w, int h, uint8_t* out_score, cudaStream_t stream, std::string* err);

// the pattern table compiled into the library
const int8_t* orb_pattern();

}  // namespace iam
