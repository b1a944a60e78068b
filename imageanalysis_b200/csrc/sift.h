// sift.h — SIFT detect + describe (sift.cu), replacing
// cv2.SIFT_create().detectAndCompute(image, None) (reference scripts/lib/image.py:236-237, :324).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

namespace iam {

// Device block kept between calls (owned by the context): the pyramid of one image size is allocated once.
struct SiftScratch {
  void* buf = nullptr;
  size_t cap = 0;
  void* pinned = nullptr;   // page-locked staging for the key point list
  size_t pinned_cap = 0;
  SiftScratch() = default;
  SiftScratch(const SiftScratch&) = delete;
  SiftScratch& operator=(const SiftScratch&) = delete;
  ~SiftScratch();
};

// gray: HOST uint8 [h][w].  Outputs (HOST): out_kp5 [max_out][5] = x, y, size, angle (degrees), response;
// out_octave [max_out] = cv2's packed octave field; out_des [max_out][128] (cv2 returns the same integers as
// float32); *out_n key points in cv2's order (sorted by x, y, size desc, angle, ... with duplicates removed).
// Returns the number of kernels launched (>= 0) or a negative code (-1 argument, -2 CUDA, -3 memory, -5 output
// buffers too small) with `err` filled in.
int sift_detect(const uint8_t* gray, int w, int h, int max_out, float* out_kp5, int32_t* out_octave, uint8_t* out_des,
                int* out_n, SiftScratch* scratch, cudaStream_t stream, std::string* err);

}  // namespace iam
