// layout.h — the HBM/shared-memory operand layout shared by the conversion
// kernels and the tcgen05 kNN kernel.
//
// One descriptor ("row") is stored as 18 K-chunks of 16 bytes:
//   chunks 0..15 : the 128 fp16 components (L2)  /  the 256 e4m3 bit values (Hamming)
//   chunks 16,17 : the augmentation K-step that folds ||q||^2 + ||t||^2 into the MMA
// Rows are grouped by 8 into UMMA "core-matrix groups":
//   byte(row r, chunk c) = (r/8)*GROUP_BYTES + c*128 + (r%8)*16
// which is exactly the K-major SWIZZLE_NONE canonical layout with LBO=128, SBO=2304,
// so a 128-row tile is one contiguous 36,864-byte slab that a single
// cp.async.bulk drops into shared memory ready for tcgen05.mma.
#pragma once
#include <cstdint>

namespace iam {

constexpr int kChunksPerRow = 18;
constexpr int kRowBytes = kChunksPerRow * 16;        // 288
constexpr int kGroupRows = 8;
constexpr int kGroupBytes = kGroupRows * kRowBytes;  // 2304
constexpr int kTileRows = 128;
constexpr int kTileBytes = kTileRows * kRowBytes;    // 36864
constexpr int kATiles = 2;                           // resident query tiles per work unit (A operand lives in TMEM)
constexpr int kSuperRows = kATiles * kTileRows;      // 256 rows of the query image one work unit owns
#ifndef IAM_BROWS
#define IAM_BROWS 96
#endif
constexpr int kBRows = IAM_BROWS;                    // train rows per streamed B tile (= UMMA N), multiple of 32
constexpr int kBTileBytes = kBRows * kRowBytes;      // 27648
constexpr int kKSteps = 9;                           // 8 data K-steps + 1 augmentation step
constexpr int kKStepBytes = 256;                     // 2 chunks * 128 B
constexpr uint32_t kLBO = 128;
constexpr uint32_t kSBO = kGroupBytes;

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline size_t form_bytes(int n_pad) { return static_cast<size_t>(n_pad) * kRowBytes; }

// Per-image device record.
struct ImgDev {
  const uint8_t* a_form;  // query-role operand (tiled layout above)
  const uint8_t* b_form;  // train-role operand
  const uint8_t* raw;     // packed u8 rows [n][raw_bytes] for the SIMT engine
  const int* kp_key;      // optional [n] keypoint-position ids for filter_duplicates (nullptr: none)
  int n;                  // valid descriptors
  int n_pad;              // rows allocated, multiple of kSuperRows
};

// One unit of work for the kNN kernels: kSuperRows query rows of q_slot against all
// descriptors of t_slot.
struct KnnUnit {
  int q_slot;
  int t_slot;
  int super;     // query supertile index (rows super*256 ..)
  int out_base;  // first output row of this directed job
};

}  // namespace iam
