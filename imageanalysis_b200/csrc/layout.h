// layout.h — the HBM/shared-memory operand layouts shared by the conversion
// kernels and the tcgen05 kNN kernel.
//
// Every layout stores a descriptor ("row") as K-chunks of 16 bytes and groups
// rows by 8 into UMMA "core-matrix groups":
//   byte(row r, chunk c) = (r/8)*GROUP_BYTES + c*128 + (r%8)*16
// which is exactly the K-major SWIZZLE_NONE canonical layout with LBO=128,
// SBO=GROUP_BYTES, so a 128-row tile is one contiguous slab that a single
// cp.async.bulk drops into shared memory ready for tcgen05.mma.
//
// Wide layout (kind::f16 for L2 on fp16 operands, kind::f8f6f4 for Hamming): 18 chunks
//   chunks 0..15 : the 128 fp16 components (L2)  /  the 256 e4m3 bit values (Hamming)
//   chunks 16,17 : the augmentation K-step that folds ||q||^2 + ||t||^2 into the MMA
//   one form per role (a_form, b_form).
//
// Byte layout (kind::i8, integer-valued L2 descriptors in 0..255): 12 chunks, ONE form for both roles
//   chunks 0..7  : the 128 components as unsigned bytes
//   chunks 8,9   : query-role augmentation K-step: the 32 constant weights {1, 255, 255 x30}
//   chunks 10,11 : train-role augmentation K-step: 32 digits of G = CAP - floor(||t||^2 / 2) in the mixed
//                  radix of the weights
//   The MMA pairs query K-step 4 (chunks 8,9) with train K-step 4 (chunks 10,11), so the s32 accumulator is
//        acc = q.t + CAP - floor(||t||^2 / 2)        (LARGER = nearer)
//   and  d^2 = ||q||^2 + 2 CAP + (||t||^2 & 1) - 2 acc  exactly.  Rows are stored in RANK order: rows with even
//   ||t||^2 first, then the odd ones, each in original order (stable partition), then all-zero padding rows
//   (acc = 0 < any real acc).  Within a parity class acc orders exactly like d^2; across classes an odd row with
//   equal acc is farther by one, and it is always seen later, so strict comparisons keep cv2's order.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace iam {

enum class Kind { F16, F8, I8 };

constexpr int kGroupRows = 8;
constexpr int kTileRows = 128;
constexpr int kATiles = 2;                           // resident query tiles per work unit (A operand lives in TMEM)
constexpr int kSuperRows = kATiles * kTileRows;      // 256 rows of the query image one work unit owns
#ifndef IAM_BROWS
#define IAM_BROWS 96
#endif
constexpr int kBRows = IAM_BROWS;                    // train rows per streamed B tile (= UMMA N), multiple of 32
constexpr int kKStepBytes = 256;                     // one K-step = 2 chunks * 128 B per core-matrix group
constexpr uint32_t kLBO = 128;

template <Kind kKind>
struct Lay {  // wide layout: F16, F8
  static constexpr int kChunksPerRow = 18;
  static constexpr int kKSteps = 9;                  // 8 data K-steps + 1 augmentation step
  static constexpr int kAugA = 8 * kKStepBytes;      // byte offset of the augmentation K-step, query role
  static constexpr int kAugB = 8 * kKStepBytes;      // ... train role
  static constexpr int kBStages = 4;
};
template <>
struct Lay<Kind::I8> {
  static constexpr int kChunksPerRow = 12;
  static constexpr int kKSteps = 5;                  // 4 data K-steps + 1 augmentation step
  static constexpr int kAugA = 4 * kKStepBytes;      // chunks 8,9
  static constexpr int kAugB = 5 * kKStepBytes;      // chunks 10,11
#ifndef IAM_I8_STAGES
#define IAM_I8_STAGES 6
#endif
  static constexpr int kBStages = IAM_I8_STAGES;
};
template <Kind kKind>
struct LayD : Lay<kKind> {
  using L = Lay<kKind>;
  static constexpr int kRowBytes = L::kChunksPerRow * 16;
  static constexpr int kGroupBytes = kGroupRows * kRowBytes;   // = SBO
  static constexpr int kTileBytes = kTileRows * kRowBytes;
  static constexpr int kBTileBytes = kBRows * kRowBytes;
  static constexpr uint32_t kSBO = kGroupBytes;
  // byte offset (inside a core-matrix group) of K-step ks of the query / train role
  __host__ __device__ static constexpr int a_koff(int ks) { return ks < L::kKSteps - 1 ? ks * kKStepBytes : L::kAugA; }
  __host__ __device__ static constexpr int b_koff(int ks) { return ks < L::kKSteps - 1 ? ks * kKStepBytes : L::kAugB; }
};

// the wide layout's constants under their historical names (conversion kernels, debug entry points)
constexpr int kChunksPerRow = LayD<Kind::F16>::kChunksPerRow;
constexpr int kRowBytes = LayD<Kind::F16>::kRowBytes;        // 288
constexpr int kGroupBytes = LayD<Kind::F16>::kGroupBytes;    // 2304
constexpr int kTileBytes = LayD<Kind::F16>::kTileBytes;      // 36864
constexpr int kBTileBytes = LayD<Kind::F16>::kBTileBytes;    // 27648
constexpr int kKSteps = LayD<Kind::F16>::kKSteps;
constexpr uint32_t kSBO = kGroupBytes;

constexpr int kI8RowBytes = LayD<Kind::I8>::kRowBytes;       // 192
constexpr int kI8GroupBytes = LayD<Kind::I8>::kGroupBytes;   // 1536
// Mixed-radix capacity of the 32 augmentation slots {1, 255, 255*255 x30}: digits (g0, g1 <= 254; up to 30 x 255).
constexpr int kI8Cap = 254 + 255 * 254 + 30 * 65025;         // 2 015 774
// Eligibility: floor(||t||^2 / 2) <= kI8Cap - 1, so G >= 1 and every real row beats a padding row (acc 0).
constexpr int kI8MaxNorm = 2 * (kI8Cap - 1) + 1;             // 4 031 547  (real SIFT rows: ~262 144)

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline size_t form_bytes(int n_pad) { return static_cast<size_t>(n_pad) * kRowBytes; }
__host__ __device__ inline size_t i8_form_bytes(int n_pad) { return static_cast<size_t>(n_pad) * kI8RowBytes; }

// Per-image words at the start of the image's device block.
enum ImgMeta { kMetaExact = 0, kMetaI8Ok = 1, kMetaNEven = 2, kMetaWords = 64 };

// Per-image device record.
struct ImgDev {
  const uint8_t* a_form;   // wide layout, query-role operand (nullptr: not built)
  const uint8_t* b_form;   // wide layout, train-role operand
  const uint8_t* raw;      // packed u8 rows [n][raw_bytes] for the SIMT engine
  const int* kp_key;       // optional [n] keypoint-position ids for filter_duplicates (nullptr: none)
  const uint8_t* i8_form;  // byte layout (both roles), rows in rank order (nullptr: not built)
  const int* perm;         // [n_pad] rank -> original row (-1 for padding)
  const int* rowc;         // [n_pad] by rank: ||q||^2 + 2*kI8Cap
  const int* meta;         // ImgMeta words
  const float2* kp_xy;     // optional [n] keypoint pixel coordinates for the GMS filter (nullptr: none)
  int n;                   // valid descriptors
  int n_pad;               // rows allocated, multiple of kSuperRows and of kBRows
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
