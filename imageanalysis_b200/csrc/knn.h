// knn.h — internal launcher declarations (device code lives in the .cu files).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "layout.h"

namespace iam {

// tcgen05 engine.  kind: kKindF16 = L2 on fp16 operands (kind::f16), kKindF8 = Hamming (kind::f8f6f4),
// kKindI8 = L2 on integer-valued byte operands (kind::i8, byte layout of layout.h).  k in {1,2,3}.
// out_idx / out_d2 are [rows][k]; out_d2 holds the exact squared L2 distance
// (or the Hamming distance) as fp32.
constexpr int kKindF16 = 0, kKindF8 = 1, kKindI8 = 2;
cudaError_t launch_knn_umma(int kind, int k, const ImgDev* imgs, const KnnUnit* units, int n_units, int* out_idx,
                            float* out_d2, int num_sms, cudaStream_t stream);

// Debug aid: accumulators of one 128x128 tile with run-time descriptor strides -> out[128*128]
// (kKindI8: a_tile and b_tile are byte-layout tiles, the strides are the layout's own).
cudaError_t launch_umma_tile_debug(int kind, const uint8_t* a_tile, const uint8_t* b_tile, uint32_t lbo, uint32_t sbo,
                                   uint32_t kstep_bytes, int ksteps, float* out, cudaStream_t stream);

// Exact CUDA-core engine on the packed u8 rows (dp4a / xor+popc); same outputs.
cudaError_t launch_knn_simt(int norm, int k, int raw_bytes, const ImgDev* imgs, const KnnUnit* units, int n_units,
                            int* out_idx, float* out_d2, cudaStream_t stream);

// Outputs of the byte-layout pass of an L2 conversion (all device pointers).
struct I8Out {
  uint8_t* form;        // [n_pad] byte-layout rows in rank order; nullptr: only norms / flags are produced
  int* perm;            // [n_pad] rank -> original row
  int* rowc;            // [n_pad] by rank: squared norm + 2 * kI8Cap
  int* nrm;             // [n_pad] scratch: squared norms by original row
  uint32_t* even_mask;  // [n_pad / 32] scratch: bit r = row r valid and its squared norm even
  int* ctx_flag;        // optional context-wide sticky word, cleared when a row is not eligible
};

// One image of a batched byte-layout conversion (launch_convert_u8_batch); all device pointers.
struct ConvJob {
  const uint8_t* src;   // [n][raw_bytes] uint8 rows (may be `raw` itself: rows copied straight to their final place)
  uint8_t* raw;         // [n_pad][raw_bytes]
  int* meta;            // the image's flag words (pre-set non-zero)
  int* nrm;
  uint32_t* even_mask;
  uint8_t* form;
  int* perm;
  int* rowc;
  int n, n_pad;
};

// The byte-layout conversion of launch_convert(norm = L2, src_dtype = u8, wide forms skipped) for n_jobs images in
// two launches (blockIdx.y = image).
cudaError_t launch_convert_u8_batch(const ConvJob* d_jobs, int n_jobs, int max_n_pad, int raw_bytes, int* ctx_flag,
                                    cudaStream_t stream);

// Descriptor conversion: host-layout rows -> raw u8 rows + tiled operand forms.
// src_dtype: 0 = u8, 1 = f32.  meta[kMetaExact] (pre-set non-zero) is cleared if any L2 component is not an
// integer in [0,255]; with `i8`, meta[kMetaI8Ok] (pre-set non-zero) is cleared if a row is not eligible for the
// byte layout and meta[kMetaNEven] receives the number of even-norm rows.  a_form == nullptr skips the wide forms.
cudaError_t launch_convert(int norm, int raw_bytes, const void* src, int src_dtype, int n, int n_pad, uint8_t* raw,
                           uint8_t* a_form, uint8_t* b_form, int* meta, const I8Out* i8, cudaStream_t stream);

// d2 -> distance (sqrt for L2, identity for Hamming), in place, and
// invalidation of neighbours that point at padding rows.
cudaError_t launch_finish_dist(int norm, float* d, int* idx, size_t count, cudaStream_t stream);

}  // namespace iam
