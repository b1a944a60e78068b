// knn.h — internal launcher declarations (device code lives in the .cu files).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "layout.h"

namespace iam {

// tcgen05 engine. norm: 0 = L2 (kind::f16), 1 = Hamming (kind::f8f6f4). k in {1,2,3}.
// out_idx / out_d2 are [rows][k]; out_d2 holds the exact squared L2 distance
// (or the Hamming distance) as fp32.
cudaError_t launch_knn_umma(int norm, int k, const ImgDev* imgs, const KnnUnit* units, int n_units, int* out_idx,
                            float* out_d2, int num_sms, cudaStream_t stream);

// Debug aid: accumulators of one 128x128 tile with run-time descriptor strides -> out[128*128].
cudaError_t launch_umma_tile_debug(int norm, const uint8_t* a_tile, const uint8_t* b_tile, uint32_t lbo, uint32_t sbo,
                                   uint32_t kstep_bytes, int ksteps, float* out, cudaStream_t stream);

// Exact CUDA-core engine on the packed u8 rows (dp4a / xor+popc); same outputs.
cudaError_t launch_knn_simt(int norm, int k, int raw_bytes, const ImgDev* imgs, const KnnUnit* units, int n_units,
                            int* out_idx, float* out_d2, cudaStream_t stream);

// Descriptor conversion: host-layout rows -> raw u8 rows + tiled A/B operand forms.
// src_dtype: 0 = u8, 1 = f32.  `exact_flag` (device int, pre-set to 1) is cleared
// if any L2 component is not an integer in [0,255].
cudaError_t launch_convert(int norm, int raw_bytes, const void* src, int src_dtype, int n, int n_pad, uint8_t* raw,
                           uint8_t* a_form, uint8_t* b_form, int* exact_flag, cudaStream_t stream);

// d2 -> distance (sqrt for L2, identity for Hamming), in place, and
// invalidation of neighbours that point at padding rows.
cudaError_t launch_finish_dist(int norm, float* d, int* idx, size_t count, cudaStream_t stream);

}  // namespace iam
