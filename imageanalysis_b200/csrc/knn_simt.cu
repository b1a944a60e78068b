// knn_simt.cu — exact CUDA-core kNN engine (dp4a for L2 on u8 rows, xor+popc
// for Hamming).  It is the on-device cross-check for the tcgen05 engine and
// serves descriptor sizes the tensor-core layout does not cover.  Same
// contract as cv2.BFMatcher.knnMatch (reference matcher.py:212): ascending
// distance, ties to the lowest train index.
#include <cuda_runtime.h>

#include "knn.h"
#include "layout.h"

namespace iam {
namespace {

constexpr int kTrainTile = 64;
constexpr float kInf = 3.0e38f;

template <int NORM, int WORDS, int KTOP>
__global__ void __launch_bounds__(128) knn_simt_kernel(const ImgDev* __restrict__ imgs,
                                                       const KnnUnit* __restrict__ units, int* __restrict__ out_idx,
                                                       float* __restrict__ out_d2) {
  __shared__ __align__(16) uint32_t s_t[kTrainTile * WORDS];
  __shared__ uint32_t s_norm[kTrainTile];

  constexpr int kSub = kSuperRows / 128;
  const KnnUnit unit = units[blockIdx.x / kSub];
  const ImgDev q = imgs[unit.q_slot];
  const ImgDev t = imgs[unit.t_slot];
  const int row = unit.super * kSuperRows + (blockIdx.x % kSub) * 128 + threadIdx.x;

  uint32_t qw[WORDS];
  {
    const uint32_t* qp = reinterpret_cast<const uint32_t*>(q.raw) + static_cast<size_t>(row) * WORDS;
#pragma unroll
    for (int w = 0; w < WORDS; ++w) qw[w] = qp[w];  // rows < n_pad are always allocated (zero padded)
  }
  uint32_t qn = 0;
  if (NORM == 0) {
#pragma unroll
    for (int w = 0; w < WORDS; ++w) qn = __dp4a(qw[w], qw[w], qn);
  }

  float bd[KTOP];
  int bi[KTOP];
#pragma unroll
  for (int s = 0; s < KTOP; ++s) {
    bd[s] = kInf;
    bi[s] = -1;
  }

  for (int t0 = 0; t0 < t.n; t0 += kTrainTile) {
    const int rows = min(kTrainTile, t.n - t0);
    __syncthreads();
    const uint32_t* tp = reinterpret_cast<const uint32_t*>(t.raw) + static_cast<size_t>(t0) * WORDS;
    for (int i = threadIdx.x; i < rows * WORDS; i += blockDim.x) s_t[i] = tp[i];
    __syncthreads();
    if (NORM == 0 && threadIdx.x < rows) {
      uint32_t nn = 0;
#pragma unroll
      for (int w = 0; w < WORDS; ++w) {
        const uint32_t x = s_t[threadIdx.x * WORDS + w];
        nn = __dp4a(x, x, nn);
      }
      s_norm[threadIdx.x] = nn;
    }
    __syncthreads();
    for (int j = 0; j < rows; ++j) {
      uint32_t acc = 0;
#pragma unroll
      for (int w = 0; w < WORDS; ++w) {
        const uint32_t x = s_t[j * WORDS + w];  // warp-uniform address: broadcast
        if (NORM == 0)
          acc = __dp4a(qw[w], x, acc);
        else
          acc += __popc(qw[w] ^ x);
      }
      const uint32_t dist = (NORM == 0) ? (qn + s_norm[j] - 2u * acc) : acc;
      const float x = static_cast<float>(dist);  // < 2^24: exact
      if (x < bd[KTOP - 1]) {
#pragma unroll
        for (int s = KTOP - 1; s >= 0; --s) {
          const bool lt_prev = (s > 0) ? (x < bd[s > 0 ? s - 1 : 0]) : false;
          const bool lt_cur = x < bd[s];
          if (s > 0) {
            bd[s] = lt_prev ? bd[s - 1] : (lt_cur ? x : bd[s]);
            bi[s] = lt_prev ? bi[s - 1] : (lt_cur ? (t0 + j) : bi[s]);
          } else {
            bd[s] = lt_cur ? x : bd[s];
            bi[s] = lt_cur ? (t0 + j) : bi[s];
          }
        }
      }
    }
  }
  if (row < q.n) {
    const size_t o = (static_cast<size_t>(unit.out_base) + row) * KTOP;
#pragma unroll
    for (int s = 0; s < KTOP; ++s) {
      out_idx[o + s] = bi[s];
      out_d2[o + s] = bd[s];
    }
  }
}

template <int NORM, int WORDS>
cudaError_t launch_w(int k, const ImgDev* imgs, const KnnUnit* units, int n_units, int* out_idx, float* out_d2,
                     cudaStream_t stream) {
  const int grid = n_units * (kSuperRows / 128);
  switch (k) {
    case 1: knn_simt_kernel<NORM, WORDS, 1><<<grid, 128, 0, stream>>>(imgs, units, out_idx, out_d2); break;
    case 2: knn_simt_kernel<NORM, WORDS, 2><<<grid, 128, 0, stream>>>(imgs, units, out_idx, out_d2); break;
    case 3: knn_simt_kernel<NORM, WORDS, 3><<<grid, 128, 0, stream>>>(imgs, units, out_idx, out_d2); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

template <int NORM>
cudaError_t launch_n(int k, int raw_bytes, const ImgDev* imgs, const KnnUnit* units, int n_units, int* out_idx,
                     float* out_d2, cudaStream_t stream) {
  switch (raw_bytes) {
    case 32: return launch_w<NORM, 8>(k, imgs, units, n_units, out_idx, out_d2, stream);
    case 64: return launch_w<NORM, 16>(k, imgs, units, n_units, out_idx, out_d2, stream);
    case 128: return launch_w<NORM, 32>(k, imgs, units, n_units, out_idx, out_d2, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

cudaError_t launch_knn_simt(int norm, int k, int raw_bytes, const ImgDev* imgs, const KnnUnit* units, int n_units,
                            int* out_idx, float* out_d2, cudaStream_t stream) {
  if (n_units <= 0) return cudaSuccess;
  return norm == 0 ? launch_n<0>(k, raw_bytes, imgs, units, n_units, out_idx, out_d2, stream)
                   : launch_n<1>(k, raw_bytes, imgs, units, n_units, out_idx, out_d2, stream);
}

}  // namespace iam
