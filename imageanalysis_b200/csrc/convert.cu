// convert.cu — descriptor upload-side layout conversion.
//
// Host rows (what OpenCV produced: float32 [N,128] integer-valued SIFT, or
// uint8 [N,32] ORB; reference image.py:160-180) become
//   raw    : packed u8 rows for the SIMT engine
//   a_form : query-role tcgen05 operand  [q , 1, 2048, 2048, 1, nq0..nq3 ...]
//   b_form : train-role tcgen05 operand  [-2t, nt0..nt3, 1, 2048, 2048, 1 ...]
// in the tiled layout of layout.h, so that A'.B' = ||q||^2 + ||t||^2 - 2 q.t
// = the exact squared distance.  For Hamming the 256 bits become e4m3
// {0,1} / {0,-2} bytes and the popcounts ride in the augmentation step.
// Integer-valued L2 descriptors additionally (or only) get the BYTE layout of
// layout.h for tcgen05 kind::i8 (convert_i8_kernel): one form for both roles,
// rows stably partitioned by the parity of their squared norm.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "knn.h"
#include "layout.h"

namespace iam {
namespace {

__device__ __forceinline__ size_t row_base(int r) {
  return static_cast<size_t>(r >> 3) * kGroupBytes + static_cast<size_t>(r & 7) * 16;
}

// e4m3 encodings of the small integers 0..16 and 256.
__device__ __forceinline__ uint8_t e4m3_small(int n) {
  const uint8_t tbl[17] = {0x00, 0x38, 0x40, 0x44, 0x48, 0x4A, 0x4C, 0x4E, 0x50,
                           0x51, 0x52, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58};
  return tbl[n];
}

// norm -> four fp16 terms with  n = x0 + 2048*x1 + 2048*x2 + x3  (exact for
// integer n < 2^24; ~33 significant bits otherwise)
__device__ __forceinline__ void split_norm(float n, __half (&x)[4]) {
  const float h2 = floorf(n * (1.0f / 4194304.0f));        // n / 2^22
  const float r = n - h2 * 4194304.0f;
  const float h1 = floorf(r * (1.0f / 2048.0f));
  const float h0 = r - h1 * 2048.0f;
  x[0] = __float2half_rn(h0);
  x[1] = __float2half_rn(h1);
  x[2] = __float2half_rn(h2 * 2048.0f);
  x[3] = __float2half_rn(h0 - __half2float(x[0]));
}

// One warp per row, 8 warps (one 8-row core-matrix group) per block.
// `bx`: block index inside the image (the batched kernel below serves many images per launch).  src may alias raw
// (a uint8 source copied straight to its final place): every thread reads its four bytes before it writes them.
template <typename SrcT>
__device__ __forceinline__ void convert_l2_body(const SrcT* src, int dim, int n, int n_pad, uint8_t* raw,
                                                uint8_t* __restrict__ a_form, uint8_t* __restrict__ b_form,
                                                int* __restrict__ meta, int* __restrict__ nrm_out,
                                                uint8_t* __restrict__ even_mask, int* __restrict__ ctx_flag, int bx) {
  __shared__ int s_even[8];
  const int lane = threadIdx.x & 31;
  const int r = bx * 8 + (threadIdx.x >> 5);  // n_pad is a multiple of 8: whole blocks only
  const bool valid = r < n;

  float v[4] = {0.f, 0.f, 0.f, 0.f};
  bool exact = true;
  if (valid) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = lane * 4 + e;
      if (k < dim) {
        const float x = static_cast<float>(src[static_cast<size_t>(r) * dim + k]);
        exact = exact && (x == rintf(x)) && (x >= 0.f) && (x <= 255.f);
        v[e] = x;
      }
    }
  }
  __half h[4];
  float nrm = 0.f;
  uint32_t rawword = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    h[e] = __float2half_rn(v[e]);
    const float hv = __half2float(h[e]);
    nrm += hv * hv;  // exact for integer inputs: every partial < 2^24
    const int q = static_cast<int>(fminf(fmaxf(rintf(v[e]), 0.f), 255.f));
    rawword |= static_cast<uint32_t>(q) << (8 * e);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
  const bool row_exact = __all_sync(0xffffffffu, exact);
  if (!row_exact && lane == 0) atomicAnd(meta + kMetaExact, 0);
  if (nrm_out) {  // inputs of the byte-layout pass: squared norm, its parity, eligibility (layout.h)
    const bool ok = row_exact && nrm <= static_cast<float>(kI8MaxNorm);
    const int inrm = ok ? static_cast<int>(nrm) : 0;
    if (lane == 0) {
      nrm_out[r] = inrm;
      s_even[threadIdx.x >> 5] = (valid && !(inrm & 1)) ? 1 : 0;
      if (!ok) {
        atomicAnd(meta + kMetaI8Ok, 0);
        if (ctx_flag) atomicAnd(ctx_flag, 0);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int m = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) m |= s_even[w] << w;
      even_mask[bx] = static_cast<uint8_t>(m);
    }
  }

  // raw u8 row (dim bytes, word `lane`)
  if (lane * 4 < dim) reinterpret_cast<uint32_t*>(raw + static_cast<size_t>(r) * dim)[lane] = rawword;

  if (!a_form) return;  // wide forms not wanted (byte layout only)
  // data chunks: lane -> chunk lane/2, half-chunk lane%2 (8 bytes = 4 halfs)
  const size_t base = row_base(r) + static_cast<size_t>(lane >> 1) * 128 + static_cast<size_t>(lane & 1) * 8;
  {
    __half2 a01 = __halves2half2(h[0], h[1]), a23 = __halves2half2(h[2], h[3]);
    uint2 av;
    av.x = *reinterpret_cast<uint32_t*>(&a01);
    av.y = *reinterpret_cast<uint32_t*>(&a23);
    *reinterpret_cast<uint2*>(a_form + base) = av;
    __half2 b01 = __halves2half2(__float2half_rn(-2.f * __half2float(h[0])), __float2half_rn(-2.f * __half2float(h[1])));
    __half2 b23 = __halves2half2(__float2half_rn(-2.f * __half2float(h[2])), __float2half_rn(-2.f * __half2float(h[3])));
    uint2 bv;
    bv.x = *reinterpret_cast<uint32_t*>(&b01);
    bv.y = *reinterpret_cast<uint32_t*>(&b23);
    *reinterpret_cast<uint2*>(b_form + base) = bv;
  }
  // augmentation chunks 16 and 17 (16 halfs): lanes 0..15 write one half each
  if (lane < 16) {
    __half x[4];
    split_norm(nrm, x);
    const __half one = __float2half_rn(1.f), k2048 = __float2half_rn(2048.f), zero = __float2half_rn(0.f);
    __half av = zero, bv = zero;
    if (valid) {
      // A: [1, 2048, 2048, 1, n0, n1, n2, n3, 0...]   B: [n0, n1, n2, n3, 1, 2048, 2048, 1, 0...]
      switch (lane) {
        case 0: av = one;   bv = x[0]; break;
        case 1: av = k2048; bv = x[1]; break;
        case 2: av = k2048; bv = x[2]; break;
        case 3: av = one;   bv = x[3]; break;
        case 4: av = x[0];  bv = one; break;
        case 5: av = x[1];  bv = k2048; break;
        case 6: av = x[2];  bv = k2048; break;
        case 7: av = x[3];  bv = one; break;
        default: break;
      }
    } else {
      // padding train rows: distance = ||q||^2 + 2^24, larger than any real one
      if (lane == 1) bv = __float2half_rn(8192.f);
    }
    const size_t abase = row_base(r) + static_cast<size_t>(16 + (lane >> 3)) * 128 + static_cast<size_t>(lane & 7) * 2;
    *reinterpret_cast<__half*>(a_form + abase) = av;
    *reinterpret_cast<__half*>(b_form + abase) = bv;
  }
}


template <typename SrcT>
__global__ void __launch_bounds__(256) convert_l2_kernel(const SrcT* __restrict__ src, int dim, int n, int n_pad,
                                                         uint8_t* __restrict__ raw, uint8_t* __restrict__ a_form,
                                                         uint8_t* __restrict__ b_form, int* __restrict__ meta,
                                                         int* __restrict__ nrm_out, uint8_t* __restrict__ even_mask,
                                                         int* __restrict__ ctx_flag) {
  convert_l2_body<SrcT>(src, dim, n, n_pad, raw, a_form, b_form, meta, nrm_out, even_mask, ctx_flag, blockIdx.x);
}

// One launch for a whole wave of uint8 images (iam_match_images): blockIdx.y = image, byte layout only.
__global__ void __launch_bounds__(256) convert_l2_u8_batch_kernel(const ConvJob* __restrict__ jobs, int dim, int* __restrict__ ctx_flag) {
  const ConvJob j = jobs[blockIdx.y];
  if (static_cast<int>(blockIdx.x) * 8 >= j.n_pad) return;   // whole block: the body's barrier is safe
  convert_l2_body<uint8_t>(j.src, dim, j.n, j.n_pad, j.raw, nullptr, nullptr, j.meta, j.nrm, reinterpret_cast<uint8_t*>(j.even_mask),
                           ctx_flag, blockIdx.x);
}

// Byte layout (layout.h): 32 rows per block (one word of the even-norm mask), one warp per 4 rows.
// rank(r) = #even rows before r                      (||r||^2 even)
//         = #even rows in all + #odd rows before r   (odd)          -> stable partition, padding rows stay in place.
__device__ __forceinline__ void convert_i8_body(const uint8_t* __restrict__ raw, int dim, int n, int n_pad,
                                                const int* __restrict__ nrm, const uint32_t* __restrict__ even_mask,
                                                uint8_t* __restrict__ form, int* __restrict__ perm,
                                                int* __restrict__ rowc, int* __restrict__ meta, int bx) {
  __shared__ int s_before[8], s_total[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_words = n_pad >> 5;
  int before = 0, total = 0;
  for (int w = threadIdx.x; w < n_words; w += 256) {
    const int c = __popc(even_mask[w]);
    total += c;
    if (w < bx) before += c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    before += __shfl_xor_sync(0xffffffffu, before, o);
    total += __shfl_xor_sync(0xffffffffu, total, o);
  }
  if (lane == 0) {
    s_before[warp] = before;
    s_total[warp] = total;
  }
  __syncthreads();
  before = total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    before += s_before[w];
    total += s_total[w];
  }
  if (bx == 0 && threadIdx.x == 0) meta[kMetaNEven] = total;
  const uint32_t word = even_mask[bx];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int bit = warp * 4 + i;
    const int r = bx * 32 + bit;
    const bool valid = r < n;
    const int ev_before = before + __popc(word & ((1u << bit) - 1u));
    const bool even = (word >> bit) & 1u;
    const int rank = !valid ? r : (even ? ev_before : total + (r - ev_before));
    uint32_t data = 0;
    if (valid && lane * 4 < dim) data = reinterpret_cast<const uint32_t*>(raw + static_cast<size_t>(r) * dim)[lane];
    const size_t base = static_cast<size_t>(rank >> 3) * kI8GroupBytes + static_cast<size_t>(rank & 7) * 16;
    *reinterpret_cast<uint32_t*>(form + base + static_cast<size_t>(lane >> 2) * 128 + (lane & 3) * 4) = data;
    if (lane < 16) {
      uint32_t aug = 0;
      if (valid) {
        if (lane < 8) {  // query role: weights {1, 255, 255 x30}
          aug = lane == 0 ? 0xFFFFFF01u : 0xFFFFFFFFu;
        } else {         // train role: digits of G = CAP - floor(norm/2)
          int g = kI8Cap - (nrm[r] >> 1);
          g = g < 0 ? 0 : g;
          const int g2 = g / 65025;
          const int rem = g - g2 * 65025;
          const int g1 = rem / 255;
          const int g0 = rem - g1 * 255;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int slot = (lane - 8) * 4 + b;
            const int v = slot == 0 ? g0 : (slot == 1 ? g1 : ((slot - 2) < g2 ? 255 : 0));
            aug |= static_cast<uint32_t>(v) << (8 * b);
          }
        }
      }
      *reinterpret_cast<uint32_t*>(form + base + static_cast<size_t>(8 + (lane >> 2)) * 128 + (lane & 3) * 4) = aug;
    }
    if (lane == 0) {
      perm[rank] = valid ? r : -1;
      rowc[rank] = valid ? nrm[r] + 2 * kI8Cap : 0;
    }
  }
}

__global__ void __launch_bounds__(256) convert_i8_kernel(const uint8_t* __restrict__ raw, int dim, int n, int n_pad,
                                                         const int* __restrict__ nrm, const uint32_t* __restrict__ even_mask,
                                                         uint8_t* __restrict__ form, int* __restrict__ perm,
                                                         int* __restrict__ rowc, int* __restrict__ meta) {
  convert_i8_body(raw, dim, n, n_pad, nrm, even_mask, form, perm, rowc, meta, blockIdx.x);
}

__global__ void __launch_bounds__(256) convert_i8_batch_kernel(const ConvJob* __restrict__ jobs, int dim) {
  const ConvJob j = jobs[blockIdx.y];
  if (static_cast<int>(blockIdx.x) * 32 >= j.n_pad) return;
  convert_i8_body(j.raw, dim, j.n, j.n_pad, j.nrm, j.even_mask, j.form, j.perm, j.rowc, j.meta, blockIdx.x);
}

__global__ void __launch_bounds__(256) convert_hamming_kernel(const uint8_t* __restrict__ src, int nbytes, int n,
                                                              int n_pad, uint8_t* __restrict__ raw,
                                                              uint8_t* __restrict__ a_form,
                                                              uint8_t* __restrict__ b_form) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n_pad) return;
  const bool valid = r < n;
  uint32_t byte = 0;
  if (valid && lane < nbytes) byte = src[static_cast<size_t>(r) * nbytes + lane];
  if (lane < nbytes) raw[static_cast<size_t>(r) * nbytes + lane] = static_cast<uint8_t>(byte);
  int pc = __popc(byte);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pc += __shfl_xor_sync(0xffffffffu, pc, o);

  // element index = 8*lane + bit (bit 0 = LSB; any fixed order works as both operands agree)
  uint32_t a_lo = 0, a_hi = 0, b_lo = 0, b_hi = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    if (byte & (1u << b)) {
      a_lo |= 0x38u << (8 * b);  // 1.0
      b_lo |= 0xE8u << (8 * b);  // -64.0 = -2 * 32 (the accumulator is a KEY: 32 * distance + column, see below)
    }
    if (byte & (1u << (b + 4))) {
      a_hi |= 0x38u << (8 * b);
      b_hi |= 0xE8u << (8 * b);
    }
  }
  const size_t base = row_base(r) + static_cast<size_t>(lane >> 1) * 128 + static_cast<size_t>(lane & 1) * 8;
  *reinterpret_cast<uint2*>(a_form + base) = make_uint2(a_lo, a_hi);
  *reinterpret_cast<uint2*>(b_form + base) = make_uint2(b_lo, b_hi);

  // augmentation: 32 e4m3 elements in chunks 16,17; one byte per lane.  The accumulator of (query q, train t) is
  //     32 * (|q| + |t| - 2 q.t) + (t's row index mod 32)  =  32 * Hamming distance + column inside the 32-column
  // slice an epilogue thread holds: an exact, unique key that orders like (distance, column) -- the packing that
  // the byte layout does with 32 IMADs per slice comes out of the tensor core for free (knn_umma.cu).
  // pc = d0 + 16*d1, d1 <= 16;  32*pc = 32*d0 + 256*(2*d1), every factor an e4m3 integer.
  const int d0 = pc & 15, d1 = pc >> 4;
  const uint8_t two_d1 = d1 ? static_cast<uint8_t>(e4m3_small(d1) + 0x08) : 0;  // exponent + 1
  uint8_t av = 0, bv = 0;
  // A: [32, 256, p0, 2*p1, 256, 1, 1, ...]   B: [p0, 2*p1, 32, 256, pad, j & 15, j & 16, ...]
  switch (lane) {
    case 0: av = 0x60; bv = valid ? e4m3_small(d0) : 0; break;
    case 1: av = 0x78; bv = valid ? two_d1 : 0; break;
    case 2: av = valid ? e4m3_small(d0) : 0; bv = 0x60; break;
    case 3: av = valid ? two_d1 : 0; bv = 0x78; break;
    case 4: av = 0x78; bv = valid ? 0x00 : 0x78; break;  // padding train rows: 256 * 256 = 65536 > any key
    case 5: av = 0x38; bv = e4m3_small(r & 15); break;
    case 6: av = 0x38; bv = (r & 16) ? 0x58 : 0x00; break;
    default: break;
  }
  const size_t abase = row_base(r) + static_cast<size_t>(16 + (lane >> 4)) * 128 + static_cast<size_t>(lane & 15);
  a_form[abase] = av;
  b_form[abase] = bv;
}

__global__ void finish_dist_kernel(int norm, float* __restrict__ d, int* __restrict__ idx, size_t count) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float x = d[i];
  const float pad_thr = (norm == 0) ? 16777216.0f : 257.0f;  // padding rows sit at >= 2^24 (L2) / >= 4096 (Hamming)
  if (x >= pad_thr || idx[i] < 0) {  // neighbour slot filled by a padding row (fewer than k train rows)
    idx[i] = -1;
    d[i] = __int_as_float(0x7f800000);
    return;
  }
  if (norm == 0) d[i] = sqrtf(x);  // correctly rounded; matches cv2's float sqrt of the exact sum
}

}  // namespace

cudaError_t launch_convert(int norm, int raw_bytes, const void* src, int src_dtype, int n, int n_pad, uint8_t* raw,
                           uint8_t* a_form, uint8_t* b_form, int* meta, const I8Out* i8, cudaStream_t stream) {
  const int blocks = n_pad / 8;
  if (blocks <= 0) return cudaSuccess;
  if (norm == 0) {
    if (raw_bytes > 128 || (raw_bytes & 3)) return cudaErrorInvalidValue;
    int* nrm = i8 ? i8->nrm : nullptr;
    uint8_t* mask = i8 ? reinterpret_cast<uint8_t*>(i8->even_mask) : nullptr;
    int* ctx_flag = i8 ? i8->ctx_flag : nullptr;
    if (src_dtype == 0)
      convert_l2_kernel<uint8_t><<<blocks, 256, 0, stream>>>(static_cast<const uint8_t*>(src), raw_bytes, n, n_pad,
                                                             raw, a_form, b_form, meta, nrm, mask, ctx_flag);
    else
      convert_l2_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(src), raw_bytes, n, n_pad, raw,
                                                           a_form, b_form, meta, nrm, mask, ctx_flag);
    if (i8 && i8->form)
      convert_i8_kernel<<<n_pad / 32, 256, 0, stream>>>(raw, raw_bytes, n, n_pad, nrm, i8->even_mask, i8->form, i8->perm,
                                                        i8->rowc, meta);
  } else {
    if (raw_bytes > 32 || src_dtype != 0) return cudaErrorInvalidValue;
    convert_hamming_kernel<<<blocks, 256, 0, stream>>>(static_cast<const uint8_t*>(src), raw_bytes, n, n_pad, raw,
                                                       a_form, b_form);
  }
  return cudaGetLastError();
}

cudaError_t launch_convert_u8_batch(const ConvJob* d_jobs, int n_jobs, int max_n_pad, int raw_bytes, int* ctx_flag,
                                    cudaStream_t stream) {
  if (n_jobs <= 0 || max_n_pad <= 0) return cudaSuccess;
  if (raw_bytes > 128 || (raw_bytes & 3) || n_jobs > 65535) return cudaErrorInvalidValue;
  convert_l2_u8_batch_kernel<<<dim3(max_n_pad / 8, n_jobs), 256, 0, stream>>>(d_jobs, raw_bytes, ctx_flag);
  convert_i8_batch_kernel<<<dim3(max_n_pad / 32, n_jobs), 256, 0, stream>>>(d_jobs, raw_bytes);
  return cudaGetLastError();
}

cudaError_t launch_finish_dist(int norm, float* d, int* idx, size_t count, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  const int threads = 256;
  const size_t blocks = (count + threads - 1) / threads;
  finish_dist_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(norm, d, idx, count);
  return cudaGetLastError();
}

}  // namespace iam
