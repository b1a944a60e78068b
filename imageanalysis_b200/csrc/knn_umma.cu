// knn_umma.cu — fused all-pairs descriptor distance + top-k on the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM), replacing
// cv2.DescriptorMatcher.knnMatch as called from raw_matches()
// (reference scripts/lib/matcher.py:203-216).
//
// One work unit = 256 query descriptors (two 128-row A tiles, resident in
// shared memory) against every descriptor of the train image (128-row B tiles
// streamed by cp.async.bulk through a 4-deep mbarrier ring).  The augmented
// K-step makes every accumulator element the exact squared L2 distance
// (integer valued < 2^23, exact in fp32) or the exact Hamming distance, so
// the N x M distance matrix never leaves the SM: eight epilogue warps read
// it from TMEM and keep a per-row running top-k in registers.
//
// Warp roles (640 threads, 1 CTA / SM, persistent over units):
//   warp 0      : B-tile producer  (bulk copy  -> b_full[stage])
//   warp 1      : MMA issuer       (one elected lane issues 18 tcgen05.mma / B tile)
//   warp 2      : TMEM allocator / deallocator
//   warp 3      : A-tile producer  (bulk copy  -> a_full[half])
//   warps 4..19 : epilogue (4 per SM sub-partition): warp -> (accumulator, 64-column part,
//                 TMEM lane quadrant = warp_id % 4); the two column parts of a row merge
//                 their top-k lists through shared memory once per unit
#include <cuda_runtime.h>

#include <cstdlib>

#include "knn.h"
#include "layout.h"
#include "ptx.cuh"

namespace iam {

namespace {

constexpr int kBStages = 4;
constexpr int kAccStages = 2;
constexpr uint32_t kTmemCols = 512;  // 2 stages x 2 accumulators x 128 fp32 columns

struct __align__(8) Barriers {
  uint64_t a_full[2];
  uint64_t a_empty[2];
  uint64_t b_full[kBStages];
  uint64_t b_empty[kBStages];
  uint64_t t_full[kAccStages];
  uint64_t t_empty[kAccStages];
  uint32_t tmem_base;
  uint32_t pad;
};

constexpr size_t kSmemA = 2 * kTileBytes;
constexpr size_t kSmemB = kBStages * kTileBytes;
constexpr size_t kSmemMerge = 2 * 128 * 3 * 8;  // [A tile][row][k] (dist, idx) hand-over between column parts
constexpr size_t kSmemTotal = kSmemA + kSmemB + sizeof(Barriers) + kSmemMerge + 128;

constexpr float kInf = 3.0e38f;

template <int KTOP>
struct TopK {
  float d[KTOP];
  int i[KTOP];
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int s = 0; s < KTOP; ++s) {
      d[s] = kInf;
      i[s] = 0x7fffffff;
    }
  }
  __device__ __forceinline__ float thr() const { return d[KTOP - 1]; }
  // Insert (x, col); caller guarantees x < thr().  Columns arrive in ascending
  // order, so strict '<' keeps the earliest (lowest) column first among equal
  // distances, which is the order cv2.BFMatcher reports ties in.
  __device__ __forceinline__ void insert(float x, int col) {
#pragma unroll
    for (int s = KTOP - 1; s >= 0; --s) {
      const bool lt_prev = (s > 0) ? (x < d[s > 0 ? s - 1 : 0]) : false;
      const bool lt_cur = x < d[s];
      if (s > 0) {
        d[s] = lt_prev ? d[s - 1] : (lt_cur ? x : d[s]);
        i[s] = lt_prev ? i[s - 1] : (lt_cur ? col : i[s]);
      } else {
        d[s] = lt_cur ? x : d[s];
        i[s] = lt_cur ? col : i[s];
      }
    }
  }
  // Order-independent insert: (distance, index) lexicographic.  Used to merge the
  // partial lists of the two column parts of a row.
  __device__ __forceinline__ void insert_lex(float x, int col) {
#pragma unroll
    for (int s = KTOP - 1; s >= 0; --s) {
      const bool lt_prev = (s > 0) ? (x < d[s > 0 ? s - 1 : 0] || (x == d[s > 0 ? s - 1 : 0] && col < i[s > 0 ? s - 1 : 0])) : false;
      const bool lt_cur = x < d[s] || (x == d[s] && col < i[s]);
      if (s > 0) {
        d[s] = lt_prev ? d[s - 1] : (lt_cur ? x : d[s]);
        i[s] = lt_prev ? i[s - 1] : (lt_cur ? col : i[s]);
      } else {
        d[s] = lt_cur ? x : d[s];
        i[s] = lt_cur ? col : i[s];
      }
    }
  }
};

__device__ __forceinline__ float min8(const float* w) {
  return fminf(fmin3(fmin3(w[0], w[1], w[2]), w[3], w[4]), fmin3(w[5], w[6], w[7]));
}

// 64 accumulator columns of one row (two 32-column TMEM loads).  All eight
// group minima are formed first (independent FMNMX3 chains), one combined test
// skips the whole block, and only groups that can beat the running k-th best
// take the insertion path.
template <int KTOP>
__device__ __forceinline__ void consume64(const float (&v0)[32], const float (&v1)[32], int col0, TopK<KTOP>& tk) {
  float m[8];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    m[g] = min8(&v0[g * 8]);
    m[4 + g] = min8(&v1[g * 8]);
  }
  const float mall = fmin3(fmin3(m[0], m[1], m[2]), fmin3(m[3], m[4], m[5]), fminf(m[6], m[7]));
  if (mall < tk.thr()) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (m[g] < tk.thr()) {
        const float* w = (g < 4) ? &v0[g * 8] : &v1[(g - 4) * 8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (w[j] < tk.thr()) tk.insert(w[j], col0 + g * 8 + j);
        }
      }
    }
  }
}

// NPART = number of column parts an accumulator tile is split into among
// epilogue warps (1: 8 epilogue warps, 2: 16 epilogue warps = 4 per SM sub-partition).
template <Kind kKind, int KTOP, int NPART>
__global__ void __launch_bounds__(128 + 256 * NPART, 1)
knn_umma_kernel(const ImgDev* __restrict__ imgs, const KnnUnit* __restrict__ units, int n_units,
                int* __restrict__ out_idx, float* __restrict__ out_d2, int dbg_flags) {
  constexpr int kEpiWarps = 8 * NPART;
  constexpr int kCols = 128 / NPART;  // accumulator columns per epilogue warp
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kSmemA;
  Barriers* bars = reinterpret_cast<Barriers*>(smem + kSmemA + kSmemB);
  float2* merge = reinterpret_cast<float2*>(smem + kSmemA + kSmemB + ((sizeof(Barriers) + 15) / 16) * 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 1 && elect_one()) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->a_full[i], 1);
      mbar_init(&bars->a_empty[i], 1);
    }
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(&bars->b_full[i], 1);
      mbar_init(&bars->b_empty[i], 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&bars->t_full[i], 1);
      mbar_init(&bars->t_empty[i], kEpiWarps);
    }
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc<kTmemCols>(&bars->tmem_base);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------ B-tile producer
    if (elect_one()) {
      uint32_t it = 0;  // running B-tile counter across units
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const KnnUnit unit = units[u];
        const ImgDev t = imgs[unit.t_slot];
        const int n_tb = (t.n + kTileRows - 1) / kTileRows;
        for (int tb = 0; tb < n_tb; ++tb, ++it) {
          const uint32_t stage = it % kBStages;
          const uint32_t par = (it / kBStages) & 1;
          mbar_wait(&bars->b_empty[stage], par ^ 1, 10);
          mbar_arrive_expect_tx(&bars->b_full[stage], kTileBytes);
          bulk_g2s(smem_b + stage * kTileBytes, t.b_form + static_cast<size_t>(tb) * kTileBytes, kTileBytes,
                   &bars->b_full[stage]);
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------ A-tile producer
    if (elect_one()) {
      uint32_t it = 0;  // unit counter of this CTA
      for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
        const KnnUnit unit = units[u];
        const ImgDev q = imgs[unit.q_slot];
        const uint8_t* src = q.a_form + static_cast<size_t>(unit.super) * kSuperRows * kRowBytes;
        const uint32_t par = it & 1;
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&bars->a_empty[h], par ^ 1, 20 + h);
          mbar_arrive_expect_tx(&bars->a_full[h], kTileBytes);
          bulk_g2s(smem_a + h * kTileBytes, src + static_cast<size_t>(h) * kTileBytes, kTileBytes, &bars->a_full[h]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, 128, 0, 0);
      const uint32_t a_addr = smem_u32(smem_a);
      const uint32_t b_addr = smem_u32(smem_b);
      uint32_t it = 0, uit = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++uit) {
        const KnnUnit unit = units[u];
        const ImgDev t = imgs[unit.t_slot];
        const int n_tb = (t.n + kTileRows - 1) / kTileRows;
        for (int tb = 0; tb < n_tb; ++tb, ++it) {
          const uint32_t stage = it % kBStages;
          const uint32_t par = (it / kBStages) & 1;
          const uint32_t acc = it % kAccStages;
          const uint32_t apar = (it / kAccStages) & 1;
          mbar_wait(&bars->b_full[stage], par, 30);
          mbar_wait(&bars->t_empty[acc], apar ^ 1, 31);
          if (tb == 0) {
            mbar_wait(&bars->a_full[0], uit & 1, 32);
            mbar_wait(&bars->a_full[1], uit & 1, 33);
          }
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t taddr = tmem_base + acc * 256 + h * 128;
#pragma unroll
            for (int ks = 0; ks < kKSteps; ++ks) {
              const uint64_t adesc = make_smem_desc(a_addr + h * kTileBytes + ks * kKStepBytes, kLBO, kSBO);
              const uint64_t bdesc = make_smem_desc(b_addr + stage * kTileBytes + ks * kKStepBytes, kLBO, kSBO);
              umma<kKind>(taddr, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
            }
            if (tb == n_tb - 1) umma_commit(&bars->a_empty[h]);  // A half free for the next unit
          }
          umma_commit(&bars->b_empty[stage]);
          umma_commit(&bars->t_full[acc]);
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue
    const int g = (warp - 4) >> 2;      // 0 .. 2*NPART-1
    const int h = g / NPART;            // which accumulator / A tile
    const int part = g % NPART;         // which column part of it
    const int quad = warp & 3;          // TMEM lane quadrant this warp may touch
    const int row_in_tile = quad * 32 + lane;
    TopK<KTOP> tk;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const KnnUnit unit = units[u];
      const ImgDev q = imgs[unit.q_slot];
      const ImgDev t = imgs[unit.t_slot];
      const int n_tb = (t.n + kTileRows - 1) / kTileRows;
      tk.reset();
      for (int tb = 0; tb < n_tb; ++tb, ++it) {
        const uint32_t acc = it % kAccStages;
        const uint32_t apar = (it / kAccStages) & 1;
        mbar_wait(&bars->t_full[acc], apar, 40);
        tc_fence_after();
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256 + h * 128 + part * kCols;
#pragma unroll 1
        for (int c = 0; c < kCols / 64; ++c) {
          if (dbg_flags == 1) break;
          float v0[32], v1[32];
          __syncwarp();
          tmem_ld32(taddr + c * 64, v0);
          tmem_ld32(taddr + c * 64 + 32, v1);
          tmem_ld_wait(v0);
          tmem_ld_wait(v1);
          if (dbg_flags == 0) {
            consume64<KTOP>(v0, v1, tb * kTileRows + part * kCols + c * 64, tk);
          } else if (dbg_flags == 2) {  // profiling aid: fast path only (results are NOT valid)
            float m = v0[0];
#pragma unroll
            for (int j = 0; j < 4; ++j) m = fminf(m, fminf(min8(&v0[j * 8]), min8(&v1[j * 8])));
            tk.d[0] = fminf(tk.d[0], m);
          }  // dbg_flags == 1: MMA/TMA pipeline only, accumulators dropped
        }
        __syncwarp();
        tc_fence_before();
        if (lane == 0) mbar_arrive(&bars->t_empty[acc]);
      }
      if (NPART == 2) {
        // hand the upper column part's list to the lower part's thread of the same row
        if (part == 1) {
#pragma unroll
          for (int s = 0; s < KTOP; ++s)
            merge[(h * 128 + row_in_tile) * KTOP + s] = make_float2(tk.d[s], __int_as_float(tk.i[s]));
        }
        asm volatile("bar.sync 1, %0;" ::"n"(256 * NPART) : "memory");
        if (part == 0) {
#pragma unroll
          for (int s = 0; s < KTOP; ++s) {
            const float2 e = merge[(h * 128 + row_in_tile) * KTOP + s];
            tk.insert_lex(e.x, __float_as_int(e.y));
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(256 * NPART) : "memory");
      }
      const int row = unit.super * kSuperRows + h * kTileRows + row_in_tile;
      if (part == 0 && row < q.n) {
        const size_t o = (static_cast<size_t>(unit.out_base) + row) * KTOP;
#pragma unroll
        for (int s = 0; s < KTOP; ++s) {
          out_idx[o + s] = tk.i[s] == 0x7fffffff ? -1 : tk.i[s];
          out_d2[o + s] = tk.d[s];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// Debug aid: one 128x128 distance tile (first A tile of `q` x first B tile of `t`)
// with the descriptor strides given at run time, accumulators dumped to global.
template <Kind kKind>
__global__ void __launch_bounds__(128, 1)
umma_tile_debug_kernel(const uint8_t* __restrict__ a_tile, const uint8_t* __restrict__ b_tile, uint32_t lbo,
                       uint32_t sbo, uint32_t kstep_bytes, int ksteps, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full, bar_done;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<128>(&s_tmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_full, 2 * kTileBytes);
    bulk_g2s(smem, a_tile, kTileBytes, &bar_full);
    bulk_g2s(smem + kTileBytes, b_tile, kTileBytes, &bar_full);
    mbar_wait(&bar_full, 0, 90);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc(128, 128, 0, 0);
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint64_t ad = make_smem_desc(smem_u32(smem) + ks * kstep_bytes, lbo, sbo);
      const uint64_t bd = make_smem_desc(smem_u32(smem + kTileBytes) + ks * kstep_bytes, lbo, sbo);
      umma<kKind>(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(&bar_done);
  }
  mbar_wait(&bar_done, 0, 91);
  tc_fence_after();
  for (int c = 0; c < 4; ++c) {
    float v[32];
    __syncwarp();
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait(v);
#pragma unroll
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 128 + c * 32 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<128>(tmem);
  }
}

template <Kind kKind, int KTOP, int NPART>
cudaError_t launch_p(const ImgDev* imgs, const KnnUnit* units, int n_units, int* out_idx, float* out_d2, int num_sms,
                     cudaStream_t stream) {
  auto kern = knn_umma_kernel<kKind, KTOP, NPART>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal);
  if (err != cudaSuccess) return err;
  const int grid = n_units < num_sms ? n_units : num_sms;
  static const int flags = [] {
    const char* e = getenv("IAM_UMMA_DEBUG");  // profiling aid, results invalid when set: 1 = no epilogue, 2 = fast path only
    return e ? atoi(e) : 0;
  }();
  kern<<<grid, 128 + 256 * NPART, kSmemTotal, stream>>>(imgs, units, n_units, out_idx, out_d2, flags);
  return cudaGetLastError();
}

template <Kind kKind, int KTOP>
cudaError_t launch_t(const ImgDev* imgs, const KnnUnit* units, int n_units, int* out_idx, float* out_d2, int num_sms,
                     cudaStream_t stream) {
  static const int parts = [] {
    const char* e = getenv("IAM_UMMA_PARTS");  // tuning / A-B aid: 1 = 8 epilogue warps, 2 = 16 (default)
    return (e && atoi(e) == 1) ? 1 : 2;
  }();
  return parts == 1 ? launch_p<kKind, KTOP, 1>(imgs, units, n_units, out_idx, out_d2, num_sms, stream)
                    : launch_p<kKind, KTOP, 2>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
}

}  // namespace

cudaError_t launch_umma_tile_debug(int norm, const uint8_t* a_tile, const uint8_t* b_tile, uint32_t lbo, uint32_t sbo,
                                   uint32_t kstep_bytes, int ksteps, float* out, cudaStream_t stream) {
  const size_t smem = 2 * kTileBytes + 1024;
  if (norm == 0) {
    cudaError_t e = cudaFuncSetAttribute(umma_tile_debug_kernel<Kind::F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    umma_tile_debug_kernel<Kind::F16><<<1, 128, smem, stream>>>(a_tile, b_tile, lbo, sbo, kstep_bytes, ksteps, out);
  } else {
    cudaError_t e = cudaFuncSetAttribute(umma_tile_debug_kernel<Kind::F8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    umma_tile_debug_kernel<Kind::F8><<<1, 128, smem, stream>>>(a_tile, b_tile, lbo, sbo, kstep_bytes, ksteps, out);
  }
  return cudaGetLastError();
}

cudaError_t launch_knn_umma(int norm, int k, const ImgDev* imgs, const KnnUnit* units, int n_units, int* out_idx,
                            float* out_d2, int num_sms, cudaStream_t stream) {
  if (n_units <= 0) return cudaSuccess;
  if (norm == 0) {
    if (k == 1) return launch_t<Kind::F16, 1>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
    if (k == 2) return launch_t<Kind::F16, 2>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
    if (k == 3) return launch_t<Kind::F16, 3>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
  } else {
    if (k == 1) return launch_t<Kind::F8, 1>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
    if (k == 2) return launch_t<Kind::F8, 2>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
    if (k == 3) return launch_t<Kind::F8, 3>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace iam
