// knn_umma.cu — fused all-pairs descriptor distance + top-k on the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM), replacing
// cv2.DescriptorMatcher.knnMatch as called from raw_matches()
// (reference scripts/lib/matcher.py:203-216).
//
// One work unit = 512 query descriptors (four 128-row A tiles, resident in
// shared memory) against every descriptor of the train image (64-row B tiles
// streamed by cp.async.bulk through a 4-deep mbarrier ring).  The augmented
// K-step makes every accumulator element the exact squared L2 distance
// (integer valued < 2^23, exact in fp32) or the exact Hamming distance, so
// the N x M distance matrix never leaves the SM: eight epilogue warps read
// it from TMEM and keep a per-row running top-k in registers.
//
// Warp roles (640 threads, 1 CTA / SM, persistent over units):
//   warp 0      : B-tile producer  (bulk copy  -> b_full[stage])
//   warp 1      : MMA issuer       (one elected lane issues 36 tcgen05.mma M128xN64xK16 / B tile)
//   warp 2      : TMEM allocator / deallocator
//   warp 3      : A-tile producer  (bulk copy  -> a_full[half])
//   warps 4..19 : epilogue (4 per SM sub-partition): warp -> (A tile, TMEM lane quadrant
//                 = warp_id % 4); each thread owns ONE query row for all train columns, so
//                 its running top-k needs no cross-thread merge
#include <cuda_runtime.h>

#include <cstdlib>

#include "knn.h"
#include "layout.h"
#include "ptx.cuh"

namespace iam {

namespace {

constexpr int kBStages = 4;
constexpr int kAccStages = 2;
constexpr int kEpiWarps = 4 * kATiles;           // one warp per (A tile, TMEM lane quadrant)
constexpr int kThreads = 128 + 32 * kEpiWarps;   // 640
constexpr uint32_t kTmemCols = 512;              // 2 stages x 4 accumulators x 64 fp32 columns
constexpr uint32_t kAccCols = kATiles * kBRows;  // 256 TMEM columns per accumulator stage

struct __align__(8) Barriers {
  uint64_t a_full[kATiles];
  uint64_t a_empty[kATiles];
  uint64_t b_full[kBStages];
  uint64_t b_empty[kBStages];
  uint64_t t_full[kAccStages];
  uint64_t t_empty[kAccStages];
  uint32_t tmem_base;
  uint32_t pad;
};

constexpr size_t kSmemA = kATiles * kTileBytes;       // 147456: four resident query tiles
constexpr size_t kSmemB = kBStages * kBTileBytes;     //  73728: streamed train tiles
constexpr size_t kSmemTotal = kSmemA + kSmemB + sizeof(Barriers) + 128;

constexpr float kInf = 3.0e38f;

template <int KTOP>
struct TopK {
  float d[KTOP];
  int i[KTOP];
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int s = 0; s < KTOP; ++s) {
      d[s] = kInf;
      i[s] = -1;
    }
  }
  __device__ __forceinline__ float thr() const { return d[KTOP - 1]; }
  // Branch-free insertion network; a no-op when x >= thr().  Columns arrive in
  // ascending order and every comparison is strict, so among equal distances the
  // earliest (lowest) column stays first: the order cv2.BFMatcher reports ties in.
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
};

// 64 accumulator columns of one row.  Fast path: one FMNMX3-based minimum per
// group of 4 columns and a warp vote; the insertion network runs (for the whole
// warp, uniformly: no divergence) only for columns where some lane can beat its
// running k-th best.  Expected insertions per row over M columns are ~k*ln(M/k),
// so almost every group takes the 4-instruction fast path.
template <int KTOP>
__device__ __forceinline__ void consume64(const float (&v0)[32], const float (&v1)[32], int col0, TopK<KTOP>& tk) {
#pragma unroll
  for (int g = 0; g < 16; ++g) {
    const float* w = (g < 8) ? &v0[g * 4] : &v1[(g - 8) * 4];
    const float m = fminf(fmin3(w[0], w[1], w[2]), w[3]);
    if (__any_sync(0xffffffffu, m < tk.thr())) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (__any_sync(0xffffffffu, w[j] < tk.thr())) tk.insert(w[j], col0 + g * 4 + j);
      }
    }
  }
}

template <Kind kKind, int KTOP>
__global__ void __launch_bounds__(kThreads, 1)
knn_umma_kernel(const ImgDev* __restrict__ imgs, const KnnUnit* __restrict__ units, int n_units,
                int* __restrict__ out_idx, float* __restrict__ out_d2, int dbg_flags) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kSmemA;
  Barriers* bars = reinterpret_cast<Barriers*>(smem + kSmemA + kSmemB);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 1 && elect_one()) {
    for (int i = 0; i < kATiles; ++i) {
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
        const int n_tb = (t.n + kBRows - 1) / kBRows;
        for (int tb = 0; tb < n_tb; ++tb, ++it) {
          const uint32_t stage = it % kBStages;
          const uint32_t par = (it / kBStages) & 1;
          mbar_wait(&bars->b_empty[stage], par ^ 1, 10);
          mbar_arrive_expect_tx(&bars->b_full[stage], kBTileBytes);
          bulk_g2s(smem_b + stage * kBTileBytes, t.b_form + static_cast<size_t>(tb) * kBTileBytes, kBTileBytes,
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
        for (int a = 0; a < kATiles; ++a) {
          mbar_wait(&bars->a_empty[a], par ^ 1, 20 + a);
          mbar_arrive_expect_tx(&bars->a_full[a], kTileBytes);
          bulk_g2s(smem_a + a * kTileBytes, src + static_cast<size_t>(a) * kTileBytes, kTileBytes, &bars->a_full[a]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, kBRows, 0, 0);
      const uint32_t a_addr = smem_u32(smem_a);
      const uint32_t b_addr = smem_u32(smem_b);
      uint32_t it = 0, uit = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++uit) {
        const KnnUnit unit = units[u];
        const ImgDev t = imgs[unit.t_slot];
        const int n_tb = (t.n + kBRows - 1) / kBRows;
        for (int tb = 0; tb < n_tb; ++tb, ++it) {
          const uint32_t stage = it % kBStages;
          const uint32_t par = (it / kBStages) & 1;
          const uint32_t acc = it % kAccStages;
          const uint32_t apar = (it / kAccStages) & 1;
          mbar_wait(&bars->b_full[stage], par, 30);
          mbar_wait(&bars->t_empty[acc], apar ^ 1, 31);
          tc_fence_after();
#pragma unroll
          for (int a = 0; a < kATiles; ++a) {
            if (tb == 0) {
              mbar_wait(&bars->a_full[a], uit & 1, 32 + a);
              tc_fence_after();
            }
            const uint32_t taddr = tmem_base + acc * kAccCols + a * kBRows;
#pragma unroll
            for (int ks = 0; ks < kKSteps; ++ks) {
              const uint64_t adesc = make_smem_desc(a_addr + a * kTileBytes + ks * kKStepBytes, kLBO, kSBO);
              const uint64_t bdesc = make_smem_desc(b_addr + stage * kBTileBytes + ks * kKStepBytes, kLBO, kSBO);
              umma<kKind>(taddr, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
            }
            if (tb == n_tb - 1) umma_commit(&bars->a_empty[a]);  // this A tile is free for the next unit
          }
          umma_commit(&bars->b_empty[stage]);
          umma_commit(&bars->t_full[acc]);
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue: one row per thread, all columns
    const int a = (warp - 4) >> 2;      // which accumulator / A tile
    const int quad = warp & 3;          // TMEM lane quadrant this warp may touch
    const int row_in_tile = quad * 32 + lane;
    TopK<KTOP> tk;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const KnnUnit unit = units[u];
      const ImgDev q = imgs[unit.q_slot];
      const ImgDev t = imgs[unit.t_slot];
      const int n_tb = (t.n + kBRows - 1) / kBRows;
      tk.reset();
      for (int tb = 0; tb < n_tb; ++tb, ++it) {
        const uint32_t acc = it % kAccStages;
        const uint32_t apar = (it / kAccStages) & 1;
        mbar_wait(&bars->t_full[acc], apar, 40);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccCols + a * kBRows;
        if (dbg_flags != 1) {
          float v0[32], v1[32];
          __syncwarp();
          tmem_ld32(taddr, v0);
          tmem_ld32(taddr + 32, v1);
          tmem_ld_wait(v0);
          tmem_ld_wait(v1);
          if (dbg_flags == 0) {
            consume64<KTOP>(v0, v1, tb * kBRows, tk);
          } else {  // profiling aid (IAM_UMMA_DEBUG=2): fast path only, results NOT valid
            float m = v0[0];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              m = fminf(m, fmin3(fminf(v0[j * 4], v0[j * 4 + 1]), fminf(v0[j * 4 + 2], v0[j * 4 + 3]),
                                 fmin3(fminf(v1[j * 4], v1[j * 4 + 1]), v1[j * 4 + 2], v1[j * 4 + 3])));
            tk.d[0] = fminf(tk.d[0], m);
          }
        }  // IAM_UMMA_DEBUG=1: MMA/TMA pipeline only, accumulators dropped
        __syncwarp();
        tc_fence_before();
        if (lane == 0) mbar_arrive(&bars->t_empty[acc]);
      }
      const int row = unit.super * kSuperRows + a * kTileRows + row_in_tile;
      if (row < q.n) {
        const size_t o = (static_cast<size_t>(unit.out_base) + row) * KTOP;
#pragma unroll
        for (int s = 0; s < KTOP; ++s) {
          out_idx[o + s] = tk.i[s];
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

template <Kind kKind, int KTOP>
cudaError_t launch_t(const ImgDev* imgs, const KnnUnit* units, int n_units, int* out_idx, float* out_d2, int num_sms,
                     cudaStream_t stream) {
  auto kern = knn_umma_kernel<kKind, KTOP>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal);
  if (err != cudaSuccess) return err;
  static const int flags = [] {
    const char* e = getenv("IAM_UMMA_DEBUG");  // profiling aid, results invalid when set: 1 = no epilogue, 2 = fast path only
    return e ? atoi(e) : 0;
  }();
  const int grid = n_units < num_sms ? n_units : num_sms;
  kern<<<grid, kThreads, kSmemTotal, stream>>>(imgs, units, n_units, out_idx, out_d2, flags);
  return cudaGetLastError();
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
