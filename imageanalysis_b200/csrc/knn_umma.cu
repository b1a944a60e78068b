// knn_umma.cu — fused all-pairs descriptor distance + top-k on the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM), replacing
// cv2.DescriptorMatcher.knnMatch as called from raw_matches()
// (reference scripts/lib/matcher.py:203-216).
//
// One work unit = 256 query descriptors (two 128-row A tiles, copied once per
// unit from a shared-memory staging buffer into TENSOR MEMORY with tcgen05.cp,
// so the MMAs read only the B operand from shared memory) against every
// descriptor of the train image (96-row B tiles streamed by cp.async.bulk
// through a 4-deep mbarrier ring; in a 2-CTA cluster each CTA fetches half of
// every tile and multicasts it to both).  The augmented K-step makes every
// accumulator element the exact squared L2 distance (integer valued < 2^23,
// exact in fp32) or the exact Hamming distance, so the N x M distance matrix
// never leaves the SM: 24 epilogue warps read it from TMEM and keep a per-row
// running top-k in registers.
//
// Warp roles (896 threads, 1 CTA / SM, persistent over units):
//   warp 0      : B-tile producer  (bulk copy  -> b_full[stage])
//   warp 1      : MMA issuer       (warp-uniform loop, one elected lane issues: 18 tcgen05.cp per unit,
//                 18 TS-form tcgen05.mma M128xN96xK16 per B tile into a ring of three 96-column accumulator slots)
//   warp 2      : TMEM allocator / deallocator
//   warp 3      : A-tile producer  (bulk copy  -> a_full[tile])
//   warps 4..27 : epilogue (6 per SM sub-partition): warp -> (A tile, 32-column part of each
//                 96-column accumulator tile, TMEM lane quadrant = warp_id % 4); the three threads
//                 that share a row exchange their running bounds through shared memory every
//                 tile (stale bounds are still valid) and merge their lists once per unit
#include <cuda_runtime.h>

#include <cstdlib>

#include "knn.h"
#include "layout.h"
#include "ptx.cuh"

namespace iam {

namespace {

constexpr int kBStages = 4;
#ifndef IAM_SHARE_EVERY
#define IAM_SHARE_EVERY 1
#endif
constexpr int kShareEvery = IAM_SHARE_EVERY;                   // tiles between exchanges of running bounds (power of two)
constexpr int kParts = kBRows / 32;              // 32-column parts of a B tile, one epilogue warp each (per A tile and lane quadrant)
constexpr int kWarpsPerATile = 4 * kParts;
constexpr int kEpiWarps = kATiles * kWarpsPerATile;  // 24 = 6 per SM sub-partition
constexpr int kThreads = 128 + 32 * kEpiWarps;   // 896
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemAColsPerTile = kKSteps * 8;  // 72 columns: 128 rows x 144 fp16 (two K elements per 32-bit cell)
// Accumulator slots of kBRows columns each, handed round-robin to successive (B tile, A tile) products.
constexpr int kSlots = (512 - kATiles * static_cast<int>(kTmemAColsPerTile)) / kBRows;  // 3 for 96-column tiles, 5 for 64
constexpr uint32_t kTmemA = kSlots * kBRows;     // A operand region behind the accumulator slots

struct __align__(8) Barriers {
  uint64_t a_full[kATiles];
  uint64_t a_empty[kATiles];
  uint64_t b_full[kBStages];
  uint64_t b_empty[kBStages];
  uint64_t t_full[kSlots];
  uint64_t t_empty[kSlots];
  uint32_t tmem_base;
  uint32_t pad;
};

constexpr size_t kSmemA = kATiles * kTileBytes;         //  73728: staging for the next unit's query tiles
constexpr size_t kSmemB = kBStages * kBTileBytes;       // 110592: streamed train tiles
constexpr size_t kSmemBars = ((sizeof(Barriers) + 127) / 128) * 128;
constexpr size_t kSmemShare = 2 * kParts * kSuperRows * 4;   //  6144: running k-th bests exchanged between the column parts of a row
constexpr size_t kSmemMerge = (kParts > 1 ? kParts - 1 : 1) * kSuperRows * 3 * 8;  // end-of-unit hand-over of the other parts' lists
constexpr size_t kSmemTotal = kSmemA + kSmemB + kSmemBars + kSmemShare + kSmemMerge + 128;
static_assert(kSmemTotal <= 232448, "shared memory budget");
static_assert(kTmemA + kATiles * kTmemAColsPerTile <= 512, "tensor memory budget");

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
  // Branch-free insertion network; a no-op when x >= thr().  A thread sees its
  // columns in ascending order and every comparison is strict, so among equal
  // distances the earliest (lowest) column stays first: the order
  // cv2.BFMatcher reports ties in.
  __device__ __forceinline__ void insert(float x, int col) {
    if constexpr (KTOP == 2) {
      // Same network as below, written with predicated moves: two compares on the ALU pipe, the six
      // moves can issue as IMAD.MOV on the FMA pipe, which the (ALU-pipe-bound) epilogue leaves idle.
      asm("{\n\t.reg .pred p0, p1;\n\t"
          "setp.lt.f32 p0, %4, %0;\n\t"
          "setp.lt.and.f32 p1, %4, %1, !p0;\n\t"
          "@p1 mov.f32 %1, %4;\n\t"
          "@p1 mov.b32 %3, %5;\n\t"
          "@p0 mov.f32 %1, %0;\n\t"
          "@p0 mov.b32 %3, %2;\n\t"
          "@p0 mov.f32 %0, %4;\n\t"
          "@p0 mov.b32 %2, %5;\n\t}"
          : "+f"(d[0]), "+f"(d[1]), "+r"(i[0]), "+r"(i[1])
          : "f"(x), "r"(col));
      return;
    }
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
  // Order-independent insert ((distance, index) lexicographic): merges the list of
  // the other column half at the end of a unit.
  __device__ __forceinline__ void insert_lex(float x, int col) {
#pragma unroll
    for (int s = KTOP - 1; s >= 0; --s) {
      const int sp = s > 0 ? s - 1 : 0;
      const bool lt_prev = (s > 0) ? (x < d[sp] || (x == d[sp] && col < i[sp])) : false;
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

__device__ __forceinline__ float next_up(float x) {  // x >= 0
  return __int_as_float(__float_as_int(x) + 1);
}

// 32 accumulator columns of one row.  Fast path: one FMNMX3-based minimum per
// group of 4 columns and a warp vote; the insertion network runs (for the whole
// warp, uniformly: no divergence) only for columns where some lane can beat its
// bound.  `pb_up` is the smallest value NOT admissible according to the partner
// thread that owns the other 32 columns of this row (ties with the partner's
// bound are admitted; the final merge orders them by index).  Expected
// insertions per row over M columns are ~k*ln(M/k), so almost every group takes
// the 5-instruction fast path.
// The running lists carry an ENCODED column index e = tp * 33 + j, where tp = tile * kParts + part numbers
// the 32-column slices of the train image and j < 32 is the column inside the slice.  e is monotone in the real
// column (tp * 32 + j), so ties order identically, and it is formed by ONE IMAD with an immediate addend: the
// epilogue is ALU-pipe bound while the FMA pipe idles, and "base + j" would cost an ALU-pipe add per insert.
constexpr int kEncMul = 33;
template <int J>
__device__ __forceinline__ int enc_index(int tp) {
  int c;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(c) : "r"(tp), "n"(kEncMul), "n"(J));
  return c;
}
__device__ __forceinline__ int dec_index(int e) {
  const int tp = e / kEncMul;
  return tp * 32 + (e - tp * kEncMul);
}

// IAM_VOTE=1: warp-uniform branches decided by votes; 0: plain divergent branches (no VOTE on the ALU pipe)
#ifndef IAM_VOTE
#define IAM_VOTE 1
#endif
__device__ __forceinline__ bool any_lane(bool p) {
#if IAM_VOTE
  return __any_sync(0xffffffffu, p);
#else
  return p;
#endif
}
// "does any lane have x < te": x and te are non-negative floats, so their bit patterns order like integers.
// IAM_REDUX routes the test through one warp-wide integer minimum (REDUX, result in a uniform register, the
// branch is then a uniform-datapath compare) instead of FSETP + VOTE on the ALU pipe.
// level 1: group tests of the fast path; level 2: also the per-element tests of a triggered group.
#ifndef IAM_REDUX
#define IAM_REDUX 0
#endif
template <int kLevel>
__device__ __forceinline__ bool any_below(float x, float te) {
  if (IAM_REDUX >= kLevel) {
    return __reduce_min_sync(0xffffffffu, __float_as_int(x) - __float_as_int(te)) < 0;
  } else {
    return any_lane(x < te);
  }
}

template <int KTOP, int J0, bool kFixed = false>
__device__ __forceinline__ void consume_group(const float* w, int tp, TopK<KTOP>& tk, float te) {
  if (kFixed) {  // profiling aid: every warp does the same slow-path work (4 element tests, 2 insertions)
    const bool f0 = any_below<2>(w[0], 3.0e38f), f1 = any_below<2>(w[1], -1.0f);
    const bool f2 = any_below<2>(w[2], 3.0e38f), f3 = any_below<2>(w[3], -1.0f);
    if (f0) tk.insert(w[0], enc_index<J0>(tp));
    if (f1) tk.insert(w[1], enc_index<J0 + 1>(tp));
    if (f2) tk.insert(w[2], enc_index<J0 + 2>(tp));
    if (f3) tk.insert(w[3], enc_index<J0 + 3>(tp));
    return;
  }
  // four votes issued back to back (computed against the bound at group entry: a superset of what
  // the tightening bound would admit), then the branch-free network only where some lane qualifies
  const bool e0 = any_below<2>(w[0], te);
  const bool e1 = any_below<2>(w[1], te);
  const bool e2 = any_below<2>(w[2], te);
  const bool e3 = any_below<2>(w[3], te);
  if (e0) tk.insert(w[0], enc_index<J0>(tp));
  if (e1) tk.insert(w[1], enc_index<J0 + 1>(tp));
  if (e2) tk.insert(w[2], enc_index<J0 + 2>(tp));
  if (e3) tk.insert(w[3], enc_index<J0 + 3>(tp));
}

// Per 16-column batch the four group tests are formed and voted on up front against the bound at
// entry (it only tightens, so the votes stay conservative): independent FMNMX3/FSETP/VOTE chains
// instead of serialised vote->branch round trips.  (Moving the tests to the idle FMA pipe with
// IMAD/IMAD.HI sign accumulation was measured and is slower: IMAD.HI is not a full-rate instruction.)
template <int KTOP, int H, bool kFixed = false>
__device__ __forceinline__ void consume16(const float* w, int tp, TopK<KTOP>& tk, float pb_up) {
  const float te = kFixed ? -1.0f : fminf(tk.thr(), pb_up);
  const bool t0 = any_below<1>(fminf(fmin3(w[0], w[1], w[2]), w[3]), te);
  const bool t1 = any_below<1>(fminf(fmin3(w[4], w[5], w[6]), w[7]), te);
  const bool t2 = any_below<1>(fminf(fmin3(w[8], w[9], w[10]), w[11]), te);
  const bool t3 = any_below<1>(fminf(fmin3(w[12], w[13], w[14]), w[15]), te);
  if (t0) consume_group<KTOP, H * 16>(w, tp, tk, fminf(tk.thr(), pb_up));
  if (t1 || kFixed) consume_group<KTOP, H * 16 + 4, kFixed>(w + 4, tp, tk, fminf(tk.thr(), pb_up));
  if (t2) consume_group<KTOP, H * 16 + 8>(w + 8, tp, tk, fminf(tk.thr(), pb_up));
  if (t3) consume_group<KTOP, H * 16 + 12>(w + 12, tp, tk, fminf(tk.thr(), pb_up));
}

template <int KTOP, bool kFixed = false>
__device__ __forceinline__ void consume32(const float (&v)[32], int tp, TopK<KTOP>& tk, float pb_up) {
  consume16<KTOP, 0, kFixed>(&v[0], tp, tk, pb_up);
  consume16<KTOP, 1, kFixed>(&v[16], tp, tk, pb_up);
}

template <Kind kKind, int KTOP, bool kATmem, bool kCluster, int kDbg>
__global__ void __launch_bounds__(kThreads, 1)
knn_umma_kernel(const ImgDev* __restrict__ imgs, const KnnUnit* __restrict__ units, int n_units,
                int* __restrict__ out_idx, float* __restrict__ out_d2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kSmemA;
  Barriers* bars = reinterpret_cast<Barriers*>(smem + kSmemA + kSmemB);
  float* share = reinterpret_cast<float*>(smem + kSmemA + kSmemB + kSmemBars);               // [unit parity][part][row]
  float2* merge = reinterpret_cast<float2*>(smem + kSmemA + kSmemB + kSmemBars + kSmemShare);  // [part-1][row][k]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // With kCluster two CTAs (one cluster) walk the unit list in lock step on the two units 2p, 2p+1 of the
  // same directed job: each CTA fetches HALF of every train tile and multicasts it to both, which halves
  // the L2 -> SM traffic of the streamed operand.
  constexpr int kCtas = kCluster ? 2 : 1;
  const int cta_rank = kCluster ? static_cast<int>(cluster_ctarank()) : 0;
  const int first_pu = blockIdx.x / kCtas;
  const int pu_stride = gridDim.x / kCtas;

  for (int i = threadIdx.x; i < 2 * kParts * kSuperRows; i += blockDim.x) share[i] = kInf;
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < kATiles; ++i) {
      mbar_init(&bars->a_full[i], 1);
      mbar_init(&bars->a_empty[i], 1);
    }
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(&bars->b_full[i], 1);
      mbar_init(&bars->b_empty[i], kCtas);  // every CTA that received the tile must be done with it
    }
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&bars->t_full[i], 1);
      mbar_init(&bars->t_empty[i], kWarpsPerATile);  // the warps of whichever A tile last used the slot
    }
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc<kTmemCols>(&bars->tmem_base);
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster) cluster_sync_all();  // peer barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------ B-tile producer
    if (elect_one()) {
      uint32_t it = 0;  // running B-tile counter across units
      for (int pu = first_pu; pu * kCtas < n_units; pu += pu_stride) {
        const int u = pu * kCtas + cta_rank;
        const KnnUnit unit = units[u];
        const ImgDev t = imgs[unit.t_slot];
        const int n_tb = (t.n + kBRows - 1) / kBRows;
        for (int tb = 0; tb < n_tb; ++tb, ++it) {
          const uint32_t stage = it % kBStages;
          const uint32_t par = (it / kBStages) & 1;
          mbar_wait(&bars->b_empty[stage], par ^ 1, 10);
          mbar_arrive_expect_tx(&bars->b_full[stage], kBTileBytes);
          if (kCluster) {
            constexpr uint32_t kHalf = kBTileBytes / 2;  // 32 train rows = 4 core-matrix groups, contiguous
            bulk_g2s_multicast(smem_b + stage * kBTileBytes + cta_rank * kHalf,
                               t.b_form + static_cast<size_t>(tb) * kBTileBytes + cta_rank * kHalf, kHalf,
                               &bars->b_full[stage], 0x3);
          } else {
            bulk_g2s(smem_b + stage * kBTileBytes, t.b_form + static_cast<size_t>(tb) * kBTileBytes, kBTileBytes,
                     &bars->b_full[stage]);
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------ A-tile producer
    if (elect_one()) {
      uint32_t it = 0;  // unit counter of this CTA
      for (int pu = first_pu; pu * kCtas < n_units; pu += pu_stride, ++it) {
        const int u = pu * kCtas + cta_rank;
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
    // The whole warp walks the loops with warp-uniform values (made provably uniform by a shuffle, so the
    // descriptor arithmetic lives in the uniform datapath instead of vector registers + R2UR); one elected lane
    // issues the tcgen05 instructions.  The issue stream of this warp is on the critical path: it shares its
    // scheduler with six ALU-bound epilogue warps.
    constexpr uint32_t idesc = make_idesc(128, kBRows, 0, 0);
    const uint32_t a_addr = smem_u32(smem_a);
    const uint32_t b_addr = smem_u32(smem_b);
    const uint32_t bars_addr = smem_u32(bars);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t tm_a = tm + kTmemA;
    const uint32_t b_lo0 = smem_desc_lo(b_addr, kLBO);
    uint32_t stage = 0, bpar = 0;   // B ring position
    uint32_t slot = 0, tpar = 1;    // accumulator ring position; parity to wait for on t_empty (fresh barrier: 1)
    uint32_t uit = 0;
    for (int pu = first_pu; pu * kCtas < n_units; pu += pu_stride, ++uit) {
      const int u = pu * kCtas + cta_rank;
      const int n_t = __shfl_sync(0xffffffffu, imgs[units[u].t_slot].n, 0);
      const int n_tb = (n_t + kBRows - 1) / kBRows;
      // query tiles: shared-memory staging -> tensor memory (tcgen05.cp), then the staging is free again.
      // tcgen05 operations of one thread execute in issue order, so these copies run after every MMA of
      // the previous unit that still reads the old A tiles.
      for (int a = 0; a < kATiles; ++a) {
        mbar_wait(&bars->a_full[a], uit & 1, 32 + a);
        tc_fence_after();
        if (kATmem) {
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < kKSteps; ++ks)
              tmem_cp_128x256b(tm + kTmemA + a * kTmemAColsPerTile + ks * 8,
                               make_smem_desc(a_addr + a * kTileBytes + ks * kKStepBytes, kLBO, kSBO));
            umma_commit(&bars->a_empty[a]);
          }
          __syncwarp();
        }
      }
      for (int tb = 0; tb < n_tb; ++tb) {
        mbar_wait_a(bars_addr + offsetof(Barriers, b_full) + stage * 8, bpar, 30);
        const uint32_t b_lo = b_lo0 + stage * (kBTileBytes >> 4);
#pragma unroll
        for (int a = 0; a < kATiles; ++a) {
          mbar_wait_a(bars_addr + offsetof(Barriers, t_empty) + slot * 8, tpar, 31);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t taddr = tm + slot * kBRows;
#pragma unroll
            for (int ks = 0; ks < kKSteps; ++ks) {
              const uint64_t bdesc = pack_desc(b_lo + ks * (kKStepBytes >> 4), smem_desc_hi(kSBO));
              if (kATmem) {
                umma_ts<kKind>(taddr, tm_a + a * kTmemAColsPerTile + ks * 8, bdesc, idesc, ks > 0 ? 1u : 0u);
              } else {
                const uint64_t adesc = make_smem_desc(a_addr + a * kTileBytes + ks * kKStepBytes, kLBO, kSBO);
                umma<kKind>(taddr, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
              }
            }
            if (!kATmem && tb == n_tb - 1) umma_commit(&bars->a_empty[a]);
            umma_commit_a(bars_addr + offsetof(Barriers, t_full) + slot * 8);
          }
          __syncwarp();
          if (++slot == kSlots) {
            slot = 0;
            tpar ^= 1;
          }
        }
        if (elect_one()) {
          if (kCluster)
            umma_commit_multicast(&bars->b_empty[stage], 0x3);
          else
            umma_commit(&bars->b_empty[stage]);
        }
        __syncwarp();
        if (++stage == kBStages) {
          stage = 0;
          bpar ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue
    // The per-tile loop is ALU-pipe bound (profiles/): everything loop-invariant lives in pinned registers as
    // ready-made shared-memory / tensor-memory addresses, and the slot / phase of the accumulator ring advance
    // incrementally instead of by division.
    const int e = (warp - 4) >> 2;      // 0 .. kATiles*kParts-1
    const int a = e / kParts;           // which A tile
    const int part = e % kParts;        // which 32 of the B tile's columns
    const int quad = warp & 3;          // TMEM lane quadrant this warp may touch
    const int urow = a * kTileRows + quad * 32 + lane;  // row within the unit
    constexpr uint32_t kPartStride = kSuperRows * 4;            // bytes between the parts' bound slots of one row
    constexpr uint32_t kParityStride = kParts * kPartStride;    // bytes between the two unit-parity buffers
    const uint32_t share_row = smem_u32(share) + urow * 4;
    const uint32_t bar_full0 = pin_reg(smem_u32(&bars->t_full[0]));
    constexpr uint32_t kEmptyOff = kSlots * 8;                  // t_empty[] follows t_full[] in Barriers
    const uint32_t tm_warp = pin_reg(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + part * 32);
    const bool lane0 = lane == 0;
    TopK<KTOP> tk;
    uint32_t slot = a % kSlots, par = 0;  // position in the accumulator ring: sq = it*kATiles + a, slot = sq % kSlots
    static_assert(kATiles <= kSlots, "one wrap per step at most");
    int uit = 0;
    for (int pu = first_pu; pu * kCtas < n_units; pu += pu_stride, ++uit) {
      const int u = pu * kCtas + cta_rank;
      const KnnUnit unit = units[u];
      const ImgDev q = imgs[unit.q_slot];
      const ImgDev t = imgs[unit.t_slot];
      const int n_tb = (t.n + kBRows - 1) / kBRows;
      tk.reset();
      float pb_up = kInf;  // smallest value NOT admissible according to the other column parts of this row
      // Bounds live in a buffer selected by the unit's parity.  At the start of unit u every thread resets its
      // slot in the OTHER buffer (the one unit u+1 will use); the end-of-unit barrier orders that reset before
      // any partner reads it, so a slot only ever holds +inf or values of the unit being processed.
      const uint32_t rd = pin_reg(share_row + (uit & 1) * kParityStride);
      const uint32_t wr = pin_reg(rd + part * kPartStride);
      sts_volatile_f32_a(share_row + ((uit + 1) & 1) * kParityStride + part * kPartStride, kInf);
      const int tp_end = n_tb * kParts + part;
      for (int tp = part; tp < tp_end; tp += kParts) {  // tp numbers the 32-column slices of the train image
        // Bound from the threads that own the column parts of this row (own slot included, it is harmless):
        // nothing worse than the smallest of the k-th bests can end up in the merged list.  Ties are admitted
        // (next_up; the final merge orders them by index); next_up(+inf) is a NaN, which fminf ignores.  Stale
        // values are still valid bounds, so plain volatile shared-memory traffic suffices.
        if (kParts > 1) {
          float g = lds_volatile_f32_a(rd);
#pragma unroll
          for (int pp = 1; pp < kParts; ++pp) g = fminf(g, lds_volatile_f32_a(rd + pp * kPartStride));
          pb_up = fminf(pb_up, __int_as_float(__float_as_int(g) + 1));
        }
        const uint32_t bar = bar_full0 + slot * 8;
        mbar_wait_bare_a(bar, par);
        tc_fence_after();
        if (kDbg != 1) {
          float v[32];
          __syncwarp();
          tmem_ld32(tm_warp + slot * kBRows, v);
          tmem_ld_wait(v);
          // the values are in registers: hand the accumulator slot back to the MMA issuer before consuming them
          __syncwarp();
          tc_fence_before();
          if (lane0) mbar_arrive_a(bar + kEmptyOff);
          if (kDbg == 0) {
            consume32<KTOP>(v, tp, tk, pb_up);
          } else if (kDbg == 5) {  // profiling aid (IAM_UMMA_DEBUG=5): identical slow-path work in every warp and tile
            consume32<KTOP, true>(v, tp, tk, pb_up);
          } else if (kDbg == 4) {  // profiling aid (IAM_UMMA_DEBUG=4): group tests + votes + branches, never taken
            consume32<KTOP>(v, tp, tk, -1.0f);
          } else if (kDbg == 3) {  // profiling aid (IAM_UMMA_DEBUG=3): accumulator read-out only, results NOT valid
            tk.d[0] = fminf(tk.d[0], v[0]);  // the load itself is volatile: all 32 columns are still read
          } else {  // profiling aid (IAM_UMMA_DEBUG=2): fast path only, results NOT valid
            float m = v[0];
#pragma unroll
            for (int j = 0; j < 8; ++j) m = fminf(m, fminf(fmin3(v[j * 4], v[j * 4 + 1], v[j * 4 + 2]), v[j * 4 + 3]));
            tk.d[0] = fminf(tk.d[0], m);
          }
        } else {  // IAM_UMMA_DEBUG=1: MMA/TMA pipeline only, accumulators dropped
          __syncwarp();
          tc_fence_before();
          if (lane0) mbar_arrive_a(bar + kEmptyOff);
        }
        if (kParts > 1) sts_volatile_f32_a(wr, tk.d[KTOP - 1]);
        slot += kATiles;
        if (slot >= kSlots) {
          slot -= kSlots;
          par ^= 1;
        }
      }
      // end of unit: parts 1.. hand their lists to part 0's thread of the same row
      if (part > 0) {
#pragma unroll
        for (int s = 0; s < KTOP; ++s)
          merge[((part - 1) * kSuperRows + urow) * KTOP + s] = make_float2(tk.d[s], __int_as_float(tk.i[s]));
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      const int row = unit.super * kSuperRows + urow;
      if (part == 0) {
#pragma unroll
        for (int pp = 0; pp < kParts - 1; ++pp) {
#pragma unroll
          for (int s = 0; s < KTOP; ++s) {
            const float2 m = merge[(pp * kSuperRows + urow) * KTOP + s];
            tk.insert_lex(m.x, __float_as_int(m.y));
          }
        }
        if (row < q.n) {
          const size_t o = (static_cast<size_t>(unit.out_base) + row) * KTOP;
#pragma unroll
          for (int s = 0; s < KTOP; ++s) {
            out_idx[o + s] = tk.i[s] == 0x7fffffff ? -1 : dec_index(tk.i[s]);
            out_d2[o + s] = tk.d[s];
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");  // merge area is free again
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kCluster) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it
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
                       uint32_t sbo, uint32_t kstep_bytes, int ksteps_in, float* __restrict__ out) {
  const bool a_in_tmem = ksteps_in < 0;  // negative K-step count selects the TS form (A staged through tcgen05.cp)
  const int ksteps = a_in_tmem ? -ksteps_in : ksteps_in;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full, bar_done;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(&s_tmem);
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
    if (a_in_tmem)
      for (int ks = 0; ks < ksteps; ++ks)
        tmem_cp_128x256b(tmem + 128 + ks * 8, make_smem_desc(smem_u32(smem) + ks * kstep_bytes, lbo, sbo));
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint64_t ad = make_smem_desc(smem_u32(smem) + ks * kstep_bytes, lbo, sbo);
      const uint64_t bd = make_smem_desc(smem_u32(smem + kTileBytes) + ks * kstep_bytes, lbo, sbo);
      if (a_in_tmem)
        umma_ts<kKind>(tmem, tmem + 128 + ks * 8, bd, idesc, ks > 0 ? 1u : 0u);
      else
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
    tmem_dealloc<256>(tmem);
  }
}

template <Kind kKind, int KTOP>
cudaError_t launch_t(const ImgDev* imgs, const KnnUnit* units, int n_units, int* out_idx, float* out_d2, int num_sms,
                     cudaStream_t stream) {
  static const bool a_tmem = [] {
    const char* e = getenv("IAM_UMMA_A_SMEM");  // A/B aid: 1 = keep the A operand in shared memory (SS form)
    return !(e && atoi(e) == 1);
  }();
  static const bool cluster = [] {
    const char* e = getenv("IAM_UMMA_NO_CLUSTER");  // A/B aid: 1 = independent CTAs, no multicast of the train tiles
    return !(e && atoi(e) == 1);
  }();
  static const int flags = [] {
    const char* e = getenv("IAM_UMMA_DEBUG");  // profiling aid, results invalid when set: 1 = no epilogue, 2 = fast path only, 3 = accumulator read-out only, 4 = group tests never taken, 5 = fixed slow-path work
    return e ? atoi(e) : 0;
  }();
  using KernT = void (*)(const ImgDev*, const KnnUnit*, int, int*, float*);
  const bool use_cluster = cluster && (n_units % 2 == 0);
  KernT kern;
  if (flags != 0) {  // profiling variants exist for the production configuration only
    if (!use_cluster || !a_tmem || flags < 0 || flags > 5) return cudaErrorInvalidValue;
    kern = flags == 1   ? knn_umma_kernel<kKind, KTOP, true, true, 1>
           : flags == 2 ? knn_umma_kernel<kKind, KTOP, true, true, 2>
           : flags == 3 ? knn_umma_kernel<kKind, KTOP, true, true, 3>
           : flags == 4 ? knn_umma_kernel<kKind, KTOP, true, true, 4>
                        : knn_umma_kernel<kKind, KTOP, true, true, 5>;
  } else if (use_cluster) {
    kern = a_tmem ? knn_umma_kernel<kKind, KTOP, true, true, 0> : knn_umma_kernel<kKind, KTOP, false, true, 0>;
  } else {
    kern = a_tmem ? knn_umma_kernel<kKind, KTOP, true, false, 0> : knn_umma_kernel<kKind, KTOP, false, false, 0>;
  }
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal);
  if (err != cudaSuccess) return err;
  int grid = n_units < num_sms ? n_units : num_sms;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (use_cluster) {
    grid &= ~1;
    if (grid < 2) grid = 2;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  cfg.gridDim = dim3(grid);
  return cudaLaunchKernelEx(&cfg, kern, imgs, units, n_units, out_idx, out_d2);
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
