// knn_umma.cu — fused all-pairs descriptor distance + top-k on the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM), replacing
// cv2.DescriptorMatcher.knnMatch as called from raw_matches()
// (reference scripts/lib/matcher.py:203-216).
//
// One work unit = 256 query descriptors (two 128-row A tiles, copied once per
// unit from a shared-memory staging buffer into TENSOR MEMORY with tcgen05.cp,
// so the MMAs read only the B operand from shared memory) against every
// descriptor of the train image (96-row B tiles streamed by cp.async.bulk
// through an mbarrier ring; in a 2-CTA cluster each CTA fetches half of
// every tile and multicasts it to both).  The augmented K-step makes every
// accumulator element an exact function of the squared L2 distance (byte layout:
// s32, larger = nearer; wide layouts: the distance itself in fp32) or the exact
// Hamming distance, so the N x M distance matrix never leaves the SM: 24 epilogue
// warps read it from TMEM and keep a per-row running top-k in registers.
//
// Warp roles (896 threads, 1 CTA / SM, persistent over units; the role warps carry the highest warp ids):
//   B-tile producer  (bulk copy -> b_full[stage])
//   MMA issuer(s)    (warp-uniform loop, one elected lane issues: per unit kKSteps tcgen05.cp per query tile,
//                     per B tile kKSteps TS-form tcgen05.mma M128 x N96 into the accumulator slots; byte layout:
//                     one issuer per query tile, each with its own pair of slots)
//   TMEM allocator   (second MMA issuer in the byte layout)
//   A-tile producer  (bulk copy -> a_full[tile])
//   warps 0..23      epilogue (6 per SM sub-partition): warp -> (A tile, 32-column part of each 96-column
//                    accumulator tile, TMEM lane quadrant = warp_id % 4); the three threads that share a row
//                    exchange their running bounds through shared memory (stale bounds are still valid) and
//                    merge their lists once per unit.  Byte layout, k = 2: packed-key top-2 (consume32_packed);
//                    other kinds / k: group extrema + warp votes + insertion network (consume32)
#include <cuda_runtime.h>

#include <cstdlib>
#include <type_traits>
#include <utility>

#include "knn.h"
#include "layout.h"
#include "ptx.cuh"

namespace iam {

namespace {

// Build-time switch of the warp layout (A/B aid):
//   IAM_ROLES_LAST   1: the four role warps (TMA, MMA issue, TMEM alloc) carry the HIGHEST warp ids of the CTA -- the SM
//                       sub-partition arbiter serves high warp ids first, and the MMA issuer must never queue
//                       behind the ALU-bound epilogue warps it shares a scheduler with
#ifndef IAM_ROLES_LAST
#define IAM_ROLES_LAST 1
#endif
//   IAM_DUAL_ISSUE   1: one MMA-issuing warp PER QUERY TILE (the TMEM allocator warp becomes the second issuer) when the
//                       accumulator slots split evenly between the tiles: each tile's products then advance on their own,
//                       so a slow epilogue warp of one tile no longer stalls the other tile's twelve warps, and the
//                       per-product issue latency (barrier waits, descriptor set-up) is paid by two warps in parallel
#ifndef IAM_GROUP_UNCOND
#define IAM_GROUP_UNCOND 0
#endif
#ifndef IAM_DUAL_ISSUE
#define IAM_DUAL_ISSUE 1
#endif
// Column parts of a B tile, one epilogue warp each (per A tile and lane quadrant).  IAM_PART_COLS = columns a thread
// holds per slice in the packed-key path (byte layout, k = 2): 32 (three parts, 24 epilogue warps), 48 (two parts, 16
// warps, 96 registers) or 96 (one part, 8 warps, no exchange of bounds between threads at all).  Wider slices spread
// the per-slice overhead (barrier wait, tensor-memory load, exchange, vote) over more columns.
#ifndef IAM_PART_COLS
#define IAM_PART_COLS 48         // measured (1990 pairs): 32 -> 15.86 ms, 48 -> 15.28 ms, 96 -> 16.5 ms
#endif

// Per-kind, per-shape kernel configuration: operand layout (layout.h) + tensor-memory / shared-memory budget.
// kT = query tiles (128 rows each) resident per CTA:
//   2: one CTA per SM owns a whole 256-row work unit (896 threads, all 512 tensor-memory columns);
//   1: a CTA owns HALF a unit (512 threads, 256 columns) and two CTAs share an SM.  The insertion work of a unit is
//      front-loaded (every row starts with empty lists: the first eight of ~53 train tiles carry 60 % of all
//      insertions), so one CTA alternates between an ALU-bound phase with an idle tensor pipe and an MMA-bound phase
//      with idle ALUs; two co-resident CTAs drift out of phase and fill each other's gaps.
template <Kind kKind, int kT, int kPC = 32>
struct Cfg : LayD<kKind> {
  using L = LayD<kKind>;
  static constexpr int kPartCols = kPC;
  static constexpr int kParts = kBRows / kPC;
  static_assert(kParts * kPC == kBRows && (kPC == 32 || kPC == 48 || kPC == 96), "part columns");
  static constexpr int kWarpsPerATile = 4 * kParts;
  static constexpr int kKeyMul = kPC <= 32 ? 32 : kPC <= 64 ? 64 : 128;   // packed key = acc * kKeyMul + (kKeyMul - 1 - column)
  static constexpr int kEpiWarps = kT * kWarpsPerATile;          // 24 (6 per SM sub-partition) / 12
  static constexpr int kThreads = 128 + 32 * kEpiWarps;          // 896 / 512
  static constexpr int kCtasPerSm = kT == 1 ? 2 : 1;
  static constexpr int kRows = kT * kTileRows;                   // query rows per CTA pass
  static constexpr int kItems = kATiles / kT;                    // CTA passes ("items") per 256-row work unit
  static constexpr uint32_t kTmemCols = kT == 1 ? 256 : 512;
  static constexpr int kBStages = kT == 1 ? 4 : L::kBStages;
  static constexpr uint32_t kTmemAColsPerTile = L::kKSteps * 8;  // 128 rows x 32 bytes per K-step = 8 columns
  // Accumulator slots of kBRows columns each, handed round-robin to successive (B tile, A tile) products:
  // kT = 2: 3 with the wide layout (144 columns of query operand), 4 with the byte layout (80 columns).
  static constexpr int kSlots = (static_cast<int>(kTmemCols) - kT * static_cast<int>(kTmemAColsPerTile)) / kBRows;
  static constexpr uint32_t kTmemA = kSlots * kBRows;            // A operand region behind the accumulator slots
  static constexpr bool kDual = IAM_DUAL_ISSUE && kT == 2 && kSlots % kT == 0;
  struct __align__(8) Barriers {
    uint64_t a_full[kT];
    uint64_t a_empty[kT];
    uint64_t b_full[kBStages];
    uint64_t b_empty[kBStages];
    uint64_t t_full[kSlots * kParts];    // per (slot, part) when the products are split by part (IAM_SPLIT_N), else [kSlots] used
    uint64_t t_empty[kSlots * kParts];
    uint32_t tmem_base;
    uint32_t pad;
  };
  static constexpr size_t kSmemA = kT * L::kTileBytes;            // staging for the next item's query tiles
  static constexpr size_t kSmemB = kBStages * L::kBTileBytes;     // streamed train tiles
  static constexpr size_t kSmemBars = ((sizeof(Barriers) + 127) / 128) * 128;
  static constexpr size_t kSmemShare = 2 * kParts * kRows * 8;    // running bests exchanged between the column parts of a row: 8-byte slots (k-th best; the pair loop of the packed path also publishes the best)
  static constexpr size_t kSmemMerge = (kParts > 1 ? kParts - 1 : 1) * kRows * 3 * 8;  // end-of-item hand-over of the other parts' lists
  static constexpr size_t kSmemTotal = kSmemA + kSmemB + kSmemBars + kSmemShare + kSmemMerge + 128;
  static_assert(kCtasPerSm * (kSmemTotal + 1024) <= 233472, "shared memory budget");
  static_assert(kTmemA + kT * kTmemAColsPerTile <= kTmemCols, "tensor memory budget");
  static_assert(kT <= kSlots, "one wrap per step at most");
  static_assert(kATiles % kT == 0, "whole CTA passes per work unit");
};

constexpr float kInf = 3.0e38f;

// Ordering of accumulator values.  Wide layouts: the accumulator IS the distance (float, smaller = nearer).
// Byte layout: acc = q.t + CAP - floor(||t||^2/2) (int, LARGER = nearer), see layout.h.
template <Kind kKind>
struct Ord {
  using T = float;
  __device__ static constexpr T worst() { return kInf; }
  __device__ static constexpr T never() { return -1.0f; }   // a bound no value beats
  __device__ static __forceinline__ bool better(T x, T y) { return x < y; }
  __device__ static __forceinline__ T best(T a, T b) { return fminf(a, b); }
  __device__ static __forceinline__ T best3(T a, T b, T c) { return fmin3(a, b, c); }
  // first value NOT admissible when a partner's k-th best is g (ties are admitted: the merge orders them by index)
  __device__ static __forceinline__ T loosen(T g) { return __int_as_float(__float_as_int(g) + 1); }
  __device__ static __forceinline__ uint32_t bits(T x) { return __float_as_uint(x); }
  __device__ static __forceinline__ T from_bits(uint32_t b) { return __uint_as_float(b); }
};
template <>
struct Ord<Kind::I8> {
  using T = int;
  __device__ static constexpr T worst() { return -1; }
  __device__ static constexpr T never() { return 0x7fffffff; }
  __device__ static __forceinline__ bool better(T x, T y) { return x > y; }
  __device__ static __forceinline__ T best(T a, T b) { return max(a, b); }
  __device__ static __forceinline__ T best3(T a, T b, T c) { return imax3(a, b, c); }
  __device__ static __forceinline__ T loosen(T g) { return g - 1; }
  __device__ static __forceinline__ uint32_t bits(T x) { return static_cast<uint32_t>(x); }
  __device__ static __forceinline__ T from_bits(uint32_t b) { return static_cast<int>(b); }
};

template <int KTOP, typename O>
struct TopK {
  using T = typename O::T;
  T d[KTOP];
  int i[KTOP];
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int s = 0; s < KTOP; ++s) {
      d[s] = O::worst();
      i[s] = 0x7fffffff;
    }
  }
  __device__ __forceinline__ T thr() const { return d[KTOP - 1]; }
  // Branch-free insertion network; a no-op when x is not better than thr().  A thread sees its
  // columns in ascending order and every comparison is strict, so among equal
  // values the earliest (lowest) column stays first: the order
  // cv2.BFMatcher reports ties in.
  __device__ __forceinline__ void insert(T x, int col) {
    if constexpr (KTOP == 2) {
      // Six ALU-pipe instructions, all guarded by "x beats the second best": the new second best is the worse of
      // (old best, x), the new best the better of them; only the two index moves need the second compare.
      // (A compare-twice / select-six network costs eight; the epilogue is ALU-pipe bound.)
      if constexpr (std::is_same<T, float>::value) {  // float, smaller is better
        asm("{\n\t.reg .pred p0, p1;\n\t"
            "setp.lt.f32 p1, %4, %1;\n\t"
            "@p1 max.f32 %1, %0, %4;\n\t"
            "setp.lt.and.f32 p0, %4, %0, p1;\n\t"
            "@p1 selp.b32 %3, %2, %5, p0;\n\t"
            "@p0 mov.b32 %2, %5;\n\t"
            "@p1 min.f32 %0, %0, %4;\n\t}"
            : "+f"(d[0]), "+f"(d[1]), "+r"(i[0]), "+r"(i[1])
            : "f"(x), "r"(col));
      } else {  // int, larger is better
        asm("{\n\t.reg .pred p0, p1;\n\t"
            "setp.gt.s32 p1, %4, %1;\n\t"
            "@p1 min.s32 %1, %0, %4;\n\t"
            "setp.gt.and.s32 p0, %4, %0, p1;\n\t"
            "@p1 selp.b32 %3, %2, %5, p0;\n\t"
            "@p0 mov.b32 %2, %5;\n\t"
            "@p1 max.s32 %0, %0, %4;\n\t}"
            : "+r"(d[0]), "+r"(d[1]), "+r"(i[0]), "+r"(i[1])
            : "r"(x), "r"(col));
      }
      return;
    }
#pragma unroll
    for (int s = KTOP - 1; s >= 0; --s) {
      const bool lt_prev = (s > 0) ? O::better(x, d[s > 0 ? s - 1 : 0]) : false;
      const bool lt_cur = O::better(x, d[s]);
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

// (distance, index)-lexicographic list: the end-of-unit merge of the column parts of a row
template <int KTOP>
struct Final {
  float d[KTOP];
  int i[KTOP];
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

// 32 accumulator columns of one row.  Fast path: one 3-input min/max based extremum per
// group of 4 columns and a warp vote; the insertion network runs (for the whole
// warp, uniformly: no divergence) only for columns where some lane can beat its
// bound.  `pb` is the first value NOT admissible according to the partner
// threads that own the other columns of this row (ties with the partner's
// bound are admitted; the final merge orders them by index).  Expected
// insertions per row over M columns are ~k*ln(M/k), so almost every group takes
// the short fast path.
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

// kKey (Hamming): the accumulators are keys 32 * distance + column-in-slice (convert_hamming_kernel) while bounds and
// lists hold distances: tests compare against 32 * bound, an inserted value is decoded with one FFMA (the column is a
// compile-time constant here).
template <int J, bool kKey, typename T>
__device__ __forceinline__ T key_value(T w) {
  if constexpr (kKey) return fmaf(w, 0.03125f, -static_cast<float>(J) * 0.03125f);
  else return w;
}
template <bool kKey, typename T>
__device__ __forceinline__ T key_bound(T te) {
  if constexpr (kKey) return te * 32.0f;
  else return te;
}
template <int KTOP, typename O, int J0, bool kFixed = false, bool kKey = false>
__device__ __forceinline__ void consume_group(const typename O::T* w, int tp, TopK<KTOP, O>& tk, typename O::T te_in) {
  const typename O::T te = key_bound<kKey>(te_in);
  if (kFixed) {  // profiling aid: every warp does the same slow-path work (4 element tests, 2 insertions)
    const bool f0 = any_lane(O::better(w[0], O::worst())), f1 = any_lane(O::better(w[1], O::never()));
    const bool f2 = any_lane(O::better(w[2], O::worst())), f3 = any_lane(O::better(w[3], O::never()));
    if (f0) tk.insert(key_value<J0, kKey>(w[0]), enc_index<J0>(tp));
    if (f1) tk.insert(key_value<J0 + 1, kKey>(w[1]), enc_index<J0 + 1>(tp));
    if (f2) tk.insert(key_value<J0 + 2, kKey>(w[2]), enc_index<J0 + 2>(tp));
    if (f3) tk.insert(key_value<J0 + 3, kKey>(w[3]), enc_index<J0 + 3>(tp));
    return;
  }
#if IAM_GROUP_UNCOND
  // A/B aid: no per-column tests; the (self-guarded) insertion network runs for all four columns of a triggered group
  (void)te;
  tk.insert(key_value<J0, kKey>(w[0]), enc_index<J0>(tp));
  tk.insert(key_value<J0 + 1, kKey>(w[1]), enc_index<J0 + 1>(tp));
  tk.insert(key_value<J0 + 2, kKey>(w[2]), enc_index<J0 + 2>(tp));
  tk.insert(key_value<J0 + 3, kKey>(w[3]), enc_index<J0 + 3>(tp));
  return;
#endif
  // four votes issued back to back (computed against the bound at group entry: a superset of what
  // the tightening bound would admit), then the branch-free network only where some lane qualifies
  const bool e0 = any_lane(O::better(w[0], te));
  const bool e1 = any_lane(O::better(w[1], te));
  const bool e2 = any_lane(O::better(w[2], te));
  const bool e3 = any_lane(O::better(w[3], te));
  if (e0) tk.insert(key_value<J0, kKey>(w[0]), enc_index<J0>(tp));
  if (e1) tk.insert(key_value<J0 + 1, kKey>(w[1]), enc_index<J0 + 1>(tp));
  if (e2) tk.insert(key_value<J0 + 2, kKey>(w[2]), enc_index<J0 + 2>(tp));
  if (e3) tk.insert(key_value<J0 + 3, kKey>(w[3]), enc_index<J0 + 3>(tp));
}

// Per 16-column batch the four group tests are formed and voted on up front against the bound at
// entry (it only tightens, so the votes stay conservative): independent min3/setp/vote chains
// instead of serialised vote->branch round trips.
template <int KTOP, typename O, int H, bool kFixed = false, bool kKey = false>
__device__ __forceinline__ void consume16(const typename O::T* w, int tp, TopK<KTOP, O>& tk, typename O::T pb) {
  using T = typename O::T;
  const T te = kFixed ? O::never() : key_bound<kKey>(O::best(tk.thr(), pb));
  const bool t0 = any_lane(O::better(O::best(O::best3(w[0], w[1], w[2]), w[3]), te));
  const bool t1 = any_lane(O::better(O::best(O::best3(w[4], w[5], w[6]), w[7]), te));
  const bool t2 = any_lane(O::better(O::best(O::best3(w[8], w[9], w[10]), w[11]), te));
  const bool t3 = any_lane(O::better(O::best(O::best3(w[12], w[13], w[14]), w[15]), te));
  if (t0) consume_group<KTOP, O, H * 16, false, kKey>(w, tp, tk, O::best(tk.thr(), pb));
  if (t1 || kFixed) consume_group<KTOP, O, H * 16 + 4, kFixed, kKey>(w + 4, tp, tk, O::best(tk.thr(), pb));
  if (t2) consume_group<KTOP, O, H * 16 + 8, false, kKey>(w + 8, tp, tk, O::best(tk.thr(), pb));
  if (t3) consume_group<KTOP, O, H * 16 + 12, false, kKey>(w + 12, tp, tk, O::best(tk.thr(), pb));
}

template <int KTOP, typename O, bool kFixed = false, bool kKey = false>
__device__ __forceinline__ void consume32(const typename O::T (&v)[32], int tp, TopK<KTOP, O>& tk, typename O::T pb) {
  consume16<KTOP, O, 0, kFixed, kKey>(&v[0], tp, tk, pb);
  consume16<KTOP, O, 1, kFixed, kKey>(&v[16], tp, tk, pb);
}

// ---------------------------------------------------------------------------------------------------------------
// Packed-key path (byte layout, k = 2).  The s32 accumulators are < 2^24 (layout.h), so
//     key_j = acc_j * 32 + (31 - j)
// is an order-preserving, UNIQUE 29-bit key of the 32 columns a thread holds: larger key = nearer descriptor, and
// among equal distances the lower column (the order cv2.BFMatcher reports ties in).  With unique keys the two best
// columns of the slice need no compare/select network and no per-column votes:
//     m1 = max_j key_j                                   16 three-input max (ALU pipe)
//     u_j = key_j - m1  (mod 2^32)                       the winner becomes 0, every other column a huge unsigned
//     m2 = m1 + umax_j u_j                               number that still orders like its key
// The 32 packing multiply-adds are IMADs (FMA pipe, otherwise idle in this kernel; the multiplier is a kernel
// parameter so that ptxas cannot turn them into ALU-pipe LEAs).  The knock-out takes IAM_PACKED_FMA_SUBS columns as
// subtraction (IMAD.IADD) + three-input umax tree and the rest as four chains of fused add-max (VIADDMNMX.U32, one
// ALU instruction per column): the measured optimum between issue slots, the two pipes and dependent-chain length
// (DESIGN.md section 4).  Only the warp vote "some lane's m1 beats its bound" guards the second half, and the two
// winners enter the running list through the ordinary six-instruction insertion.
// The cost per slice no longer depends on how many columns qualify, which is what made the first tiles of every
// unit (empty lists, everything qualifies) cost seven times a late tile.
#ifndef IAM_PACKED
#define IAM_PACKED 1
#endif
#ifndef IAM_ROW_BOUND
#define IAM_ROW_BOUND 1             // pair loop: bound = the row's true second best (parts publish best and second best)
#endif
#ifndef IAM_RAW_FIRST
#define IAM_RAW_FIRST 0             // A/B aid: 1 = per-slice test on the raw accumulators, keys built only on a hit (measured 18.0 vs 16.4 ms: the second tree lands on the ALU pipe)
#endif
#ifndef IAM_LAT_EX
#define IAM_LAT_EX 0                // A/B aid: 1 = the exchange of bounds runs while the tensor-memory load is in flight
#endif
#ifndef IAM_LAT_CHAIN
#define IAM_LAT_CHAIN 1             // four knock-out chains and literal subtractions (0: two chains, IMAD with a register multiplier)
#endif
#ifndef IAM_FMA_DECODE
#define IAM_FMA_DECODE 0            // A/B aid: winners' (value, index) by IMAD.HI / IMAD instead of SHF / LOP3
#endif
#ifndef IAM_PACKED_HAMMING
#define IAM_PACKED_HAMMING 1        // packed-key epilogue for kind::f8f6f4 (Hamming), k = 2
#endif
#ifndef IAM_EPI_NOSYNC
#define IAM_EPI_NOSYNC 0           // A/B aid: 1 = no __syncwarp around the tensor-memory load of the pair loop
#endif
#ifndef IAM_PACK_IMAD
#define IAM_PACK_IMAD 1          // A/B aid: 0 = literal multipliers (ptxas emits LEA / IADD on the ALU pipe)
#endif
#ifndef IAM_SPARSE2
#define IAM_SPARSE2 0            // A/B aid: 1 = slices where few lanes have a candidate take a divergent second half (only those
#endif                           // lanes run; runner-up from the first half's group maxima by a switch on the winner's column).
                                 // Fewer instructions (~36 against 59) but measured SLOWER: 17.2 / 17.5 / 17.9 / 18.6 ms with at
                                 // most 2 / 4 / 8 / 32 lanes on that route against 15.85 ms (the compare tree of the switch and
                                 // the re-convergence lengthen the warp's critical path; 72 registers spill)
#ifndef IAM_SPARSE2_MAX
#define IAM_SPARSE2_MAX 4        // at most this many lanes with a candidate take the sparse route; more: the knock-out
#endif
#ifndef IAM_SPLIT_N
#define IAM_SPLIT_N 0            // A/B aid: 1 = every column part of an accumulator slot is its own product (kKSteps MMAs of
#endif                           // N = 32 / 48 instead of N = 96) with its own full / empty barriers, so that a slot waits for
                                 // the 4 warps of ONE part instead of all 12 of the tile (profile: 15 % of the epilogue warps'
                                 // time goes into waiting for accumulators while the tensor pipe is 45 % busy).  Measured
                                 // SLOWER: 18.8 against 15.9 ms (32-column parts), 15.8 against 15.3 ms (48-column parts):
                                 // three times the MMA instructions, and narrow MMAs do not run at the N = 96 rate
#ifndef IAM_PACKED_FMA_SUBS
#define IAM_PACKED_FMA_SUBS 8    // columns whose knock-out subtraction is an IMAD + tree (FMA pipe); the others: fused add-max chains. 32: all IMAD
#endif
template <int J>
__device__ __forceinline__ int pack_key(int acc, uint32_t mul32) {
  int k;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(k) : "r"(acc), "r"(mul32), "n"(31 - J));
  return k;
}
template <int J>
__device__ __forceinline__ uint32_t knock(int key, int m1, int neg_m1, uint32_t one) {
  (void)m1;
  uint32_t u;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(u) : "r"(key), "r"(one), "r"(neg_m1));
  return u;
}
__device__ __forceinline__ uint32_t umax3(uint32_t a, uint32_t b, uint32_t c) { return max(max(a, b), c); }
__device__ __forceinline__ uint32_t umin3(uint32_t a, uint32_t b, uint32_t c) { return min(min(a, b), c); }
// maximum of N unsigned values by three-input maxima
template <int N>
__device__ __forceinline__ uint32_t umax_tree(const uint32_t* x) {
  if constexpr (N == 1) {
    return x[0];
  } else if constexpr (N == 2) {
    return max(x[0], x[1]);
  } else if constexpr (N == 3) {
    return umax3(x[0], x[1], x[2]);
  } else {
    constexpr int kM = (N + 2) / 3;
    uint32_t y[kM];
#pragma unroll
    for (int i = 0; i < N / 3; ++i) y[i] = umax3(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
    if constexpr (N % 3 == 1) y[kM - 1] = x[N - 1];
    if constexpr (N % 3 == 2) y[kM - 1] = max(x[N - 2], x[N - 1]);
    return umax_tree<kM>(y);
  }
}

// maximum of N signed values by three-input maxima
template <int N>
__device__ __forceinline__ int imax_tree(const int* x) {
  if constexpr (N == 1) {
    return x[0];
  } else if constexpr (N == 2) {
    return max(x[0], x[1]);
  } else if constexpr (N == 3) {
    return imax3(x[0], x[1], x[2]);
  } else {
    constexpr int kM = (N + 2) / 3;
    int y[kM];
#pragma unroll
    for (int i = 0; i < N / 3; ++i) y[i] = imax3(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
    if constexpr (N % 3 == 1) y[kM - 1] = x[N - 1];
    if constexpr (N % 3 == 2) y[kM - 1] = max(x[N - 2], x[N - 1]);
    return imax_tree<kM>(y);
  }
}
// The first half leaves the maxima b[g] of the column groups {0..8}, {9..17}, {18..26}, {27..31}.  When column J holds the
// slice's best key, the runner-up is the largest of the other groups' maxima and the other columns of J's own group:
// 11 (or 7) values picked at compile time -- five (three) three-input maxima, against ~40 instructions of knock-out.
template <int J>
__device__ __forceinline__ int runner_up(const int (&k)[32], const int (&b)[4]) {
  constexpr int kG = J < 27 ? J / 9 : 3;
  constexpr int kLo = kG * 9, kHi = kG < 3 ? kLo + 9 : 32;
  int x[3 + (kHi - kLo) - 1];
  int n = 0;
#pragma unroll
  for (int g = 0; g < 4; ++g)
    if (g != kG) x[n++] = b[g];
#pragma unroll
  for (int j = kLo; j < kHi; ++j)
    if (j != J) x[n++] = k[j];
  return imax_tree<3 + (kHi - kLo) - 1>(x);
}
template <typename F, int... Js>
__device__ __forceinline__ void static_for32(F&& f, std::integer_sequence<int, Js...>) {
  (f(std::integral_constant<int, Js>{}), ...);
}
// kMode 0: production; 1: profiling aid, second half never taken; 2: profiling aid, second half always taken.
// The running lists carry the index  e = tp32 | (31 - j)  (tp32 = 32 * slice number): the low five bits come
// straight out of the key (one LOP3 per winner), and the real column is  e ^ 31.
template <int kMode = 0>
__device__ __forceinline__ void consume32_packed(const int (&v)[32], int tp32, TopK<2, Ord<Kind::I8>>& tk, int pb,
                                                 uint32_t mul32, uint32_t one) {
  const int te = max(tk.d[1], pb);
  const uint32_t two27 = mul32 << 22;                      // opaque like mul32: stays a register operand
  const int neg32 = -static_cast<int>(mul32);
  (void)two27;
  (void)neg32;
#if IAM_RAW_FIRST
  // the test needs no keys: largest raw accumulator of the slice against the bound
  int ra[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) ra[i] = imax3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
  const int vmax = max(imax3(imax3(ra[0], ra[1], ra[2]), imax3(ra[3], ra[4], ra[5]), imax3(ra[6], ra[7], ra[8])),
                       imax3(ra[9], v[30], v[31]));
  const bool hit = kMode == 2 ? any_lane(vmax > -1) : any_lane(vmax > te);
  if (hit && kMode != 1) {
    int k[32];
    static_for32([&](auto j) { k[j] = pack_key<j>(v[j], mul32); }, std::make_integer_sequence<int, 32>{});
    int a[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) a[i] = imax3(k[3 * i], k[3 * i + 1], k[3 * i + 2]);
    const int m1 = max(imax3(imax3(a[0], a[1], a[2]), imax3(a[3], a[4], a[5]), imax3(a[6], a[7], a[8])),
                       imax3(a[9], k[30], k[31]));
#else
  int k[32];
  static_for32([&](auto j) { k[j] = pack_key<j>(v[j], mul32); }, std::make_integer_sequence<int, 32>{});
  int a[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) a[i] = imax3(k[3 * i], k[3 * i + 1], k[3 * i + 2]);
  const int b0 = imax3(a[0], a[1], a[2]), b1 = imax3(a[3], a[4], a[5]), b2 = imax3(a[6], a[7], a[8]);
  const int b3 = imax3(a[9], k[30], k[31]);
  const int m1 = max(imax3(b0, b1, b2), b3);
  // admissible: acc > te  <=>  key > te * 32 + 31
  int te_key;
  asm("mad.lo.s32 %0, %1, %2, 31;" : "=r"(te_key) : "r"(te), "r"(mul32));
#if IAM_SPARSE2
  if constexpr (kMode == 0) {
    // Late in a unit one or two of a warp's 32 rows have a candidate in a slice: those lanes alone run the second
    // half.  The low five bits of the best key name its column, the switch picks the compile-time set of registers
    // the runner-up can sit in.  Slices where many lanes qualify (the first tiles of a unit) take the knock-out below.
    const uint32_t cand = __ballot_sync(0xffffffffu, m1 > te_key);
    if (cand == 0) return;
    if (__popc(cand) <= IAM_SPARSE2_MAX) {
      if (m1 > te_key) {
        const int bb[4] = {b0, b1, b2, b3};
        int m2 = 0;
        switch (m1 & 31) {  // = 31 - column
#define IAM_RU(J) case 31 - J: m2 = runner_up<J>(k, bb); break;
          IAM_RU(0) IAM_RU(1) IAM_RU(2) IAM_RU(3) IAM_RU(4) IAM_RU(5) IAM_RU(6) IAM_RU(7)
          IAM_RU(8) IAM_RU(9) IAM_RU(10) IAM_RU(11) IAM_RU(12) IAM_RU(13) IAM_RU(14) IAM_RU(15)
          IAM_RU(16) IAM_RU(17) IAM_RU(18) IAM_RU(19) IAM_RU(20) IAM_RU(21) IAM_RU(22) IAM_RU(23)
          IAM_RU(24) IAM_RU(25) IAM_RU(26) IAM_RU(27) IAM_RU(28) IAM_RU(29) IAM_RU(30) IAM_RU(31)
#undef IAM_RU
        }
        tk.insert(m1 >> 5, (m1 & 31) | tp32);
        if (m2 > te_key) tk.insert(m2 >> 5, (m2 & 31) | tp32);
      }
      return;
    }
  }
  const bool hit = kMode == 0 ? true : kMode == 2 ? any_lane(m1 > -1) : any_lane(m1 > te_key);
#else
  const bool hit = kMode == 2 ? any_lane(m1 > -1) : any_lane(m1 > te_key);
#endif
  if (hit && kMode != 1) {
#endif
    const int neg_m1 = -m1;
    uint32_t best;
    if constexpr (IAM_PACKED_FMA_SUBS >= 32) {  // every subtraction an IMAD, one three-input tree
      uint32_t u[32];
      static_for32([&](auto j) { u[j] = knock<j>(k[j], m1, neg_m1, one); }, std::make_integer_sequence<int, 32>{});
      uint32_t c[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) c[i] = umax3(u[3 * i], u[3 * i + 1], u[3 * i + 2]);
      const uint32_t d0 = umax3(c[0], c[1], c[2]), d1 = umax3(c[3], c[4], c[5]), d2 = umax3(c[6], c[7], c[8]);
      const uint32_t d3 = umax3(c[9], u[30], u[31]);
      best = max(umax3(d0, d1, d2), d3);
    } else {
      // Issue slots are the scarce resource once both pipes are loaded: the first kNF columns take the IMAD +
      // three-input tree route (kNF + kNF/2 instructions), the others two chains of fused add-max
      // (VIADDMNMX.U32, one ALU-pipe instruction per column).
      constexpr int kNF = IAM_PACKED_FMA_SUBS;
      static_assert(kNF >= 32 || (kNF >= 2 && kNF % 2 == 0), "IAM_PACKED_FMA_SUBS");
      uint32_t u[kNF];
      const uint32_t nm = static_cast<uint32_t>(neg_m1);
#if IAM_LAT_CHAIN
      // plain subtractions (ptxas issues them as IMAD.IADD with a literal 1: no multiplier register to reload) and
      // FOUR add-max chains: the second half sits on every warp's critical path, and the warps of a tile move in
      // lock step with the MMAs, so dependent-chain length counts as much as instruction count
#pragma unroll
      for (int j = 0; j < kNF; ++j) u[j] = static_cast<uint32_t>(k[j]) - static_cast<uint32_t>(m1);
      uint32_t s0 = umax_tree<kNF / 2>(u), s1 = umax_tree<kNF / 2>(u + kNF / 2), s2 = 0, s3 = 0;
      constexpr int kQ = (32 - kNF) / 4;
#pragma unroll
      for (int j = 0; j < kQ; ++j) {
        s0 = max(s0, static_cast<uint32_t>(k[kNF + j]) + nm);
        s1 = max(s1, static_cast<uint32_t>(k[kNF + kQ + j]) + nm);
        s2 = max(s2, static_cast<uint32_t>(k[kNF + 2 * kQ + j]) + nm);
      }
#pragma unroll
      for (int j = kNF + 3 * kQ; j < 32; ++j) s3 = max(s3, static_cast<uint32_t>(k[j]) + nm);
      best = max(umax3(s0, s1, s2), s3);
#else
      static_for32([&](auto j) { u[j] = knock<j>(k[j], m1, neg_m1, one); }, std::make_integer_sequence<int, kNF>{});
      uint32_t s0 = umax_tree<kNF / 2>(u), s1 = umax_tree<kNF / 2>(u + kNF / 2);
      constexpr int kMid = kNF + (32 - kNF) / 2;
#pragma unroll
      for (int j = kNF; j < kMid; ++j) s0 = max(s0, static_cast<uint32_t>(k[j]) + nm);
#pragma unroll
      for (int j = kMid; j < 32; ++j) s1 = max(s1, static_cast<uint32_t>(k[j]) + nm);
      best = max(s0, s1);
#endif
    }
    const int m2 = m1 + static_cast<int>(best);
#if IAM_FMA_DECODE
    // value = key >> 5 as the high half of key * 2^27, index = key - 32 * value + tp32: multiply-adds (FMA pipe)
    // instead of a shift and a LOP3 on the ALU pipe, which is the busier one
    const int x1 = __umulhi(static_cast<uint32_t>(m1), two27), x2 = __umulhi(static_cast<uint32_t>(m2), two27);
    int e1, e2;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(e1) : "r"(x1), "r"(neg32), "r"(m1 + tp32));
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(e2) : "r"(x2), "r"(neg32), "r"(m2 + tp32));
    tk.insert(x1, e1);
    tk.insert(x2, e2);
#else
    tk.insert(m1 >> 5, (m1 & 31) | tp32);
    tk.insert(m2 >> 5, (m2 & 31) | tp32);
#endif
  }
}

// The packed-key scheme for slices of NC columns (IAM_PART_COLS = 48 / 96): key = acc * kMul + (kMul - 1 - column),
// kMul = 64 / 128 (acc <= 128 * 255^2 + kI8Cap < 2^24, so the key stays below 2^31); the running lists carry the index
// e = tpk | (kMul - 1 - j) with tpk = kMul * slice number.  Same two halves as consume32_packed: kMul-ary packing by
// IMAD (FMA pipe), a three-input max tree, ONE vote, then the knock-out (first kNF columns as subtraction + tree, the
// rest as four fused add-max chains) and two insertions.
template <int NC, int kMode = 0>
__device__ __forceinline__ void consume_packed_n(const int (&v)[NC], int tpk, TopK<2, Ord<Kind::I8>>& tk, int pb,
                                                 uint32_t mul, uint32_t one) {
  constexpr int kMul = NC <= 32 ? 32 : NC <= 64 ? 64 : 128;
  constexpr int kShift = kMul == 32 ? 5 : kMul == 64 ? 6 : 7;
  (void)one;
  const int te = max(tk.d[1], pb);
  int k[NC];
  static_for32([&](auto j) { asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(k[j]) : "r"(v[j]), "r"(mul), "n"(kMul - 1 - decltype(j)::value)); },
               std::make_integer_sequence<int, NC>{});
  const int m1 = imax_tree<NC>(k);
  int te_key;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(te_key) : "r"(te), "r"(mul), "n"(kMul - 1));
  const bool hit = kMode == 2 ? any_lane(m1 > -1) : any_lane(m1 > te_key);
  if (hit && kMode != 1) {
    constexpr int kNF = IAM_PACKED_FMA_SUBS >= 32 ? 8 : IAM_PACKED_FMA_SUBS;
    const uint32_t nm = static_cast<uint32_t>(-m1);
    uint32_t u[kNF];
#pragma unroll
    for (int j = 0; j < kNF; ++j) u[j] = static_cast<uint32_t>(k[j]) - static_cast<uint32_t>(m1);
    uint32_t s0 = umax_tree<kNF / 2>(u), s1 = umax_tree<kNF / 2>(u + kNF / 2), s2 = 0, s3 = 0;
    constexpr int kQ = (NC - kNF) / 4;
#pragma unroll
    for (int j = 0; j < kQ; ++j) {
      s0 = max(s0, static_cast<uint32_t>(k[kNF + j]) + nm);
      s1 = max(s1, static_cast<uint32_t>(k[kNF + kQ + j]) + nm);
      s2 = max(s2, static_cast<uint32_t>(k[kNF + 2 * kQ + j]) + nm);
    }
#pragma unroll
    for (int j = kNF + 3 * kQ; j < NC; ++j) s3 = max(s3, static_cast<uint32_t>(k[j]) + nm);
    const int m2 = m1 + static_cast<int>(max(umax3(s0, s1, s2), s3));
    tk.insert(m1 >> kShift, (m1 & (kMul - 1)) | tpk);
    tk.insert(m2 >> kShift, (m2 & (kMul - 1)) | tpk);
  }
}

// Packed-key path for Hamming (kind::f8f6f4, k = 2).  The operand layout makes every fp32 accumulator the exact
// integer  key_j = 32 * distance + j  (convert_hamming_kernel; j = column inside the thread's 32-column slice),
// unique, ordered like (distance, column): SMALLER = nearer, ties to the lower column.  m1 = min_j key_j by
// three-input float minima; the runner-up by a knock-out on the raw bit patterns (below).
// The running lists stay in the float domain of Ord<F8>; the index is  e = tp32 | j  (the real column).
template <int kMode = 0>
__device__ __forceinline__ void consume32_packed_f(const float (&v)[32], int tp32, TopK<2, Ord<Kind::F8>>& tk, float pb) {
  const float (&k)[32] = v;  // the accumulators ARE the keys (convert_hamming_kernel)
  float a[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) a[i] = fmin3(k[3 * i], k[3 * i + 1], k[3 * i + 2]);
  const float b0 = fmin3(a[0], a[1], a[2]), b1 = fmin3(a[3], a[4], a[5]), b2 = fmin3(a[6], a[7], a[8]);
  const float b3 = fmin3(a[9], k[30], k[31]);
  const float m1 = fminf(fmin3(b0, b1, b2), b3);
  // admissible: acc < te (te an integer or +big: own second best, or a partner's second best + 1)  <=>  key < te * 32
  const float te_key = fminf(tk.d[1], pb) * 32.0f;
  const bool hit = kMode == 2 ? any_lane(m1 > -1.0f) : any_lane(m1 < te_key);
  if (hit && kMode != 1) {
    // Knock-out: t_j = key_j - (m1 + 1/4) is -1/4 for the winner and n - 1/4 (n >= 1) for every other column.  Read
    // as UNSIGNED integers, non-negative floats order like their values and the one negative float is larger than
    // all of them, so the runner-up is an unsigned minimum over the raw bit patterns (three-input VIMNMX3.U32).
    const float m1q = m1 + 0.25f;
    uint32_t u[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(k[j] - m1q);
    uint32_t c[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) c[i] = umin3(u[3 * i], u[3 * i + 1], u[3 * i + 2]);
    const uint32_t d0 = umin3(c[0], c[1], c[2]), d1 = umin3(c[3], c[4], c[5]), d2 = umin3(c[6], c[7], c[8]);
    const uint32_t d3 = umin3(c[9], u[30], u[31]);
    const float m2 = __uint_as_float(min(umin3(d0, d1, d2), d3)) + m1q;
    // key -> (distance, column): key + 2^23 carries the integer key in its mantissa, the column in its low five bits
    const uint32_t q1 = __float_as_uint(m1 + 8388608.0f), q2 = __float_as_uint(m2 + 8388608.0f);
    const float j1 = __uint_as_float((q1 & 31u) | 0x4b000000u) - 8388608.0f;
    const float j2 = __uint_as_float((q2 & 31u) | 0x4b000000u) - 8388608.0f;
    tk.insert((m1 - j1) * 0.03125f, static_cast<int>(q1 & 31u) | tp32);
    tk.insert((m2 - j2) * 0.03125f, static_cast<int>(q2 & 31u) | tp32);
  }
}

template <Kind kKind>
__device__ __forceinline__ const uint8_t* a_src(const ImgDev& im) { return kKind == Kind::I8 ? im.i8_form : im.a_form; }
template <Kind kKind>
__device__ __forceinline__ const uint8_t* b_src(const ImgDev& im) { return kKind == Kind::I8 ? im.i8_form : im.b_form; }

template <Kind kKind, int KTOP, bool kATmem, bool kCluster, int kDbg, int kT, int kPC>
__global__ void __launch_bounds__((Cfg<kKind, kT, kPC>::kThreads), (Cfg<kKind, kT, kPC>::kCtasPerSm))
knn_umma_kernel(const ImgDev* __restrict__ imgs, const KnnUnit* __restrict__ units, int n_units,
                int* __restrict__ out_idx, float* __restrict__ out_d2, uint32_t pack_mul, uint32_t pack_one) {
  using C = Cfg<kKind, kT, kPC>;
  constexpr int kEpiWarps = C::kEpiWarps;
  constexpr int kParts = C::kParts;
  constexpr int kWarpsPerATile = C::kWarpsPerATile;
  constexpr int kKeyMul = C::kKeyMul;
  constexpr int kRows = C::kRows;
  constexpr int kItems = C::kItems;
  constexpr uint32_t kTmemCols = C::kTmemCols;
  const int n_items = n_units * kItems;   // a work item = kRows query rows of one unit against its whole train image
  using O = Ord<kKind>;
  using T = typename O::T;
  using Barriers = typename C::Barriers;
  constexpr int kBStages = C::kBStages;
  constexpr int kSlots = C::kSlots;
  constexpr bool kDual = C::kDual && kATmem;
  constexpr bool kPackedK = IAM_PACKED && kKind == Kind::I8 && KTOP == 2;
  constexpr bool kPairK = kPackedK && kDual && kSlots == 2 * kT && (kParts > 1 || kPC != 32) &&
                          (kDbg == 0 || kDbg == 1 || kDbg == 4 || kDbg == 5);
  constexpr bool kSplit = IAM_SPLIT_N && kPairK && kParts > 1;   // one product (and one pair of barriers) per column part
  constexpr int kBarsPerSlot = kSplit ? kParts : 1;
  constexpr int kKSteps = C::kKSteps;
  constexpr uint32_t kTileBytes = C::kTileBytes;
  constexpr uint32_t kBTileBytes = C::kBTileBytes;
  constexpr uint32_t kRowBytes = C::kRowBytes;
  constexpr uint32_t kSBO = C::kSBO;
  constexpr uint32_t kTmemA = C::kTmemA;
  constexpr uint32_t kTmemAColsPerTile = C::kTmemAColsPerTile;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kSmemA;
  Barriers* bars = reinterpret_cast<Barriers*>(smem + C::kSmemA + C::kSmemB);
  uint32_t* share = reinterpret_cast<uint32_t*>(smem + C::kSmemA + C::kSmemB + C::kSmemBars);               // [unit parity][part][row]
  float2* merge = reinterpret_cast<float2*>(smem + C::kSmemA + C::kSmemB + C::kSmemBars + C::kSmemShare);  // [part-1][row][k]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // role of this warp: 0 B-tile producer, 1 MMA issuer, 2 TMEM allocator, 3 A-tile producer, 4.. epilogue
  const int role = IAM_ROLES_LAST ? (warp >= kEpiWarps ? warp - kEpiWarps : warp + 4) : warp;
  // With kCluster two CTAs (one cluster) walk the unit list in lock step on the two units 2p, 2p+1 of the
  // same directed job: each CTA fetches HALF of every train tile and multicasts it to both, which halves
  // the L2 -> SM traffic of the streamed operand.
  constexpr int kCtas = kCluster ? 2 : 1;
  const int cta_rank = kCluster ? static_cast<int>(cluster_ctarank()) : 0;
  const int first_pu = blockIdx.x / kCtas;
  const int pu_stride = gridDim.x / kCtas;

  for (int i = threadIdx.x; i < 2 * 2 * kParts * kRows; i += blockDim.x) share[i] = O::bits(O::worst());
  if (role == 1 && elect_one()) {
    for (int i = 0; i < kT; ++i) {
      mbar_init(&bars->a_full[i], 1);
      mbar_init(&bars->a_empty[i], 1);
    }
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(&bars->b_full[i], 1);
      mbar_init(&bars->b_empty[i], kCtas * (kDual ? kT : 1));  // every consumer of every CTA that received the tile
    }
    for (int i = 0; i < kSlots * kBarsPerSlot; ++i) {
      mbar_init(&bars->t_full[i], 1);
      mbar_init(&bars->t_empty[i], kWarpsPerATile / kBarsPerSlot);  // the warps of whichever A tile last used the slot (of one part when split)
    }
    fence_barrier_init();
  } else if (role == 2) {
    tmem_alloc<kTmemCols>(&bars->tmem_base);
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster) cluster_sync_all();  // peer barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (role == 0) {
    // ------------------------------------------------ B-tile producer
    if (elect_one()) {
      uint32_t it = 0;  // running B-tile counter across units
      for (int pu = first_pu; pu * kCtas < n_items; pu += pu_stride) {
        const int u = (pu * kCtas + cta_rank) / kItems;
        const KnnUnit unit = units[u];
        const ImgDev t = imgs[unit.t_slot];
        const uint8_t* tsrc = b_src<kKind>(t);
        const int n_tb = (t.n + kBRows - 1) / kBRows;
        for (int tb = 0; tb < n_tb; ++tb, ++it) {
          const uint32_t stage = it % kBStages;
          const uint32_t par = (it / kBStages) & 1;
          mbar_wait(&bars->b_empty[stage], par ^ 1, 10);
          mbar_arrive_expect_tx(&bars->b_full[stage], kBTileBytes);
          if (kCluster) {
            constexpr uint32_t kHalf = kBTileBytes / 2;  // half the train rows = whole core-matrix groups, contiguous
            bulk_g2s_multicast(smem_b + stage * kBTileBytes + cta_rank * kHalf,
                               tsrc + static_cast<size_t>(tb) * kBTileBytes + cta_rank * kHalf, kHalf,
                               &bars->b_full[stage], 0x3);
          } else {
            bulk_g2s(smem_b + stage * kBTileBytes, tsrc + static_cast<size_t>(tb) * kBTileBytes, kBTileBytes,
                     &bars->b_full[stage]);
          }
        }
      }
    }
  } else if (role == 3) {
    // ------------------------------------------------ A-tile producer
    if (elect_one()) {
      uint32_t it = 0;  // unit counter of this CTA
      for (int pu = first_pu; pu * kCtas < n_items; pu += pu_stride, ++it) {
        const int w = pu * kCtas + cta_rank;
        const KnnUnit unit = units[w / kItems];
        const ImgDev q = imgs[unit.q_slot];
        const uint8_t* src = a_src<kKind>(q) + (static_cast<size_t>(unit.super) * kSuperRows + (w % kItems) * kRows) * kRowBytes;
        const uint32_t par = it & 1;
        for (int a = 0; a < kT; ++a) {
          mbar_wait(&bars->a_empty[a], par ^ 1, 20 + a);
          mbar_arrive_expect_tx(&bars->a_full[a], kTileBytes);
          bulk_g2s(smem_a + a * kTileBytes, src + static_cast<size_t>(a) * kTileBytes, kTileBytes, &bars->a_full[a]);
        }
      }
    }
  } else if (role == 1 || (kDual && role == 2)) {
    // ------------------------------------------------ MMA issuer(s)
    // The whole warp walks the loops with warp-uniform values (made provably uniform by a shuffle, so the
    // descriptor arithmetic lives in the uniform datapath instead of vector registers + R2UR); one elected lane
    // issues the tcgen05 instructions.  With kDual each query tile has its own issuer (roles 1 and 2) with its own
    // half of the accumulator slots; otherwise one warp issues the products of both tiles in turn.
    constexpr uint32_t idesc = make_idesc_kind<kKind>(128, kSplit ? kPC : kBRows);
    constexpr int kMyTiles = kDual ? 1 : kT;                  // query tiles this warp issues for
    constexpr int kSlotStep = kDual ? kT : 1;                 // distance between successive slots of this warp
    const int a0 = kDual ? role - 1 : 0;
    const uint32_t a_addr = smem_u32(smem_a);
    const uint32_t b_addr = smem_u32(smem_b);
    const uint32_t bars_addr = smem_u32(bars);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t tm_a = tm + kTmemA;
    const uint32_t b_lo0 = smem_desc_lo(b_addr, kLBO);
    uint32_t stage = 0, bpar = 0;   // B ring position
    uint32_t slot = a0, tpar = 1;   // accumulator ring position; parity to wait for on t_empty (fresh barrier: 1)
    uint32_t uit = 0;
    for (int pu = first_pu; pu * kCtas < n_items; pu += pu_stride, ++uit) {
      const int u = (pu * kCtas + cta_rank) / kItems;
      const int n_t = __shfl_sync(0xffffffffu, imgs[units[u].t_slot].n, 0);
      const int n_tb = (n_t + kBRows - 1) / kBRows;
      // query tiles: shared-memory staging -> tensor memory (tcgen05.cp), then the staging is free again.
      // tcgen05 operations of one thread execute in issue order, so these copies run after every MMA of
      // the previous unit that still reads the old A tiles.
      for (int a = a0; a < a0 + kMyTiles; ++a) {
        mbar_wait(&bars->a_full[a], uit & 1, 32 + a);
        tc_fence_after();
        if (kATmem) {
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < kKSteps; ++ks)
              tmem_cp_128x256b(tm + kTmemA + a * kTmemAColsPerTile + ks * 8,
                               make_smem_desc(a_addr + a * kTileBytes + C::a_koff(ks), kLBO, kSBO));
            umma_commit(&bars->a_empty[a]);
          }
          __syncwarp();
        }
      }
      for (int tb = 0; tb < n_tb; ++tb) {
        mbar_wait_a(bars_addr + offsetof(Barriers, b_full) + stage * 8, bpar, 30);
        const uint32_t b_lo = b_lo0 + stage * (kBTileBytes >> 4);
#pragma unroll
        for (int am = 0; am < kMyTiles; ++am) {
          const int a = a0 + am;
          if constexpr (kSplit) {
            // one product per column part: kKSteps MMAs of N = kPC on that part's train rows (whole core-matrix
            // groups: kPC / 8 groups of kSBO bytes) into its own columns of the slot, gated by its own barriers
#pragma unroll
            for (int pt = 0; pt < kParts; ++pt) {
              mbar_wait_a(bars_addr + offsetof(Barriers, t_empty) + (slot * kParts + pt) * 8, tpar, 31);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t taddr = tm + slot * kBRows + pt * kPC;
                const uint32_t b_lo_p = b_lo + ((pt * (kPC / 8) * kSBO) >> 4);
#pragma unroll
                for (int ks = 0; ks < kKSteps; ++ks) {
                  const uint64_t bdesc = pack_desc(b_lo_p + (C::b_koff(ks) >> 4), smem_desc_hi(kSBO));
                  umma_ts<kKind>(taddr, tm_a + a * kTmemAColsPerTile + ks * 8, bdesc, idesc, ks > 0 ? 1u : 0u);
                }
                umma_commit_a(bars_addr + offsetof(Barriers, t_full) + (slot * kParts + pt) * 8);
              }
              __syncwarp();
            }
          } else {
          mbar_wait_a(bars_addr + offsetof(Barriers, t_empty) + slot * 8, tpar, 31);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t taddr = tm + slot * kBRows;
#pragma unroll
            for (int ks = 0; ks < kKSteps; ++ks) {
              const uint64_t bdesc = pack_desc(b_lo + (C::b_koff(ks) >> 4), smem_desc_hi(kSBO));
              if (kATmem) {
                umma_ts<kKind>(taddr, tm_a + a * kTmemAColsPerTile + ks * 8, bdesc, idesc, ks > 0 ? 1u : 0u);
              } else {
                const uint64_t adesc = make_smem_desc(a_addr + a * kTileBytes + C::a_koff(ks), kLBO, kSBO);
                umma<kKind>(taddr, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
              }
            }
            if (!kATmem && tb == n_tb - 1) umma_commit(&bars->a_empty[a]);
            umma_commit_a(bars_addr + offsetof(Barriers, t_full) + slot * 8);
          }
          }
          __syncwarp();
          slot += kSlotStep;
          if (slot >= static_cast<uint32_t>(kSlots)) {
            slot -= kSlots;
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
  } else if (role >= 4) {
    // ------------------------------------------------ epilogue
    // The per-tile loop is ALU-pipe bound (profiles/): everything loop-invariant lives in pinned registers as
    // ready-made shared-memory / tensor-memory addresses, and the slot / phase of the accumulator ring advance
    // incrementally instead of by division.
    const int e = (role - 4) >> 2;      // 0 .. kT*kParts-1
    const int a = e / kParts;           // which A tile
    const int part = e % kParts;        // which 32 of the B tile's columns
    const int quad = warp & 3;          // TMEM lane quadrant this warp may touch
    const int urow = a * kTileRows + quad * 32 + lane;  // row within the unit
    constexpr uint32_t kPartStride = kRows * 8;                 // bytes between the parts' bound slots of one row
    constexpr uint32_t kParityStride = kParts * kPartStride;    // bytes between the two unit-parity buffers
    const uint32_t share_row = smem_u32(share) + urow * 8;
    const uint32_t bar_full0 = pin_reg(smem_u32(&bars->t_full[0]));
    constexpr uint32_t kEmptyOff = kSlots * kParts * 8;         // t_empty[] follows t_full[] in Barriers
    const uint32_t tm_warp = pin_reg(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + part * kPC);
    const bool lane0 = lane == 0;
    TopK<KTOP, O> tk;
    constexpr bool kPacked = IAM_PACKED && kKind == Kind::I8 && KTOP == 2;
    static_assert(kPC == 32 || kPacked, "wide slices exist for the packed-key path only");
    constexpr bool kPackedF = IAM_PACKED && IAM_PACKED_HAMMING && kKind == Kind::F8 && KTOP == 2;
    constexpr bool kKeyF = kKind == Kind::F8;  // accumulators are keys 32 * distance + column (convert_hamming_kernel)
    // Multipliers ptxas cannot fold (kernel parameters 32 and 1): with a literal 32 the packing multiply-adds are strength-reduced
    // to LEA, an ALU-pipe instruction; as register operands they stay IMADs on the otherwise idle FMA pipe.
    const uint32_t mul32 = IAM_PACK_IMAD ? pack_mul : static_cast<uint32_t>(kKeyMul), one = IAM_PACK_IMAD ? pack_one : 1u;
    (void)mul32;
    (void)one;
    uint32_t slot = a % kSlots, par = 0;  // position in the accumulator ring: sq = it*kT + a, slot = sq % kSlots
    constexpr bool kPair = kPairK;
    static_assert(kPairK == (kPacked && kDual && kSlots == 2 * kT && (kParts > 1 || kPC != 32) && (kDbg == 0 || kDbg == 1 || kDbg == 4 || kDbg == 5)), "pair-loop condition");
    // kPair: t_full of this tile's first slot (of this warp's column part when the products are split); second slot: + kSlotBar
    constexpr uint32_t kSlotBar = kT * kBarsPerSlot * 8;
    const uint32_t bar_a = pin_reg(bar_full0 + (kSplit ? (a * kParts + part) : a) * 8);
    const uint32_t tm_a = pin_reg(tm_warp + a * kBRows);     // kPair: its accumulator columns (second slot: + kT * kBRows)
    int phase = 0;
    (void)bar_a;
    (void)tm_a;
    (void)phase;
    (void)slot;
    const int group_bar = 1 + a * 4 + quad;   // named barrier of the kParts warps that share these 32 rows
    int uit = 0;
    for (int pu = first_pu; pu * kCtas < n_items; pu += pu_stride, ++uit) {
      const int w = pu * kCtas + cta_rank;
      const KnnUnit unit = units[w / kItems];
      const ImgDev q = imgs[unit.q_slot];
      const ImgDev t = imgs[unit.t_slot];
      const int n_tb = (t.n + kBRows - 1) / kBRows;
      tk.reset();
      T pb = O::worst();  // first value NOT admissible according to the other column parts of this row
      // Bounds live in a buffer selected by the unit's parity.  At the start of unit u every thread resets its
      // slot in the OTHER buffer (the one unit u+1 will use); the end-of-unit barrier orders that reset before
      // any partner reads it, so a slot only ever holds the worst value or values of the unit being processed.
      const uint32_t rd = pin_reg(share_row + (uit & 1) * kParityStride);
      const uint32_t wr = pin_reg(rd + part * kPartStride);
      sts_volatile_v2b32_a(share_row + ((uit + 1) & 1) * kParityStride + part * kPartStride, O::bits(O::worst()), O::bits(O::worst()));
      if constexpr (kPair) {
        // Production shape: this query tile owns the accumulator slots a and a + 2 and uses them alternately, so the
        // tiles are walked in PAIRS with every barrier / tensor-memory address a constant offset from a pinned base,
        // one parity flip and one exchange of bounds per pair.  `phase` says which slot the unit's first tile sits in
        // (the ring runs on across units; an odd tile count flips it).
        // Bound from the threads that own the other column parts of this row (own slot included, it is harmless):
        // nothing worse than the best of the k-th bests can end up in the merged list.  Ties are admitted (loosen;
        // the final merge orders them by index).  Stale values are still valid bounds: plain volatile traffic.
        auto exchange = [&]() {
          if constexpr (kParts == 1) return;  // the whole row is this thread's: nothing to exchange
#if IAM_ROW_BOUND
          // The row's true second best so far: every part publishes (second best, best); the second largest of the
          // six values is  max(second largest of the three bests, largest of the three second bests).
          static_assert(!IAM_ROW_BOUND || kParts <= 3, "row-exact bound: up to three column parts");
          if constexpr (kParts == 3) {
            const uint2 e0 = lds_volatile_v2b32_a(rd), e1 = lds_volatile_v2b32_a(rd + kPartStride), e2 = lds_volatile_v2b32_a(rd + 2 * kPartStride);
            const T h0 = O::from_bits(e0.y), h1 = O::from_bits(e1.y), h2 = O::from_bits(e2.y);
            const T second_h = O::best3(min(h0, h1), min(h0, h2), min(h1, h2));
            const T g = O::best(second_h, O::best3(O::from_bits(e0.x), O::from_bits(e1.x), O::from_bits(e2.x)));
            pb = O::best(pb, O::loosen(g));
          } else if constexpr (kParts == 2) {  // four values: max(worse of the two bests, better of the two second bests)
            const uint2 e0 = lds_volatile_v2b32_a(rd), e1 = lds_volatile_v2b32_a(rd + kPartStride);
            const T g = O::best3(min(O::from_bits(e0.y), O::from_bits(e1.y)), O::from_bits(e0.x), O::from_bits(e1.x));
            pb = O::best(pb, O::loosen(g));
          }
#else
          T g = O::from_bits(lds_volatile_b32_a(rd));
#pragma unroll
          for (int pp = 1; pp < kParts; ++pp) g = O::best(g, O::from_bits(lds_volatile_b32_a(rd + pp * kPartStride)));
          pb = O::best(pb, O::loosen(g));
#endif
        };
        // `ex`: this is the pair's first tile -- the exchange of bounds runs while the tensor-memory load is in flight
        auto tile = [&](auto sl, int tpk, bool ex) {
          constexpr uint32_t kSl = decltype(sl)::value;
          const uint32_t bar = bar_a + kSl * kSlotBar;
          mbar_wait_bare_a(bar, par);
          tc_fence_after();
          if constexpr (kDbg != 1) {
            int v[kPC];
#if !IAM_EPI_NOSYNC
            __syncwarp();
#endif
            if constexpr (kPC == 32) tmem_ld32(tm_a + kSl * (kT * kBRows), v);
            else tmem_ld_n<kPC>(tm_a + kSl * (kT * kBRows), v);
            if (IAM_LAT_EX && ex) exchange();
            if constexpr (kPC == 32) tmem_ld_wait(v);
            else tmem_ld_wait_n<kPC>(v);
            // the values are in registers: hand the accumulator slot back to the MMA issuer before consuming them
#if !IAM_EPI_NOSYNC
            __syncwarp();
#endif
            tc_fence_before();
            if (lane0) mbar_arrive_a(bar + kEmptyOff);
            if constexpr (kPC == 32) consume32_packed<kDbg == 5 ? 2 : kDbg == 4 ? 1 : 0>(v, tpk, tk, pb, mul32, one);
            else consume_packed_n<kPC, kDbg == 5 ? 2 : kDbg == 4 ? 1 : 0>(v, tpk, tk, pb, mul32, one);
          } else {  // IAM_UMMA_DEBUG=1: MMA/TMA pipeline only, accumulators dropped
            __syncwarp();
            tc_fence_before();
            if (lane0) mbar_arrive_a(bar + kEmptyOff);
          }
        };
        constexpr int kTileStep = kKeyMul * kParts;           // key-space distance between successive tiles (96 = kBRows for 32-column parts)
        int tpk = part * kKeyMul - phase * kTileStep;          // kKeyMul * (tile * kParts + part) of the pair's first tile
        for (int tb = -phase; tb < n_tb; tb += 2, tpk += 2 * kTileStep) {
          if (!IAM_LAT_EX || kDbg == 1) exchange();
          if (tb >= 0) tile(std::integral_constant<uint32_t, 0>{}, tpk, true);
          if (tb + 1 < n_tb) {
            tile(std::integral_constant<uint32_t, 1>{}, tpk + kTileStep, tb < 0);
            par ^= 1;
          }
          if constexpr (kParts > 1) {
#if IAM_ROW_BOUND
            sts_volatile_v2b32_a(wr, O::bits(tk.d[KTOP - 1]), O::bits(tk.d[0]));
#else
            sts_volatile_b32_a(wr, O::bits(tk.d[KTOP - 1]));
#endif
          }
        }
        phase = (phase + n_tb) & 1;
      } else {
        const int tp_end = n_tb * kParts + part;
        for (int tp = part; tp < tp_end; tp += kParts) {  // tp numbers the 32-column slices of the train image
          // Bound from the threads that own the column parts of this row (own slot included, it is harmless):
          // nothing worse than the best of the k-th bests can end up in the merged list.  Ties are admitted
          // (loosen; the final merge orders them by index).  Stale
          // values are still valid bounds, so plain volatile shared-memory traffic suffices.
          if (kParts > 1) {
            T g = O::from_bits(lds_volatile_b32_a(rd));
  #pragma unroll
            for (int pp = 1; pp < kParts; ++pp) g = O::best(g, O::from_bits(lds_volatile_b32_a(rd + pp * kPartStride)));
            if constexpr (kKeyF) pb = O::best(pb, g + 1.0f);  // distances are integers: the first value NOT admissible
            else pb = O::best(pb, O::loosen(g));
          }
          const uint32_t bar = bar_full0 + slot * 8;
          mbar_wait_bare_a(bar, par);
          tc_fence_after();
          if (kDbg != 1) {
            T v[kPC];
            __syncwarp();
            if constexpr (kPC == 32) {
              tmem_ld32(tm_warp + slot * kBRows, v);
              tmem_ld_wait(v);
            } else {
              tmem_ld_n<kPC>(tm_warp + slot * kBRows, v);
              tmem_ld_wait_n<kPC>(v);
            }
            // the values are in registers: hand the accumulator slot back to the MMA issuer before consuming them
            __syncwarp();
            tc_fence_before();
            if (lane0) mbar_arrive_a(bar + kEmptyOff);
            if (kDbg == 0) {
              if constexpr (kPacked && kPC != 32) consume_packed_n<kPC, 0>(v, tp * kKeyMul, tk, pb, mul32, one);
              else if constexpr (kPacked) consume32_packed<0>(v, tp * 32, tk, pb, mul32, one);
              else if constexpr (kPackedF) consume32_packed_f<0>(v, tp * 32, tk, pb);
              else consume32<KTOP, O, false, kKeyF>(v, tp, tk, pb);
            } else if (kDbg == 5) {  // profiling aid (IAM_UMMA_DEBUG=5): identical slow-path work in every warp and tile
              if constexpr (kPacked && kPC != 32) consume_packed_n<kPC, 2>(v, tp * kKeyMul, tk, pb, mul32, one);
              else if constexpr (kPacked) consume32_packed<2>(v, tp * 32, tk, pb, mul32, one);
              else if constexpr (kPackedF) consume32_packed_f<2>(v, tp * 32, tk, pb);
              else consume32<KTOP, O, true, kKeyF>(v, tp, tk, pb);
            } else if (kDbg == 4) {  // profiling aid (IAM_UMMA_DEBUG=4): tests + votes + branches, never taken
              if constexpr (kPacked && kPC != 32) consume_packed_n<kPC, 1>(v, tp * kKeyMul, tk, pb, mul32, one);
              else if constexpr (kPacked) consume32_packed<1>(v, tp * 32, tk, pb, mul32, one);
              else if constexpr (kPackedF) consume32_packed_f<1>(v, tp * 32, tk, pb);
              else consume32<KTOP, O, false, kKeyF>(v, tp, tk, O::never());
            } else if (kDbg == 3) {  // profiling aid (IAM_UMMA_DEBUG=3): accumulator read-out only, results NOT valid
              tk.d[0] = O::best(tk.d[0], v[0]);  // the load itself is volatile: all 32 columns are still read
            } else {  // profiling aid (IAM_UMMA_DEBUG=2): fast path only, results NOT valid
              T m = v[0];
  #pragma unroll
              for (int j = 0; j < 8; ++j) m = O::best(m, O::best(O::best3(v[j * 4], v[j * 4 + 1], v[j * 4 + 2]), v[j * 4 + 3]));
              tk.d[0] = O::best(tk.d[0], m);
            }
          } else {  // IAM_UMMA_DEBUG=1: MMA/TMA pipeline only, accumulators dropped
            __syncwarp();
            tc_fence_before();
            if (lane0) mbar_arrive_a(bar + kEmptyOff);
          }
          if (kParts > 1) sts_volatile_b32_a(wr, O::bits(tk.d[KTOP - 1]));
          slot += kT;
          if (slot >= kSlots) {
            slot -= kSlots;
            par ^= 1;
          }
        }
      }
      // end of unit: every list becomes (squared distance, train row); parts 1.. hand theirs to part 0's thread
      // of the same row, which merges (distance, index)-lexicographically
      const int row = unit.super * kSuperRows + (w % kItems) * kRows + urow;  // wide layouts: the query row; byte layout: its rank
      Final<KTOP> fin;
      {
        int rowc = 0, n_even = 0;
        if (kKind == Kind::I8) {
          rowc = q.rowc[row];
          n_even = t.meta[kMetaNEven];
        }
#pragma unroll
        for (int s = 0; s < KTOP; ++s) {
          const int enc = tk.i[s];
          if (kKind == Kind::I8) {
            int rank;
            if constexpr (kPacked && kPC != 32)  // e = kKeyMul * slice | (kKeyMul - 1 - column in slice), slices of kPC columns
              rank = enc == 0x7fffffff ? 0x7fffffff : (enc / kKeyMul) * kPC + (kKeyMul - 1 - (enc & (kKeyMul - 1)));
            else
              rank = enc == 0x7fffffff ? 0x7fffffff : (kPacked ? (enc ^ 31) : dec_index(enc));
            if (rank < t.n) {  // d^2 = ||q||^2 + 2 CAP + (||t||^2 & 1) - 2 acc, exact (layout.h)
              fin.d[s] = static_cast<float>(rowc + (rank >= n_even ? 1 : 0) - 2 * static_cast<int>(tk.d[s]));
              fin.i[s] = t.perm[rank];
            } else {
              fin.d[s] = kInf;
              fin.i[s] = -1;
            }
          } else {
            fin.d[s] = static_cast<float>(tk.d[s]);
            fin.i[s] = enc == 0x7fffffff ? -1 : (kPackedF ? enc : dec_index(enc));
          }
        }
      }
      if (part > 0) {
#pragma unroll
        for (int s = 0; s < KTOP; ++s)
          merge[((part - 1) * kRows + urow) * KTOP + s] = make_float2(fin.d[s], __int_as_float(fin.i[s]));
      }
      // only the kParts warps that own these 32 rows meet here: the other row groups run on
      if constexpr (kParts > 1) asm volatile("bar.sync %0, %1;" ::"r"(group_bar), "n"(32 * kParts) : "memory");
      if (part == 0) {
#pragma unroll
        for (int pp = 0; pp < kParts - 1; ++pp) {
#pragma unroll
          for (int s = 0; s < KTOP; ++s) {
            const float2 m = merge[(pp * kRows + urow) * KTOP + s];
            fin.insert_lex(m.x, __float_as_int(m.y));
          }
        }
        if (row < q.n) {
          const int orow = kKind == Kind::I8 ? q.perm[row] : row;
          const size_t o = (static_cast<size_t>(unit.out_base) + orow) * KTOP;
#pragma unroll
          for (int s = 0; s < KTOP; ++s) {
            out_idx[o + s] = fin.i[s];
            out_d2[o + s] = fin.d[s];
          }
        }
      }
      if constexpr (kParts > 1) asm volatile("bar.sync %0, %1;" ::"r"(group_bar), "n"(32 * kParts) : "memory");  // merge area is free again
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kCluster) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it
  if (role == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// Debug aid: one 128x128 accumulator tile (first A tile of `q` x first B tile of `t`)
// with the descriptor strides given at run time, accumulators dumped to global as floats.
// The last K-step of each role sits at aug_a / aug_b (byte offsets) when those are >= 0.
template <Kind kKind>
__global__ void __launch_bounds__(128, 1)
umma_tile_debug_kernel(const uint8_t* __restrict__ a_tile, const uint8_t* __restrict__ b_tile, uint32_t tile_bytes,
                       uint32_t lbo, uint32_t sbo, uint32_t kstep_bytes, int ksteps_in, int aug_a, int aug_b,
                       float* __restrict__ out) {
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
    mbar_arrive_expect_tx(&bar_full, 2 * tile_bytes);
    bulk_g2s(smem, a_tile, tile_bytes, &bar_full);
    bulk_g2s(smem + tile_bytes, b_tile, tile_bytes, &bar_full);
    mbar_wait(&bar_full, 0, 90);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_kind<kKind>(128, 128);
    auto a_off = [&](int ks) { return (ks == ksteps - 1 && aug_a >= 0) ? uint32_t(aug_a) : ks * kstep_bytes; };
    auto b_off = [&](int ks) { return (ks == ksteps - 1 && aug_b >= 0) ? uint32_t(aug_b) : ks * kstep_bytes; };
    if (a_in_tmem)
      for (int ks = 0; ks < ksteps; ++ks)
        tmem_cp_128x256b(tmem + 128 + ks * 8, make_smem_desc(smem_u32(smem) + a_off(ks), lbo, sbo));
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint64_t ad = make_smem_desc(smem_u32(smem) + a_off(ks), lbo, sbo);
      const uint64_t bd = make_smem_desc(smem_u32(smem + tile_bytes) + b_off(ks), lbo, sbo);
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
    typename Ord<kKind>::T v[32];
    __syncwarp();
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait(v);
#pragma unroll
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 128 + c * 32 + j] = static_cast<float>(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem);
  }
}

template <Kind kKind, int KTOP, int kT, int kPC = 32>
cudaError_t launch_shape(const ImgDev* imgs, const KnnUnit* units, int n_units, int* out_idx, float* out_d2, int num_sms,
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
  using C = Cfg<kKind, kT, kPC>;
  using KernT = void (*)(const ImgDev*, const KnnUnit*, int, int*, float*, uint32_t, uint32_t);
  const int n_items = n_units * C::kItems;
  const bool use_cluster = cluster && (n_items % 2 == 0);
  KernT kern;
  if (flags != 0) {  // profiling variants exist for the production configuration only
    if (!use_cluster || !a_tmem || flags < 0 || flags > 5) return cudaErrorInvalidValue;
    kern = flags == 1   ? knn_umma_kernel<kKind, KTOP, true, true, 1, kT, kPC>
           : flags == 2 ? knn_umma_kernel<kKind, KTOP, true, true, 2, kT, kPC>
           : flags == 3 ? knn_umma_kernel<kKind, KTOP, true, true, 3, kT, kPC>
           : flags == 4 ? knn_umma_kernel<kKind, KTOP, true, true, 4, kT, kPC>
                        : knn_umma_kernel<kKind, KTOP, true, true, 5, kT, kPC>;
  } else if (use_cluster) {
    kern = a_tmem ? knn_umma_kernel<kKind, KTOP, true, true, 0, kT, kPC> : knn_umma_kernel<kKind, KTOP, false, true, 0, kT, kPC>;
  } else {
    kern = a_tmem ? knn_umma_kernel<kKind, KTOP, true, false, 0, kT, kPC> : knn_umma_kernel<kKind, KTOP, false, false, 0, kT, kPC>;
  }
  constexpr size_t kSmemTotal = C::kSmemTotal;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal);
  if (err != cudaSuccess) return err;
  if (C::kCtasPerSm > 1) {  // both CTAs of an SM must fit: all of the unified L1 as shared memory
    err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (err != cudaSuccess) return err;
  }
  const int slots = num_sms * C::kCtasPerSm;
  int grid = n_items < slots ? n_items : slots;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(C::kThreads);
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
  return cudaLaunchKernelEx(&cfg, kern, imgs, units, n_units, out_idx, out_d2, static_cast<uint32_t>(C::kKeyMul), 1u);  // the key-packing multipliers, as run-time values
}

template <Kind kKind, int KTOP>
cudaError_t launch_t(const ImgDev* imgs, const KnnUnit* units, int n_units, int* out_idx, float* out_d2, int num_sms,
                     cudaStream_t stream) {
  if constexpr (kKind == Kind::I8) {
    // A/B aid: IAM_UMMA_CTA_TILES=1 runs two half-unit CTAs per SM (Cfg).  Measured slower than one whole-unit CTA
    // (22.1 vs 19.6 ms per 1990 pairs): co-resident CTAs start their units together and stay in phase.
    static const int tiles = [] {
      const char* e = getenv("IAM_UMMA_CTA_TILES");
      return e ? atoi(e) : 2;
    }();
    if (tiles == 1) return launch_shape<kKind, KTOP, 1>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
    if constexpr (KTOP == 2 && IAM_PACKED && IAM_PART_COLS != 32)
      return launch_shape<kKind, KTOP, 2, IAM_PART_COLS>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
  }
  return launch_shape<kKind, KTOP, 2>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
}

template <Kind kKind>
cudaError_t launch_debug_t(const uint8_t* a_tile, const uint8_t* b_tile, uint32_t tile_bytes, uint32_t lbo, uint32_t sbo,
                           uint32_t kstep_bytes, int ksteps, int aug_a, int aug_b, float* out, cudaStream_t stream) {
  const size_t smem = 2 * tile_bytes + 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_tile_debug_kernel<kKind>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  umma_tile_debug_kernel<kKind><<<1, 128, smem, stream>>>(a_tile, b_tile, tile_bytes, lbo, sbo, kstep_bytes, ksteps, aug_a,
                                                          aug_b, out);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_umma_tile_debug(int kind, const uint8_t* a_tile, const uint8_t* b_tile, uint32_t lbo, uint32_t sbo,
                                   uint32_t kstep_bytes, int ksteps, float* out, cudaStream_t stream) {
  if (kind == kKindF16) return launch_debug_t<Kind::F16>(a_tile, b_tile, kTileBytes, lbo, sbo, kstep_bytes, ksteps, -1, -1, out, stream);
  if (kind == kKindF8) return launch_debug_t<Kind::F8>(a_tile, b_tile, kTileBytes, lbo, sbo, kstep_bytes, ksteps, -1, -1, out, stream);
  if (kind == kKindI8) {
    using L = LayD<Kind::I8>;
    const int ks = ksteps < 0 ? -L::kKSteps : L::kKSteps;  // only the sign (SS / TS form) is taken from the caller
    return launch_debug_t<Kind::I8>(a_tile, b_tile, L::kTileBytes, kLBO, L::kSBO, kKStepBytes, ks, L::kAugA, L::kAugB, out, stream);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_knn_umma(int kind, int k, const ImgDev* imgs, const KnnUnit* units, int n_units, int* out_idx,
                            float* out_d2, int num_sms, cudaStream_t stream) {
  if (n_units <= 0) return cudaSuccess;
  if (kind == kKindF16) {
    if (k == 1) return launch_t<Kind::F16, 1>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
    if (k == 2) return launch_t<Kind::F16, 2>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
    if (k == 3) return launch_t<Kind::F16, 3>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
  } else if (kind == kKindF8) {
    if (k == 1) return launch_t<Kind::F8, 1>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
    if (k == 2) return launch_t<Kind::F8, 2>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
    if (k == 3) return launch_t<Kind::F8, 3>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
  } else if (kind == kKindI8) {
    if (k == 1) return launch_t<Kind::I8, 1>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
    if (k == 2) return launch_t<Kind::I8, 2>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
    if (k == 3) return launch_t<Kind::I8, 3>(imgs, units, n_units, out_idx, out_d2, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace iam
