// gms.cu — grid-based motion statistics filter on the per-direction match tables,
// replacing cv2.xfeatures2d.matchGMS as called from basic_pair_matches()
// (reference scripts/lib/matcher.py:285; algorithm as restated by the reference
// itself in scripts/lib/archive/gms_matcher.py:74-285).
//
// One CTA per directed job (<= 4096 matches).  The reference's 400 x G dense
// motion-statistics matrix holds at most one non-zero per match, so it lives in
// shared memory as an open-addressing hash (key = left cell * 2048 + right cell);
// the per-left-cell arg-max (first maximum, like the strict '>' scan of
// gms_matcher.py:261-265) is an atomicMax over (count << 11 | 2047 - right cell).
// The statistics do not depend on the rotation hypothesis, so each of the four
// half-cell shifted grids is built once and verified against all 8 rotations;
// every match carries one inlier bit per rotation, OR-ed over the four grids
// exactly as GmsMatcher.run() accumulates its mask (:187-209).
#include <cuda_runtime.h>

#include "gms.h"
#include "layout.h"
#include "reduce.h"

namespace iam {
namespace {

constexpr int kGrid = 20;                    // gms_matcher.py:83
constexpr int kLeftCells = kGrid * kGrid;
constexpr int kMaxRightCells = 1600;         // scale 2.0 -> 40 x 40
constexpr int kEmpty = -1;

// which right-cell neighbour faces left-cell neighbour j under each rotation (gms_matcher.py:29-60, 0-based)
__constant__ int8_t c_rot[8][9] = {{0, 1, 2, 3, 4, 5, 6, 7, 8}, {3, 0, 1, 6, 4, 2, 7, 8, 5}, {6, 3, 0, 7, 4, 1, 8, 5, 2},
                                   {7, 6, 3, 8, 4, 0, 5, 2, 1}, {8, 7, 6, 5, 4, 3, 2, 1, 0}, {5, 8, 7, 2, 4, 6, 1, 0, 3},
                                   {2, 5, 8, 1, 4, 7, 0, 3, 6}, {1, 2, 5, 0, 4, 8, 3, 6, 7}};

__device__ __forceinline__ int grid_w_right(int scale) {  // int(20 * ratio), gms_matcher.py:63,178-181
  const int w[5] = {20, 10, 14, 28, 40};
  return w[scale];
}

__device__ __forceinline__ uint32_t hash_slot(int key, uint32_t mask) {
  return (static_cast<uint32_t>(key) * 2654435761u >> 12) & mask;
}

__device__ __forceinline__ int neighbour(int cell, int s, int gw, int gh) {  // get_nb9, gms_matcher.py:112-127
  const int x = cell % gw + (s % 3 - 1), y = cell / gw + (s / 3 - 1);
  return (x < 0 || x >= gw || y < 0 || y >= gh) ? -1 : x + y * gw;
}

struct Smem {
  int2* pairs;       // [n]   the job's matches (query, train), original order
  short* lcell;      // [4][n] left cell per shifted grid, -1 outside
  short* rcell;      // [n]
  int* hkey;         // [slots]
  int* hcnt;         // [slots]
  int* per_left;     // [400] matches per left cell
  int* best;         // [400] count << 11 | (2047 - right cell)
  short* pair;       // [8][400] accepted right cell per rotation, -1 none, -2 rejected
  uint8_t* bits;     // [n] inlier bit per rotation (current scale)
  uint8_t* keep;     // [n] mask of the best hypothesis so far
};

__device__ __forceinline__ int hash_lookup(const Smem& s, uint32_t mask, int key) {
  uint32_t h = hash_slot(key, mask);
  while (true) {
    const int k = s.hkey[h];
    if (k == key) return s.hcnt[h];
    if (k == kEmpty) return 0;
    h = (h + 1) & mask;
  }
}

__global__ void __launch_bounds__(256)
gms_kernel(const RedJob* __restrict__ jobs, const ImgDev* __restrict__ imgs, GmsParams prm, int cap, int n_max, int slots,
           int* __restrict__ job_table, int* __restrict__ job_count) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  __shared__ int s_cnt[8];
  __shared__ int s_best, s_best_rot, s_total;
  __shared__ int s_scan[256];

  const int job = blockIdx.x;
  const int n = job_count[job];
  if (n == 0) return;
  const RedJob jb = jobs[job];
  const float2* kp1 = imgs[jb.q_slot].kp_xy;
  const float2* kp2 = imgs[jb.t_slot].kp_xy;
  int* table = job_table + static_cast<size_t>(job) * cap * 2;

  Smem s;
  uint8_t* p = s_raw;
  s.pairs = reinterpret_cast<int2*>(p);      p += static_cast<size_t>(n_max) * 8;
  s.hkey = reinterpret_cast<int*>(p);        p += static_cast<size_t>(slots) * 4;
  s.hcnt = reinterpret_cast<int*>(p);        p += static_cast<size_t>(slots) * 4;
  s.per_left = reinterpret_cast<int*>(p);    p += kLeftCells * 4;
  s.best = reinterpret_cast<int*>(p);        p += kLeftCells * 4;
  s.lcell = reinterpret_cast<short*>(p);     p += static_cast<size_t>(n_max) * 8;
  s.rcell = reinterpret_cast<short*>(p);     p += static_cast<size_t>(n_max) * 2;
  s.pair = reinterpret_cast<short*>(p);      p += 8 * kLeftCells * 2;
  s.bits = p;                                p += n_max;
  s.keep = p;
  const uint32_t hmask = static_cast<uint32_t>(slots - 1);
  const int tid = threadIdx.x;

  // left cells of the four shifted grids (GetGridIndexLeft, gms_matcher.py:228-246); coordinates are normalised in
  // double like the Python floats of NormalizePoints (:91-97)
  for (int m = tid; m < n; m += blockDim.x) {
    const int2 qt = make_int2(table[2 * m], table[2 * m + 1]);
    s.pairs[m] = qt;
    const float2 a = kp1[qt.x];
    const double nx = static_cast<double>(a.x) / prm.width, ny = static_cast<double>(a.y) / prm.height;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int x = static_cast<int>(floor(nx * kGrid + ((g & 1) ? 0.5 : 0.0)));
      const int y = static_cast<int>(floor(ny * kGrid + ((g & 2) ? 0.5 : 0.0)));
      // as the reference: only the upper edges are tested (:243-245); a negative coordinate aliases into the
      // neighbouring row of cells, and a negative index drops the match (:222-223)
      const int l = (x >= kGrid || y >= kGrid) ? -1 : x + y * kGrid;
      s.lcell[g * n_max + m] = static_cast<short>(l < 0 ? -1 : l);
    }
    s.keep[m] = 0;
  }
  if (tid == 0) {
    s_best = 0;
    s_best_rot = -1;
  }
  const int n_scales = prm.with_scale ? 5 : 1;
  const int n_rot = prm.with_rotation ? 8 : 1;
  for (int sc = 0; sc < n_scales; ++sc) {
    const int gw = grid_w_right(sc), gh = gw;
    __syncthreads();
    for (int m = tid; m < n; m += blockDim.x) {  // GetGridIndexRight, gms_matcher.py:248-251
      const float2 b = kp2[s.pairs[m].y];
      const int x = static_cast<int>(floor(static_cast<double>(b.x) / prm.width * gw));
      const int y = static_cast<int>(floor(static_cast<double>(b.y) / prm.height * gh));
      // No range test in the reference (:248-251): out-of-frame points alias into other cells.  A negative
      // index is skipped by the statistics (:222-223) but still COMPARED when inliers are marked (:205-207),
      // where -1 / -2 equal the "no pair" / "rejected" marks of a cell: the true value is kept for that.
      const int r = x + y * gw;
      s.rcell[m] = static_cast<short>(r >= gw * gh ? -32768 : (r < -32767 ? -32767 : r));
      s.bits[m] = 0;
    }
    for (int g = 0; g < 4; ++g) {
      const short* lc = s.lcell + g * n_max;
      __syncthreads();
      for (int i = tid; i < slots; i += blockDim.x) s.hkey[i] = kEmpty;
      for (int i = tid; i < kLeftCells; i += blockDim.x) {
        s.per_left[i] = 0;
        s.best[i] = -1;
      }
      __syncthreads();
      // AssignMatchPairs (:211-226): sparse motion statistics
      for (int m = tid; m < n; m += blockDim.x) {
        const int l = lc[m], r = s.rcell[m];
        if (l < 0 || r < 0) continue;
        const int key = l * 2048 + r;
        uint32_t h = hash_slot(key, hmask);
        while (true) {
          const int prev = atomicCAS(&s.hkey[h], kEmpty, key);
          if (prev == kEmpty) s.hcnt[h] = 0;   // the claimer zeroes; adds follow the barrier below
          if (prev == kEmpty || prev == key) break;
          h = (h + 1) & hmask;
        }
        atomicAdd(&s.per_left[l], 1);
      }
      __syncthreads();
      for (int m = tid; m < n; m += blockDim.x) {
        const int l = lc[m], r = s.rcell[m];
        if (l < 0 || r < 0) continue;
        const int key = l * 2048 + r;
        uint32_t h = hash_slot(key, hmask);
        while (s.hkey[h] != key) h = (h + 1) & hmask;
        atomicAdd(&s.hcnt[h], 1);
      }
      __syncthreads();
      for (int i = tid; i < slots; i += blockDim.x) {  // first maximum of every left cell's row (:261-265)
        const int key = s.hkey[i];
        if (key != kEmpty) atomicMax(&s.best[key >> 11], (s.hcnt[i] << 11) | (2047 - (key & 2047)));
      }
      __syncthreads();
      // VerifyCellPairs (:253-285) for every rotation
      for (int w = tid; w < n_rot * kLeftCells; w += blockDim.x) {
        const int rot = w / kLeftCells, i = w - rot * kLeftCells;
        const int b = s.best[i];
        int res = -1;
        if (b >= 0) {
          const int j = 2047 - (b & 2047);
          int score = 0, thresh = 0, numpair = 0;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int ll = neighbour(i, t, kGrid, kGrid);
            const int rr = neighbour(j, c_rot[rot][t], gw, gh);
            if (ll < 0 || rr < 0) continue;
            score += hash_lookup(s, hmask, ll * 2048 + rr);
            thresh += s.per_left[ll];
            ++numpair;
          }
          const double th = prm.threshold_factor * sqrt(static_cast<double>(thresh) / static_cast<double>(numpair));
          res = (static_cast<double>(score) < th) ? -2 : j;
        }
        s.pair[rot * kLeftCells + i] = static_cast<short>(res);
      }
      __syncthreads();
      for (int m = tid; m < n; m += blockDim.x) {  // mark inliers (:203-207), one bit per rotation
        int l = lc[m];
        if (l < 0) {
          if (!prm.archive_wrap) continue;   // OpenCV's C++ skips the match for this grid
          l = kLeftCells - 1;                // the archive Python reads mCellPairs[-1]: the last cell
        }
        const int r = s.rcell[m];
        uint32_t bits = s.bits[m];
        for (int rot = 0; rot < n_rot; ++rot)
          if (s.pair[rot * kLeftCells + l] == r) bits |= 1u << rot;
        s.bits[m] = static_cast<uint8_t>(bits);
      }
    }
    __syncthreads();
    // inliers per rotation; the first hypothesis that beats the best so far wins (GetInlierMask :129-176)
    if (tid < 8) s_cnt[tid] = 0;
    __syncthreads();
    {
      int local[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int m = tid; m < n; m += blockDim.x) {
        const uint32_t b = s.bits[m];
#pragma unroll
        for (int rot = 0; rot < 8; ++rot) local[rot] += (b >> rot) & 1u;
      }
#pragma unroll
      for (int rot = 0; rot < 8; ++rot) {
        int v = local[rot];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0 && v) atomicAdd(&s_cnt[rot], v);
      }
    }
    __syncthreads();
    if (tid == 0) {
      s_best_rot = -1;
      for (int rot = 0; rot < n_rot; ++rot)
        if (s_cnt[rot] > s_best) {
          s_best = s_cnt[rot];
          s_best_rot = rot;
        }
    }
    __syncthreads();
    const int br = s_best_rot;
    if (br >= 0)
      for (int m = tid; m < n; m += blockDim.x) s.keep[m] = (s.bits[m] >> br) & 1u;
  }
  __syncthreads();
  // order-preserving compaction of the table
  const int per = (n + blockDim.x - 1) / blockDim.x;
  const int m0 = tid * per, m1 = min(n, m0 + per);
  int mine = 0;
  for (int m = m0; m < m1; ++m) mine += s.keep[m];
  s_scan[tid] = mine;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int i = 0; i < static_cast<int>(blockDim.x); ++i) {
      const int v = s_scan[i];
      s_scan[i] = run;
      run += v;
    }
    s_total = run;
  }
  __syncthreads();
  int o = s_scan[tid];
  for (int m = m0; m < m1; ++m)
    if (s.keep[m]) {
      table[2 * o] = s.pairs[m].x;
      table[2 * o + 1] = s.pairs[m].y;
      ++o;
    }
  if (tid == 0) job_count[job] = (prm.gate_min_pairs > 0 && s_total < prm.gate_min_pairs) ? 0 : s_total;
}

}  // namespace

size_t gms_smem_bytes(int n_max, int* slots_out) {
  int slots = 1024;
  while (slots < 2 * n_max) slots *= 2;
  if (slots_out) *slots_out = slots;
  return static_cast<size_t>(n_max) * (8 + 8 + 2 + 1 + 1) + static_cast<size_t>(slots) * 8 + kLeftCells * 8 +
         8 * kLeftCells * 2 + 64;
}

cudaError_t launch_gms(const RedJob* jobs, int n_jobs, const ImgDev* imgs, const GmsParams& prm, int cap,
                       int* job_table, int* job_count, cudaStream_t stream) {
  if (n_jobs <= 0) return cudaSuccess;
  if (cap > kGmsMaxMatches || prm.width <= 0 || prm.height <= 0) return cudaErrorInvalidValue;
  int slots = 0;
  const int n_max = (cap + 7) / 8 * 8;
  const size_t smem = gms_smem_bytes(n_max, &slots);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(gms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  gms_kernel<<<n_jobs, 256, smem, stream>>>(jobs, imgs, prm, cap, n_max, slots, job_table, job_count);
  return cudaGetLastError();
}

}  // namespace iam
