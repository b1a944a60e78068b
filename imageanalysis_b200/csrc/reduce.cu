// reduce.cu — the per-direction ratio/metric reduction and the reciprocal
// cross-check that follow knnMatch in the reference's 'traditional' strategy:
//   basic_pair_matches   scripts/lib/matcher.py:247-273
//   filter_cross_check   scripts/lib/matcher.py:187-200
// The reference does this arithmetic on Python floats (IEEE double) over
// float32 distances, so the device code uses __ddiv_rn/__dmul_rn on the
// widened values to reproduce every rounding, and orders by (metric, queryIdx)
// which is what Python's stable sorted() yields.
#include <cuda_runtime.h>

#include "reduce.h"

#include "layout.h"

namespace iam {
namespace {

constexpr int kRedThreads = 1024;
constexpr int kRankTile = 1024;

__global__ void __launch_bounds__(kRedThreads)
metric_reduce_kernel(const RedJob* __restrict__ jobs, const int* __restrict__ knn_idx, const float* __restrict__ knn_dist,
              int k, ReduceParams prm, double* __restrict__ cand_metric, int2* __restrict__ cand_qt, int cand_stride,
              int* __restrict__ job_table, int* __restrict__ job_count, int sort_cap, int raw_d2) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ int s_count;
  __shared__ double s_m[kRankTile];
  __shared__ int s_q[kRankTile];

  const int job = blockIdx.x;
  const RedJob jb = jobs[job];
  double* cm = cand_metric + static_cast<size_t>(job) * cand_stride;
  int2* cq = cand_qt + static_cast<size_t>(job) * cand_stride;
  int* table = job_table + static_cast<size_t>(job) * prm.cap * 2;

  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();

  for (int r = threadIdx.x; r < jb.n_q; r += blockDim.x) {
    const size_t o = (static_cast<size_t>(jb.out_base) + r) * k;
    const int i0 = knn_idx[o], i1 = knn_idx[o + 1];
    if (i0 < 0 || i1 < 0) continue;
    float f0 = knn_dist[o], f1 = knn_dist[o + 1];
    if (raw_d2 != 0) {
      // the lists still hold what the kNN kernel wrote (squared L2 distance / Hamming distance): the finishing pass
      // (padding rows sit at >= 2^24 / >= 257; L2: correctly rounded float sqrt, as cv2 reports it) is folded in here
      const float pad_thr = raw_d2 == 1 ? 16777216.0f : 257.0f;
      if (f0 >= pad_thr || f1 >= pad_thr) continue;
      if (raw_d2 == 1) {
        f0 = sqrtf(f0);
        f1 = sqrtf(f1);
      }
    }
    const double d0 = static_cast<double>(f0);
    const double d1 = static_cast<double>(f1);
    bool keep;
    double metric;
    if (prm.mode == 0) {  // plain Lowe gate, matcher.py:227
      keep = d0 <= __dmul_rn(d1, prm.ratio);
      metric = 0.0;       // table keeps query order
    } else {              // matcher.py:255-261
      if (d1 == 0.0) continue;  // the reference would raise ZeroDivisionError here; we drop the row
      const double ratio = __ddiv_rn(d0, d1);
      metric = __dmul_rn(d0, ratio);
      keep = metric < prm.thresh;
    }
    if (keep) {
      const int pos = atomicAdd(&s_count, 1);
      cm[pos] = metric;
      cq[pos] = make_int2(r, i0);
    }
  }
  __syncthreads();
  const int c = s_count;
  if (c < prm.min_pairs) {  // matcher.py:271-273
    if (threadIdx.x == 0) job_count[job] = 0;
    return;
  }
  __threadfence_block();

  // Order the candidates by (metric, queryIdx) = what Python's stable sorted() yields (matcher.py:258).
  // Up to kSortMax candidates: bitonic sort in shared memory on the raw IEEE bits (positive doubles
  // order like unsigned integers); beyond that, rank every candidate by counting.
  int npow2 = 1;
  while (npow2 < c) npow2 <<= 1;
  if (npow2 <= sort_cap) {
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(s_dyn);
    int2* s_qt = reinterpret_cast<int2*>(s_key + npow2);
    for (int e = threadIdx.x; e < npow2; e += blockDim.x) {
      if (e < c) {
        s_key[e] = static_cast<unsigned long long>(__double_as_longlong(cm[e]));
        s_qt[e] = cq[e];
      } else {
        s_key[e] = ~0ull;
        s_qt[e] = make_int2(0x7fffffff, 0);
      }
    }
    __syncthreads();
    // a warp's 32 consecutive elements exchange among themselves while the stride is below 32: runs of such steps
    // need only warp-level synchronisation (34 of the 66 steps at 2048 elements)
    for (int k2 = 2; k2 <= npow2; k2 <<= 1) {
      for (int j = k2 >> 1; j > 0; j >>= 1) {
        for (int e = threadIdx.x; e < npow2; e += blockDim.x) {
          const int x = e ^ j;
          if (x > e) {
            const unsigned long long ka = s_key[e], kb = s_key[x];
            const int2 qa = s_qt[e], qb = s_qt[x];
            const bool a_gt_b = ka > kb || (ka == kb && qa.x > qb.x);
            const bool up = (e & k2) == 0;
            if (a_gt_b == up) {
              s_key[e] = kb;
              s_key[x] = ka;
              s_qt[e] = qb;
              s_qt[x] = qa;
            }
          }
        }
        // block-wide barrier when this step or the next one exchanges across warps; otherwise the warp's own
        const int next_j = j > 1 ? (j >> 1) : k2;   // first stride of the next merge size: (2 k2) / 2
        if (j >= 32 || next_j >= 32 || (k2 == npow2 && j == 1)) __syncthreads();
        else __syncwarp();
      }
    }
    const int n_out = min(c, prm.cap);  // matcher.py:265-269
    for (int e = threadIdx.x; e < n_out; e += blockDim.x) {
      table[e * 2 + 0] = s_qt[e].x;
      table[e * 2 + 1] = s_qt[e].y;
    }
  } else {
  const int n_slots = (c + blockDim.x - 1) / blockDim.x;
  for (int slot = 0; slot < n_slots; ++slot) {
    const int e = slot * blockDim.x + threadIdx.x;
    double me = 0.0;
    int2 qe = make_int2(0, 0);
    if (e < c) {
      me = cm[e];
      qe = cq[e];
    }
    int rank = 0;
    for (int f0 = 0; f0 < c; f0 += kRankTile) {
      __syncthreads();
      const int f = f0 + threadIdx.x;
      if (threadIdx.x < kRankTile && f < c) {
        s_m[threadIdx.x] = cm[f];
        s_q[threadIdx.x] = cq[f].x;
      }
      __syncthreads();
      const int lim = min(kRankTile, c - f0);
      if (e < c) {
        for (int j = 0; j < lim; ++j) {
          const double mf = s_m[j];
          rank += (mf < me) || (mf == me && s_q[j] < qe.x);
        }
      }
    }
    if (e < c && rank < prm.cap) {  // matcher.py:265-269
      table[rank * 2 + 0] = qe.x;
      table[rank * 2 + 1] = qe.y;
    }
  }
  }
  if (threadIdx.x == 0) job_count[job] = min(c, prm.cap);
}

// filter_duplicates, matcher.py:157-182: walk the table in order; an entry is
// dropped when the query keypoint's position key or the train keypoint's
// position key was already claimed by an earlier KEPT entry.  The walk is
// inherently sequential (only kept entries claim keys), so the block gathers
// keys in parallel and lane 0 does the <= cap-step scan out of shared memory.
__global__ void __launch_bounds__(256)
dedupe_kernel(const RedJob* __restrict__ jobs, const ImgDev* __restrict__ imgs, int cap, int min_pairs, int words,
              int* __restrict__ job_table, int* __restrict__ job_count) {
  extern __shared__ int s_mem[];
  int* s_q = s_mem;
  int* s_t = s_q + cap;
  int* s_k1 = s_t + cap;
  int* s_k2 = s_k1 + cap;
  unsigned* s_u1 = reinterpret_cast<unsigned*>(s_k2 + cap);
  unsigned* s_u2 = s_u1 + words;
  __shared__ int s_new;

  const int job = blockIdx.x;
  const int count = job_count[job];
  if (count == 0) return;
  const RedJob jb = jobs[job];
  const int* kq = imgs[jb.q_slot].kp_key;
  const int* kt = imgs[jb.t_slot].kp_key;
  int* table = job_table + static_cast<size_t>(job) * cap * 2;
  for (int e = threadIdx.x; e < count; e += blockDim.x) {
    const int q = table[e * 2], t = table[e * 2 + 1];
    s_q[e] = q;
    s_t[e] = t;
    s_k1[e] = kq ? kq[q] : q;
    s_k2[e] = kt ? kt[t] : t;
  }
  for (int w = threadIdx.x; w < 2 * words; w += blockDim.x) s_u1[w] = 0u;
  __syncthreads();
  if (threadIdx.x == 0) {
    int out = 0;
    for (int e = 0; e < count; ++e) {
      const int a = s_k1[e], b = s_k2[e];
      const bool used = ((s_u1[a >> 5] >> (a & 31)) & 1u) || ((s_u2[b >> 5] >> (b & 31)) & 1u);
      if (!used) {
        s_u1[a >> 5] |= 1u << (a & 31);
        s_u2[b >> 5] |= 1u << (b & 31);
        s_k1[out] = e;  // safe: out <= e, and entry e's keys were already consumed
        ++out;
      }
    }
    s_new = out;
  }
  __syncthreads();
  const int n_new = s_new;
  for (int o = threadIdx.x; o < n_new; o += blockDim.x) {
    const int e = s_k1[o];
    table[o * 2] = s_q[e];
    table[o * 2 + 1] = s_t[e];
  }
  if (threadIdx.x == 0) job_count[job] = (n_new < min_pairs) ? 0 : n_new;  // matcher.py:296-298
}

// One CTA per pair.  fwd = job 2p, rev = job 2p+1.
__global__ void __launch_bounds__(256)
crosscheck_kernel(const RedJob* __restrict__ jobs, const int* __restrict__ job_table,
                  const int* __restrict__ job_count, int cap, int cross_check, int* __restrict__ out_table,
                  int* __restrict__ out_count) {
  extern __shared__ int s_rev[];  // [n_t] reverse lookup: rev_match[t] = q
  __shared__ int s_scan[256];

  const int p = blockIdx.x;
  const int* ft = job_table + static_cast<size_t>(2 * p) * cap * 2;
  const int* rt = job_table + static_cast<size_t>(2 * p + 1) * cap * 2;
  const int fc = job_count[2 * p];
  const int rc = job_count[2 * p + 1];
  int* ot = out_table + static_cast<size_t>(p) * cap * 2;
  const int n_t = jobs[2 * p].n_t;

  if (!cross_check) {
    for (int e = threadIdx.x; e < fc * 2; e += blockDim.x) ot[e] = ft[e];
    if (threadIdx.x == 0) out_count[p] = fc;
    return;
  }
  for (int t = threadIdx.x; t < n_t; t += blockDim.x) s_rev[t] = -1;
  __syncthreads();
  // the reference only runs the reverse match when the forward one survived (matcher.py:312-316)
  const int rc_eff = (fc > 0) ? rc : 0;
  for (int e = threadIdx.x; e < rc_eff; e += blockDim.x) s_rev[rt[e * 2]] = rt[e * 2 + 1];
  __syncthreads();

  // order-preserving compaction of the forward table
  const int per = (fc + blockDim.x - 1) / blockDim.x;
  const int b = threadIdx.x * per;
  int local = 0;
  for (int e = b; e < min(b + per, fc); ++e) local += (s_rev[ft[e * 2 + 1]] == ft[e * 2]);
  s_scan[threadIdx.x] = local;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {
    int v = 0;
    if (threadIdx.x >= off) v = s_scan[threadIdx.x - off];
    __syncthreads();
    s_scan[threadIdx.x] += v;
    __syncthreads();
  }
  int pos = s_scan[threadIdx.x] - local;
  for (int e = b; e < min(b + per, fc); ++e) {
    const int q = ft[e * 2], t = ft[e * 2 + 1];
    if (s_rev[t] == q) {
      ot[pos * 2] = q;
      ot[pos * 2 + 1] = t;
      ++pos;
    }
  }
  if (threadIdx.x == 255) out_count[p] = s_scan[255];
}


// Compact form of the per-pair tables for the multi-GPU gather ("counts first, then the payload"):
// offsets[p] = exclusive prefix sum of the counts (one CTA walks the list with a running carry), rows = the valid
// rows of every pair back to back, in pair order.
__global__ void __launch_bounds__(1024)
scan_counts_kernel(const int* __restrict__ count, int n, int* __restrict__ offsets) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < n ? count[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int carry = s_carry;
    const int incl = carry + x + (warp > 0 ? s_warp[warp - 1] : 0);
    if (i < n) offsets[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = s_carry;
}

__global__ void __launch_bounds__(256)
pack_tables_kernel(const int2* __restrict__ table, const int* __restrict__ count, const int* __restrict__ offsets, int cap,
                   int2* __restrict__ rows) {
  const int p = blockIdx.x;
  const int c = count[p];
  const int2* src = table + static_cast<size_t>(p) * cap;
  int2* dst = rows + offsets[p];
  for (int e = threadIdx.x; e < c; e += blockDim.x) dst[e] = src[e];
}

}  // namespace

cudaError_t launch_scan_counts(const int* count, int n, int* offsets, cudaStream_t stream) {
  scan_counts_kernel<<<1, 1024, 0, stream>>>(count, n, offsets);
  return cudaGetLastError();
}

cudaError_t launch_pack_tables(const int* table, const int* count, const int* offsets, int n_pairs, int cap, int* rows,
                               cudaStream_t stream) {
  if (n_pairs <= 0) return cudaSuccess;
  pack_tables_kernel<<<n_pairs, 256, 0, stream>>>(reinterpret_cast<const int2*>(table), count, offsets, cap,
                                                  reinterpret_cast<int2*>(rows));
  return cudaGetLastError();
}

cudaError_t launch_reduce(const RedJob* jobs, int n_jobs, const int* knn_idx, const float* knn_dist, int k,
                          const ReduceParams& prm, double* cand_metric, int2* cand_qt, int cand_stride,
                          int* job_table, int* job_count, int raw_d2, cudaStream_t stream) {
  if (n_jobs <= 0) return cudaSuccess;
  // shared-memory sort capacity: next power of two of the largest possible candidate count, at most 8192
  int sort_cap = 1;
  while (sort_cap < cand_stride && sort_cap < 8192) sort_cap <<= 1;
  const size_t smem = static_cast<size_t>(sort_cap) * 16;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(metric_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  metric_reduce_kernel<<<n_jobs, kRedThreads, smem, stream>>>(jobs, knn_idx, knn_dist, k, prm, cand_metric, cand_qt,
                                                           cand_stride, job_table, job_count, sort_cap, raw_d2);
  return cudaGetLastError();
}

cudaError_t launch_dedupe(const RedJob* jobs, int n_jobs, const ImgDev* imgs, int cap, int min_pairs, int max_n,
                          int* job_table, int* job_count, cudaStream_t stream) {
  if (n_jobs <= 0) return cudaSuccess;
  const int words = (max_n + 31) / 32 + 1;
  const size_t smem = (static_cast<size_t>(cap) * 4 + static_cast<size_t>(words) * 2) * sizeof(int);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(dedupe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dedupe_kernel<<<n_jobs, 256, smem, stream>>>(jobs, imgs, cap, min_pairs, words, job_table, job_count);
  return cudaGetLastError();
}

cudaError_t launch_crosscheck(const RedJob* jobs, int n_pairs, const int* job_table, const int* job_count, int cap,
                              int cross_check, int max_n_t, int* out_table, int* out_count, cudaStream_t stream) {
  if (n_pairs <= 0) return cudaSuccess;
  const size_t smem = static_cast<size_t>(max_n_t) * sizeof(int);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(crosscheck_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  crosscheck_kernel<<<n_pairs, 256, smem, stream>>>(jobs, job_table, job_count, cap, cross_check, out_table,
                                                    out_count);
  return cudaGetLastError();
}

}  // namespace iam
