// reduce.h — launcher declarations for the post-kNN reduction kernels.
#pragma once
#include <cuda_runtime.h>

namespace iam {

// One directed job (query image -> train image) as the reduction kernels see it.
struct RedJob {
  int out_base;  // first row of this job in the kNN output arrays
  int n_q;       // valid query descriptors
  int n_t;       // valid train descriptors
  int q_slot;    // image slots (for the keypoint keys of filter_duplicates)
  int t_slot;
  int pad[3];
};

struct ReduceParams {
  double ratio;   // match_ratio
  double thresh;  // max_distance * match_ratio, evaluated by the host in double like Python does
  int mode;       // 0 = Lowe, 1 = reference metric
  int cap;
  int min_pairs;
  int pad;
};

// raw_d2: 0 = knn_dist holds finished distances; 1 = squared L2 distances, 2 = Hamming distances straight from the
// kNN kernel (the finishing pass -- sqrt, padding rows -- is folded into the reduction: no extra sweep over the lists)
cudaError_t launch_reduce(const RedJob* jobs, int n_jobs, const int* knn_idx, const float* knn_dist, int k,
                          const ReduceParams& prm, double* cand_metric, int2* cand_qt, int cand_stride,
                          int* job_table, int* job_count, int raw_d2, cudaStream_t stream);

// filter_duplicates (matcher.py:157-182) + the min_pairs gate that follows it (:296-298),
// in place on the per-job tables.  `imgs` supplies the per-image keypoint key arrays.
struct ImgDev;
cudaError_t launch_dedupe(const RedJob* jobs, int n_jobs, const ImgDev* imgs, int cap, int min_pairs, int max_n,
                          int* job_table, int* job_count, cudaStream_t stream);

cudaError_t launch_crosscheck(const RedJob* jobs, int n_pairs, const int* job_table, const int* job_count, int cap,
                              int cross_check, int max_n_t, int* out_table, int* out_count, cudaStream_t stream);

// Compact (CSR) form of the per-pair tables: offsets [n+1] = exclusive prefix sums of count [n]; rows = the valid
// rows of every pair back to back in pair order (what the multi-GPU gather moves instead of the padded tables).
cudaError_t launch_scan_counts(const int* count, int n, int* offsets, cudaStream_t stream);
cudaError_t launch_pack_tables(const int* table, const int* count, const int* offsets, int n_pairs, int cap, int* rows,
                               cudaStream_t stream);

}  // namespace iam
