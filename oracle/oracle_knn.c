/*
 * oracle_knn.c — CPU restatement of the brute-force k-nearest-neighbour
 * search the reference runs through OpenCV.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's shared object, and only as the
 * checker / the timed CPU baseline.  The product path (imageanalysis_b200/)
 * never links or calls it.
 *
 * What it restates: `the_matcher.knnMatch(des1, des2, k)` as issued by
 * raw_matches(), reference scripts/lib/matcher.py:203-216, with the exact
 * (brute force) matcher the reference itself uses elsewhere
 * (scripts/lib/find_obj.py:46 `cv2.BFMatcher(norm)`; SURVEY.md D1 explains
 * why FLANN, matcher.py:62-79, cannot be a parity oracle).  The arithmetic
 * lives in OpenCV (features2d BFMatcher::knnMatchImpl -> batchDistance),
 * a dependency not vendored under /root/reference (environment.yml:13 pins
 * opencv 4.0.1; this container has 4.13.0).  Its published behaviour, which
 * tests/test_oracle_golden.py pins against live cv2 output stored in
 * tests/golden/, is:
 *   - NORM_L2     : dist = (float) sqrt( sum_k (q_k - t_k)^2 )
 *   - NORM_HAMMING: dist = (float) popcount( q xor t )
 *   - per query the k smallest distances in ascending order; equal distances
 *     are reported in ascending trainIdx order.
 * Threading: the functions are single-threaded and re-entrant; oracle/oracle.py
 * fans row blocks out over a thread pool (ctypes drops the GIL).
 * Parity status: PINNED against cv2.BFMatcher golden vectors (not against a
 * reference-repo test, the reference has none: SURVEY.md section 4).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void insert_topk(int k, int64_t d, int j, int64_t* bd, int32_t* bi) {
  /* strict '<' keeps the earlier train index ahead among equal distances */
  if (d >= bd[k - 1]) return;
  int s = k - 1;
  while (s > 0 && d < bd[s - 1]) {
    bd[s] = bd[s - 1];
    bi[s] = bi[s - 1];
    --s;
  }
  bd[s] = d;
  bi[s] = j;
}

/* L2 on uint8 rows (integer-valued SIFT): exact integer squared distance. */
void oracle_knn_l2_u8(const uint8_t* q, int nq, const uint8_t* t, int nt, int dim, int k, int32_t* idx, float* dist) {
  for (int i = 0; i < nq; ++i) {
    int64_t bd[8];
    int32_t bi[8];
    for (int s = 0; s < k; ++s) {
      bd[s] = INT64_MAX;
      bi[s] = -1;
    }
    const uint8_t* a = q + (size_t)i * dim;
    for (int j = 0; j < nt; ++j) {
      const uint8_t* b = t + (size_t)j * dim;
      int32_t acc = 0;
      for (int c = 0; c < dim; ++c) {
        const int32_t e = (int32_t)a[c] - (int32_t)b[c];
        acc += e * e;
      }
      insert_topk(k, acc, j, bd, bi);
    }
    for (int s = 0; s < k; ++s) {
      idx[(size_t)i * k + s] = bi[s];
      dist[(size_t)i * k + s] = bi[s] < 0 ? INFINITY : (float)sqrt((double)bd[s]);
    }
  }
}

/* L2 on float rows (general descriptors): double accumulation. */
void oracle_knn_l2_f32(const float* q, int nq, const float* t, int nt, int dim, int k, int32_t* idx, float* dist) {
  for (int i = 0; i < nq; ++i) {
    double bd[8];
    int32_t bi[8];
    for (int s = 0; s < k; ++s) {
      bd[s] = INFINITY;
      bi[s] = -1;
    }
    const float* a = q + (size_t)i * dim;
    for (int j = 0; j < nt; ++j) {
      const float* b = t + (size_t)j * dim;
      double acc = 0.0;
      for (int c = 0; c < dim; ++c) {
        const double e = (double)a[c] - (double)b[c];
        acc += e * e;
      }
      if (acc < bd[k - 1]) {
        int s = k - 1;
        while (s > 0 && acc < bd[s - 1]) {
          bd[s] = bd[s - 1];
          bi[s] = bi[s - 1];
          --s;
        }
        bd[s] = acc;
        bi[s] = j;
      }
    }
    for (int s = 0; s < k; ++s) {
      idx[(size_t)i * k + s] = bi[s];
      dist[(size_t)i * k + s] = bi[s] < 0 ? INFINITY : (float)sqrt(bd[s]);
    }
  }
}

/* Hamming on packed bytes (ORB). */
void oracle_knn_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int nbytes, int k, int32_t* idx,
                        float* dist) {
  for (int i = 0; i < nq; ++i) {
    int64_t bd[8];
    int32_t bi[8];
    for (int s = 0; s < k; ++s) {
      bd[s] = INT64_MAX;
      bi[s] = -1;
    }
    const uint8_t* a = q + (size_t)i * nbytes;
    for (int j = 0; j < nt; ++j) {
      const uint8_t* b = t + (size_t)j * nbytes;
      int32_t acc = 0;
      for (int c = 0; c < nbytes; ++c) acc += __builtin_popcount((unsigned)(a[c] ^ b[c]));
      insert_topk(k, acc, j, bd, bi);
    }
    for (int s = 0; s < k; ++s) {
      idx[(size_t)i * k + s] = bi[s];
      dist[(size_t)i * k + s] = bi[s] < 0 ? INFINITY : (float)bd[s];
    }
  }
}
