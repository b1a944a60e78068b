// ransac.cu — batched robust two-view model fitting on the GPU, standing in for
//   cv2.findEssentialMat(p1, p2, K, cv2.RANSAC, threshold=tol)   (reference scripts/lib/matcher.py:126)
//   cv2.findHomography(p1, p2, cv2.RANSAC, tol)                  (matcher.py:122, :532, :637, :803)
// as used by filter_by_transform and the bin-fitting strategies.
//
// One CTA per image pair.  Hypotheses are generated 32 at a time (one minimal
// sample per lane of warp 0: Nister's 5-point solver in fp64, or a 4-point
// DLT), every candidate model is scored against all correspondences by all
// 8 warps with warp-shuffle reductions (fp32 Sampson / transfer error on
// points held in shared memory, coalesced float4-free SoA loads), and the
// number of rounds adapts to the best inlier ratio exactly like OpenCV's
// RANSACUpdateNumIters (prob 0.999, max_iters 1000 by default).
//
// The algorithmic recipe follows the published methods (Nister 2004 for the
// 5-point problem; OpenCV documents Sampson-distance scoring against
// (threshold / mean focal)^2).  The sampler differs from OpenCV's RNG, so
// results agree as inlier SETS (tests: IoU >= 0.95), not bit for bit.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "ransac.h"

#include "layout.h"
#include "reduce.h"

namespace iam {
namespace {

constexpr int kThreads = 32;     // one warp per image pair
constexpr int kRound = 32;       // minimal samples per round
constexpr int kMaxCand = 10;     // models per sample

#ifndef IAM_HD
#define IAM_HD __host__ __device__
#endif
struct Cplx {
  double re, im;
};
IAM_HD inline Cplx cmul(Cplx a, Cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
IAM_HD inline Cplx csub(Cplx a, Cplx b) { return {a.re - b.re, a.im - b.im}; }
IAM_HD inline Cplx cdiv(Cplx a, Cplx b) {
  const double d = b.re * b.re + b.im * b.im;
  return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

// ---- tiny polynomial algebra in (x, y, z) -------------------------------------
// degree-1: [x, y, z, 1]; degree-2: [x2, y2, z2, xy, xz, yz, x, y, z, 1];
// degree-3 (Nister's elimination order):
//   [x3, y3, x2y, xy2, x2z, x2, y2z, y2, xyz, xy, xz2, xz, x, yz2, yz, y, z3, z2, z, 1]
IAM_HD inline int exp2(int m, int v) {  // exponent of variable v in degree-2 monomial m
  constexpr signed char t[10][3] = {{2, 0, 0}, {0, 2, 0}, {0, 0, 2}, {1, 1, 0}, {1, 0, 1},
                                    {0, 1, 1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};
  return t[m][v];
}
IAM_HD inline int exp1(int m, int v) { return (m < 3 && m == v) ? 1 : 0; }

IAM_HD inline int idx2(int a, int b, int c) {
  // exponents -> degree-2 monomial index
  if (a == 2) return 0;
  if (b == 2) return 1;
  if (c == 2) return 2;
  if (a == 1 && b == 1) return 3;
  if (a == 1 && c == 1) return 4;
  if (b == 1 && c == 1) return 5;
  if (a == 1) return 6;
  if (b == 1) return 7;
  if (c == 1) return 8;
  return 9;
}
IAM_HD inline int idx3(int a, int b, int c) {
  const int key = a * 16 + b * 4 + c;
  switch (key) {
    case 3 * 16: return 0;            // x3
    case 3 * 4: return 1;             // y3
    case 2 * 16 + 4: return 2;        // x2y
    case 16 + 2 * 4: return 3;        // xy2
    case 2 * 16 + 1: return 4;        // x2z
    case 2 * 16: return 5;            // x2
    case 2 * 4 + 1: return 6;         // y2z
    case 2 * 4: return 7;             // y2
    case 16 + 4 + 1: return 8;        // xyz
    case 16 + 4: return 9;            // xy
    case 16 + 2: return 10;           // xz2
    case 16 + 1: return 11;           // xz
    case 16: return 12;               // x
    case 4 + 2: return 13;            // yz2
    case 4 + 1: return 14;            // yz
    case 4: return 15;                // y
    case 3: return 16;                // z3
    case 2: return 17;                // z2
    case 1: return 18;                // z
    default: return 19;               // 1
  }
}

IAM_HD inline void p1p1(const double* a, const double* b, double* out /*10, accumulated*/, double s) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      out[idx2(exp1(i, 0) + exp1(j, 0), exp1(i, 1) + exp1(j, 1), exp1(i, 2) + exp1(j, 2))] += s * a[i] * b[j];
}
IAM_HD inline void p2p1(const double* a, const double* b, double* out /*20, accumulated*/, double s) {
  for (int i = 0; i < 10; ++i)
    for (int j = 0; j < 4; ++j)
      out[idx3(exp2(i, 0) + exp1(j, 0), exp2(i, 1) + exp1(j, 1), exp2(i, 2) + exp1(j, 2))] += s * a[i] * b[j];
}

// polynomial product in z: out[da+db+1] += a[da+1]*b[db+1]  (coefficients low -> high)
IAM_HD inline void zmul(const double* a, int da, const double* b, int db, double* out) {
  for (int i = 0; i <= da + db; ++i) out[i] = 0.0;
  for (int i = 0; i <= da; ++i)
    for (int j = 0; j <= db; ++j) out[i + j] += a[i] * b[j];
}

// ---- Nister 5-point: 5 correspondences (normalised coords) -> up to 10 E (row-major, unit Frobenius norm)
IAM_HD int five_point(const float* x1, const float* y1, const float* x2, const float* y2, const int* s,
                          float* E_out /*[10][9]*/) {
  // 1. epipolar constraint rows  x2^T E x1 = 0  ->  Q e = 0, e row-major
  double Q[5][9];
  for (int r = 0; r < 5; ++r) {
    const double a = x1[s[r]], b = y1[s[r]], c = x2[s[r]], d = y2[s[r]];
    Q[r][0] = c * a; Q[r][1] = c * b; Q[r][2] = c;
    Q[r][3] = d * a; Q[r][4] = d * b; Q[r][5] = d;
    Q[r][6] = a;     Q[r][7] = b;     Q[r][8] = 1.0;
  }
  // 2. null space by Gauss-Jordan with full pivoting
  int pivcol[5];
  bool is_piv[9] = {false, false, false, false, false, false, false, false, false};
  for (int r = 0; r < 5; ++r) {
    int br = r, bc = -1;
    double best = 0.0;
    for (int i = r; i < 5; ++i)
      for (int j = 0; j < 9; ++j)
        if (!is_piv[j] && fabs(Q[i][j]) > best) {
          best = fabs(Q[i][j]);
          br = i;
          bc = j;
        }
    if (bc < 0 || best < 1e-14) return 0;  // degenerate sample
    if (br != r)
      for (int j = 0; j < 9; ++j) {
        const double t = Q[r][j];
        Q[r][j] = Q[br][j];
        Q[br][j] = t;
      }
    pivcol[r] = bc;
    is_piv[bc] = true;
    const double inv = 1.0 / Q[r][bc];
    for (int j = 0; j < 9; ++j) Q[r][j] *= inv;
    for (int i = 0; i < 5; ++i)
      if (i != r) {
        const double f = Q[i][bc];
        if (f != 0.0)
          for (int j = 0; j < 9; ++j) Q[i][j] -= f * Q[r][j];
      }
  }
  double N[4][9];
  {
    int k = 0;
    for (int f = 0; f < 9; ++f)
      if (!is_piv[f]) {
        for (int j = 0; j < 9; ++j) N[k][j] = 0.0;
        N[k][f] = 1.0;
        for (int r = 0; r < 5; ++r) N[k][pivcol[r]] = -Q[r][f];
        ++k;
      }
    // Gram-Schmidt for conditioning
    for (int a = 0; a < 4; ++a) {
      for (int b = 0; b < a; ++b) {
        double dot = 0.0;
        for (int j = 0; j < 9; ++j) dot += N[a][j] * N[b][j];
        for (int j = 0; j < 9; ++j) N[a][j] -= dot * N[b][j];
      }
      double nn = 0.0;
      for (int j = 0; j < 9; ++j) nn += N[a][j] * N[a][j];
      nn = 1.0 / sqrt(nn);
      for (int j = 0; j < 9; ++j) N[a][j] *= nn;
    }
  }
  // E(x,y,z) = x N0 + y N1 + z N2 + N3 : entry e -> degree-1 poly [N0[e], N1[e], N2[e], N3[e]]
  double Ep[9][4];
  for (int e = 0; e < 9; ++e)
    for (int k = 0; k < 4; ++k) Ep[e][k] = N[k][e];

  // 3. ten cubic constraints -> A (10 x 20)
  double A[10][20];
  for (int r = 0; r < 10; ++r)
    for (int c = 0; c < 20; ++c) A[r][c] = 0.0;
  {
    // EEt (symmetric), degree 2
    double EEt[3][3][10];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        for (int m = 0; m < 10; ++m) EEt[i][j][m] = 0.0;
        for (int k = 0; k < 3; ++k) p1p1(Ep[i * 3 + k], Ep[j * 3 + k], EEt[i][j], 1.0);
      }
    double tr[10];
    for (int m = 0; m < 10; ++m) tr[m] = 0.5 * (EEt[0][0][m] + EEt[1][1][m] + EEt[2][2][m]);
    for (int i = 0; i < 3; ++i)
      for (int m = 0; m < 10; ++m) EEt[i][i][m] -= tr[m];       // Lambda = EEt - tr/2 I
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) p2p1(EEt[i][k], Ep[k * 3 + j], A[i * 3 + j], 1.0);   // Lambda E = 0
    // det E = 0
    double m2[10];
    for (int m = 0; m < 10; ++m) m2[m] = 0.0;
    p1p1(Ep[4], Ep[8], m2, 1.0);
    p1p1(Ep[5], Ep[7], m2, -1.0);
    p2p1(m2, Ep[0], A[9], 1.0);
    for (int m = 0; m < 10; ++m) m2[m] = 0.0;
    p1p1(Ep[3], Ep[8], m2, 1.0);
    p1p1(Ep[5], Ep[6], m2, -1.0);
    p2p1(m2, Ep[1], A[9], -1.0);
    for (int m = 0; m < 10; ++m) m2[m] = 0.0;
    p1p1(Ep[3], Ep[7], m2, 1.0);
    p1p1(Ep[4], Ep[6], m2, -1.0);
    p2p1(m2, Ep[2], A[9], 1.0);
  }
  // 4. Gauss-Jordan on the first 10 columns (partial pivoting)
  for (int c = 0; c < 10; ++c) {
    int br = c;
    double best = fabs(A[c][c]);
    for (int i = c + 1; i < 10; ++i)
      if (fabs(A[i][c]) > best) {
        best = fabs(A[i][c]);
        br = i;
      }
    if (best < 1e-14) return 0;
    if (br != c)
      for (int j = 0; j < 20; ++j) {
        const double t = A[c][j];
        A[c][j] = A[br][j];
        A[br][j] = t;
      }
    const double inv = 1.0 / A[c][c];
    for (int j = c; j < 20; ++j) A[c][j] *= inv;
    for (int i = 0; i < 10; ++i)
      if (i != c) {
        const double f = A[i][c];
        if (f != 0.0)
          for (int j = c; j < 20; ++j) A[i][j] -= f * A[c][j];
      }
  }
  // 5. B(z): rows <k> = <e> - z<f>, <l> = <g> - z<h>, <m> = <i> - z<j>
  double B[3][3][5];  // [row][x|y|1][coeff of z^p]
  const int hi[3] = {4, 6, 8}, lo[3] = {5, 7, 9};
  for (int r = 0; r < 3; ++r) {
    const double* e = A[hi[r]];
    const double* f = A[lo[r]];
    for (int part = 0; part < 2; ++part) {
      const int c0 = 10 + part * 3;  // [.z2, .z, .]
      B[r][part][0] = e[c0 + 2];
      B[r][part][1] = e[c0 + 1] - f[c0 + 2];
      B[r][part][2] = e[c0] - f[c0 + 1];
      B[r][part][3] = -f[c0];
      B[r][part][4] = 0.0;
    }
    B[r][2][0] = e[19];
    B[r][2][1] = e[18] - f[19];
    B[r][2][2] = e[17] - f[18];
    B[r][2][3] = e[16] - f[17];
    B[r][2][4] = -f[16];
  }
  // det B(z): degree 10
  double poly[11];
  for (int i = 0; i < 11; ++i) poly[i] = 0.0;
  {
    double t1[8], t2[8], t3[11];
    // k1 (l2 m3 - l3 m2)
    zmul(B[1][1], 3, B[2][2], 4, t1);
    zmul(B[1][2], 4, B[2][1], 3, t2);
    for (int i = 0; i < 8; ++i) t1[i] -= t2[i];
    zmul(B[0][0], 3, t1, 7, t3);
    for (int i = 0; i < 11; ++i) poly[i] += t3[i];
    // - k2 (l1 m3 - l3 m1)
    zmul(B[1][0], 3, B[2][2], 4, t1);
    zmul(B[1][2], 4, B[2][0], 3, t2);
    for (int i = 0; i < 8; ++i) t1[i] -= t2[i];
    zmul(B[0][1], 3, t1, 7, t3);
    for (int i = 0; i < 11; ++i) poly[i] -= t3[i];
    // + k3 (l1 m2 - l2 m1)
    double u1[7], u2[7];
    zmul(B[1][0], 3, B[2][1], 3, u1);
    zmul(B[1][1], 3, B[2][0], 3, u2);
    for (int i = 0; i < 7; ++i) u1[i] -= u2[i];
    zmul(B[0][2], 4, u1, 6, t3);
    for (int i = 0; i < 11; ++i) poly[i] += t3[i];
  }
  // 6. REAL roots of the degree-10 polynomial by Sturm sequences (the method of Nister's paper): the k-th smallest
  //    real root is bracketed by bisection on the root count N(x) = V(-R) - V(x) until it is alone in its bracket,
  //    then plain bisection on the sign of the polynomial narrows it; Newton steps below polish it.  (The previous
  //    Durand-Kerner iteration on all ten complex roots took 80 % of the kernel's time: a warp runs until its slowest
  //    lane has converged.)
  double mx = 0.0;
  for (int i = 0; i < 11; ++i) mx = fmax(mx, fabs(poly[i]));
  if (!(mx > 0.0) || !isfinite(mx)) return 0;
  int deg = 10;
  while (deg > 0 && fabs(poly[deg]) < 1e-13 * mx) --deg;
  if (deg < 1) return 0;
  double st[11][11];   // Sturm chain, st[k] has degree sd[k] (coefficients low -> high)
  int sd[11];
  int ns = 2;
  for (int i = 0; i <= deg; ++i) st[0][i] = poly[i] / poly[deg];
  sd[0] = deg;
  for (int i = 0; i < deg; ++i) st[1][i] = st[0][i + 1] * (i + 1);
  sd[1] = deg - 1;
  double rad = 0.0;
  for (int i = 0; i < deg; ++i) rad = fmax(rad, fabs(st[0][i]));
  rad = fmin(1.0 + rad, 1e6);
  while (sd[ns - 1] > 0 && ns < 11) {   // st[ns] = -rem(st[ns-2], st[ns-1]), scaled to unit size
    double r[11];
    const int da = sd[ns - 2], db = sd[ns - 1];
    for (int i = 0; i <= da; ++i) r[i] = st[ns - 2][i];
    const double lead = st[ns - 1][db];
    for (int k = da; k >= db; --k) {
      const double q = r[k] / lead;
      for (int j = 0; j <= db; ++j) r[k - db + j] -= q * st[ns - 1][j];
    }
    int dr = db - 1;
    double big = 0.0;
    for (int i = 0; i <= dr; ++i) big = fmax(big, fabs(r[i]));
    while (dr > 0 && fabs(r[dr]) <= 1e-14 * big) --dr;
    if (!(big > 0.0)) break;            // exact division: repeated roots, the chain ends here
    const double sc = -1.0 / big;
    for (int i = 0; i <= dr; ++i) st[ns][i] = r[i] * sc;
    sd[ns] = dr;
    ++ns;
  }
  auto sign_changes = [&](double x) {
    int v = 0, last = 0;
    for (int k = 0; k < ns; ++k) {
      double p = st[k][sd[k]];
      for (int i = sd[k] - 1; i >= 0; --i) p = p * x + st[k][i];
      const int sg = p > 0.0 ? 1 : (p < 0.0 ? -1 : 0);
      if (sg != 0) {
        if (last != 0 && sg != last) ++v;
        last = sg;
      }
    }
    return v;
  };
  auto pval = [&](double x) {
    double p = st[0][deg];
    for (int i = deg - 1; i >= 0; --i) p = p * x + st[0][i];
    return p;
  };
  const int v_lo = sign_changes(-rad);
  const int n_real = v_lo - sign_changes(rad);
  double rroot[10];
  int n_roots = 0;
  for (int k = 1; k <= n_real && n_roots < 10; ++k) {
    // smallest x with N(x) >= k
    double lo = -rad, hi = rad;
    int n_lo = 0, n_hi = n_real;        // N(lo) < k <= N(hi)
    int it = 0;
    for (; it < 64 && !(n_hi - n_lo == 1 && it >= 6); ++it) {
      const double mid = 0.5 * (lo + hi);
      const int nm = v_lo - sign_changes(mid);
      if (nm >= k) {
        hi = mid;
        n_hi = nm;
      } else {
        lo = mid;
        n_lo = nm;
      }
      if (hi - lo <= 1e-15 * (1.0 + fabs(lo))) break;
    }
    if (n_hi - n_lo == 1) {             // alone in (lo, hi]: bisection on the sign of the polynomial
      const double plo = pval(lo);
      for (int b = 0; b < 40; ++b) {
        const double mid = 0.5 * (lo + hi);
        const double pm = pval(mid);
        if ((pm > 0.0) == (plo > 0.0) && pm != 0.0) lo = mid; else hi = mid;
      }
    } else {
      k += n_hi - n_lo - 1;             // a cluster narrower than the resolution: one representative
    }
    rroot[n_roots++] = 0.5 * (lo + hi);
  }
  const double* polyn = st[0];
  // 7. back-substitute every real root
  int n_out = 0;
  for (int i = 0; i < n_roots && n_out < kMaxCand; ++i) {
    double z = rroot[i];
    for (int nw = 0; nw < 3; ++nw) {  // Newton polish on the real polynomial
      double p = polyn[deg], dp = 0.0;
      for (int k = deg - 1; k >= 0; --k) {
        dp = dp * z + p;
        p = p * z + polyn[k];
      }
      if (fabs(dp) > 0.0) z -= p / dp;
    }
    double Bz[3][3];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double v = B[r][c][4];
        for (int k = 3; k >= 0; --k) v = v * z + B[r][c][k];
        Bz[r][c] = v;
      }
    // null vector (x, y, 1) of Bz from the best-conditioned pair of rows
    double bx = 0, by = 0, bw = 0;
    for (int r0 = 0; r0 < 3; ++r0) {
      const int r1 = (r0 + 1) % 3;
      const double vx = Bz[r0][1] * Bz[r1][2] - Bz[r0][2] * Bz[r1][1];
      const double vy = Bz[r0][2] * Bz[r1][0] - Bz[r0][0] * Bz[r1][2];
      const double vw = Bz[r0][0] * Bz[r1][1] - Bz[r0][1] * Bz[r1][0];
      if (fabs(vw) > fabs(bw)) {
        bx = vx;
        by = vy;
        bw = vw;
      }
    }
    if (fabs(bw) < 1e-300) continue;
    const double x = bx / bw, y = by / bw;
    double E[9], nn = 0.0;
    for (int e = 0; e < 9; ++e) {
      E[e] = x * N[0][e] + y * N[1][e] + z * N[2][e] + N[3][e];
      nn += E[e] * E[e];
    }
    if (!(nn > 0.0) || !isfinite(nn)) continue;
    nn = 1.0 / sqrt(nn);
    for (int e = 0; e < 9; ++e) E_out[n_out * 9 + e] = static_cast<float>(E[e] * nn);
    ++n_out;
  }
  return n_out;
}

// ---- 2-point similarity (cv2.estimateAffinePartial2D's model: rotation, uniform scale, translation) ----
// [a -b tx; b a ty; 0 0 1] through two correspondences: (a + ib) = (q1 - q0) / (p1 - p0) as complex numbers.
IAM_HD int two_point(const float* x1, const float* y1, const float* x2, const float* y2, const int* s, float* M_out) {
  const double px = double(x1[s[1]]) - x1[s[0]], py = double(y1[s[1]]) - y1[s[0]];
  const double qx = double(x2[s[1]]) - x2[s[0]], qy = double(y2[s[1]]) - y2[s[0]];
  const double den = px * px + py * py;
  if (!(den > 1e-24)) return 0;
  const double a = (qx * px + qy * py) / den, b = (qy * px - qx * py) / den;
  const double tx = x2[s[0]] - (a * x1[s[0]] - b * y1[s[0]]), ty = y2[s[0]] - (b * x1[s[0]] + a * y1[s[0]]);
  M_out[0] = float(a); M_out[1] = float(-b); M_out[2] = float(tx);
  M_out[3] = float(b); M_out[4] = float(a); M_out[5] = float(ty);
  M_out[6] = 0.f; M_out[7] = 0.f; M_out[8] = 1.f;
  return 1;
}

// ---- 4-point homography (normalised coordinates), h33 fixed by unit norm ----
IAM_HD int four_point(const float* x1, const float* y1, const float* x2, const float* y2, const int* s,
                          float* H_out) {
  double M[8][9];
  for (int r = 0; r < 4; ++r) {
    const double x = x1[s[r]], y = y1[s[r]], u = x2[s[r]], v = y2[s[r]];
    double* a = M[2 * r];
    double* b = M[2 * r + 1];
    a[0] = x; a[1] = y; a[2] = 1; a[3] = 0; a[4] = 0; a[5] = 0; a[6] = -u * x; a[7] = -u * y; a[8] = u;
    b[0] = 0; b[1] = 0; b[2] = 0; b[3] = x; b[4] = y; b[5] = 1; b[6] = -v * x; b[7] = -v * y; b[8] = v;
  }
  for (int c = 0; c < 8; ++c) {
    int br = c;
    double best = fabs(M[c][c]);
    for (int i = c + 1; i < 8; ++i)
      if (fabs(M[i][c]) > best) {
        best = fabs(M[i][c]);
        br = i;
      }
    if (best < 1e-12) return 0;
    if (br != c)
      for (int j = 0; j < 9; ++j) {
        const double t = M[c][j];
        M[c][j] = M[br][j];
        M[br][j] = t;
      }
    const double inv = 1.0 / M[c][c];
    for (int j = c; j < 9; ++j) M[c][j] *= inv;
    for (int i = 0; i < 8; ++i)
      if (i != c) {
        const double f = M[i][c];
        if (f != 0.0)
          for (int j = c; j < 9; ++j) M[i][j] -= f * M[c][j];
      }
  }
  for (int e = 0; e < 8; ++e) H_out[e] = static_cast<float>(M[e][8]);
  H_out[8] = 1.0f;
  return 1;
}

// ---- 7-point fundamental matrix (normalised coordinates): up to 3 rank-2 F with x2^T F x1 = 0 ----
// The published minimal method behind cv2.findFundamentalMat(..., RANSAC): the 7 x 9 epipolar system has a
// two-dimensional null space {F1, F2}; det(l F1 + (1 - l) F2) = 0 is a cubic in l (Hartley & Zisserman 11.1.2).
IAM_HD inline double det3(const double* m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
IAM_HD int seven_point(const float* x1, const float* y1, const float* x2, const float* y2, const int* s, float* F_out) {
  double M[7][9];
  for (int r = 0; r < 7; ++r) {
    const double x = x1[s[r]], y = y1[s[r]], u = x2[s[r]], v = y2[s[r]];
    double* a = M[r];
    a[0] = u * x; a[1] = u * y; a[2] = u; a[3] = v * x; a[4] = v * y; a[5] = v; a[6] = x; a[7] = y; a[8] = 1.0;
  }
  // Gauss-Jordan with full pivoting: 7 pivot columns, 2 free ones
  int pcol[7];
  bool used[9] = {false, false, false, false, false, false, false, false, false};
  for (int r = 0; r < 7; ++r) {
    int br = r, bc = -1;
    double best = 0.0;
    for (int i = r; i < 7; ++i)
      for (int c = 0; c < 9; ++c)
        if (!used[c] && fabs(M[i][c]) > best) {
          best = fabs(M[i][c]);
          br = i;
          bc = c;
        }
    if (bc < 0 || best < 1e-12) return 0;
    if (br != r)
      for (int c = 0; c < 9; ++c) {
        const double t = M[r][c];
        M[r][c] = M[br][c];
        M[br][c] = t;
      }
    used[bc] = true;
    pcol[r] = bc;
    const double inv = 1.0 / M[r][bc];
    for (int c = 0; c < 9; ++c) M[r][c] *= inv;
    for (int i = 0; i < 7; ++i)
      if (i != r) {
        const double f = M[i][bc];
        if (f != 0.0)
          for (int c = 0; c < 9; ++c) M[i][c] -= f * M[r][c];
      }
  }
  double N[2][9];
  int nf = 0;
  for (int c = 0; c < 9 && nf < 2; ++c)
    if (!used[c]) {
      for (int e = 0; e < 9; ++e) N[nf][e] = 0.0;
      N[nf][c] = 1.0;
      for (int r = 0; r < 7; ++r) N[nf][pcol[r]] = -M[r][c];
      ++nf;
    }
  if (nf != 2) return 0;
  // det(N1 + l * D), D = N0 - N1:  c0 + c1 l + c2 l^2 + c3 l^3 from the determinant at l = 0, +1, -1 and det(D)
  double D[9], P[9], Q[9];
  for (int e = 0; e < 9; ++e) {
    D[e] = N[0][e] - N[1][e];
    P[e] = N[1][e] + D[e];
    Q[e] = N[1][e] - D[e];
  }
  const double c0 = det3(N[1]), c3 = det3(D), dp = det3(P), dm = det3(Q);
  const double c2 = 0.5 * (dp + dm) - c0, c1 = 0.5 * (dp - dm) - c3;
  double roots[3];
  int nr = 0;
  const double scale = fabs(c0) + fabs(c1) + fabs(c2) + fabs(c3);
  if (!(scale > 0.0)) return 0;
  if (fabs(c3) < 1e-12 * scale) {  // quadratic (or lower)
    if (fabs(c2) < 1e-12 * scale) {
      if (fabs(c1) > 0.0) roots[nr++] = -c0 / c1;
    } else {
      const double disc = c1 * c1 - 4.0 * c2 * c0;
      if (disc >= 0.0) {
        const double q = -0.5 * (c1 + (c1 >= 0.0 ? sqrt(disc) : -sqrt(disc)));
        roots[nr++] = q / c2;
        if (q != 0.0) roots[nr++] = c0 / q;
      }
    }
  } else {  // x^3 + a x^2 + b x + c = 0 (Numerical Recipes 5.6)
    const double a = c2 / c3, b = c1 / c3, c = c0 / c3;
    const double Qq = (a * a - 3.0 * b) / 9.0, R = (2.0 * a * a * a - 9.0 * a * b + 27.0 * c) / 54.0;
    if (R * R < Qq * Qq * Qq) {
      const double th = acos(fmax(-1.0, fmin(1.0, R / sqrt(Qq * Qq * Qq)))), sq = -2.0 * sqrt(Qq);
      roots[nr++] = sq * cos(th / 3.0) - a / 3.0;
      roots[nr++] = sq * cos((th + 6.283185307179586) / 3.0) - a / 3.0;
      roots[nr++] = sq * cos((th - 6.283185307179586) / 3.0) - a / 3.0;
    } else {
      const double A = -(R >= 0.0 ? 1.0 : -1.0) * cbrt(fabs(R) + sqrt(R * R - Qq * Qq * Qq));
      const double B = A != 0.0 ? Qq / A : 0.0;
      roots[nr++] = (A + B) - a / 3.0;
    }
  }
  int n_out = 0;
  for (int i = 0; i < nr; ++i) {
    double l = roots[i];
    for (int nw = 0; nw < 2; ++nw) {  // Newton polish
      const double f = ((c3 * l + c2) * l + c1) * l + c0, df = (3.0 * c3 * l + 2.0 * c2) * l + c1;
      if (fabs(df) > 0.0) l -= f / df;
    }
    double F[9], nn = 0.0;
    for (int e = 0; e < 9; ++e) {
      F[e] = N[1][e] + l * D[e];
      nn += F[e] * F[e];
    }
    if (!(nn > 0.0) || !isfinite(nn)) continue;
    nn = 1.0 / sqrt(nn);
    for (int e = 0; e < 9; ++e) F_out[n_out * 9 + e] = static_cast<float>(F[e] * nn);
    ++n_out;
  }
  return n_out;
}

struct PairXform {  // per-pair point normalisation
  float a1x, a1y, s1x, s1y;  // image 1: xn = (x - a1x) * s1x
  float a2x, a2y, s2x, s2y;
  float thr2;                // squared threshold in normalised units (errors measured in image 2)
  float thr_ratio;           // fundamental: thr2 / (squared threshold of the error measured in image 1)
};

__device__ __forceinline__ float model_error(int model, const float* M, float x1, float y1, float x2, float y2,
                                             float thr_ratio) {
  if (model == 0) {  // Sampson distance of x2^T E x1
    const float ex = M[0] * x1 + M[1] * y1 + M[2];
    const float ey = M[3] * x1 + M[4] * y1 + M[5];
    const float ez = M[6] * x1 + M[7] * y1 + M[8];
    const float tx = M[0] * x2 + M[3] * y2 + M[6];
    const float ty = M[1] * x2 + M[4] * y2 + M[7];
    const float r = x2 * ex + y2 * ey + ez;
    return r * r / (ex * ex + ey * ey + tx * tx + ty * ty);
  }
  if (model == 2) {
    // cv2.findFundamentalMat's RANSAC error: the larger of the two squared point-to-epipolar-line distances
    // (line F x1 in image 2, line F^T x2 in image 1).  The two images carry their own normalisation scales, so the
    // distance in image 1 is rescaled to image-2 units (thr_ratio) and one comparison serves both.
    const float a = M[0] * x1 + M[1] * y1 + M[2];
    const float b = M[3] * x1 + M[4] * y1 + M[5];
    const float c = M[6] * x1 + M[7] * y1 + M[8];
    const float ta = M[0] * x2 + M[3] * y2 + M[6];
    const float tb = M[1] * x2 + M[4] * y2 + M[7];
    const float r = x2 * a + y2 * b + c;
    const float r2 = r * r;
    return fmaxf(r2 / (a * a + b * b), thr_ratio * r2 / (ta * ta + tb * tb));
  }
  const float w = M[6] * x1 + M[7] * y1 + M[8];
  const float iw = 1.0f / w;
  const float dx = (M[0] * x1 + M[1] * y1 + M[2]) * iw - x2;
  const float dy = (M[3] * x1 + M[4] * y1 + M[5]) * iw - y2;
  return dx * dx + dy * dy;
}

__device__ __forceinline__ double warp_sum(double v) {
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

struct RansacArgs {
  int model;
  int max_iters;
  uint32_t seed;
  float fx, fy, cx, cy;   // essential: K
  float thr_px;
  double prob;
  // direct mode: packed point lists
  const float* pts1;
  const float* pts2;
  const int* off;
  // gather mode (pts1 == nullptr): the device match tables of the last match call + the images' key points
  const int* table;       // [P][cap][2]
  int* count;             // [P]
  int cap;
  int compact;            // gather mode: drop the outliers from the tables in place (filter_by_transform, matcher.py:134-141)
  int min_pairs;          // gather mode: pairs with fewer rows are emptied without a fit (matcher.py:99-101)
  const RedJob* jobs;     // forward job of pair p = jobs[2p]: q_slot / t_slot
  const ImgDev* imgs;
  // outputs
  unsigned char* out_mask;  // direct: [off[P]]; gather: [P][cap] (may be null)
  float* out_model;         // [P][9]: E for normalised image coordinates, H / F for pixels
  int* out_inliers;         // [P]
};

// One WARP per image pair: its 32 lanes solve 32 minimal samples (the solvers are long fp64 routines: with one
// lane per sample no lane idles, where a wider CTA would leave every warp but one waiting), then the same warp scores
// every candidate model, lanes striding the points.  Several pairs share an SM (registers: ~250 per thread).
#ifndef IAM_RANSAC_MIN_BLOCKS
#define IAM_RANSAC_MIN_BLOCKS 1   // A/B aid: e.g. 16 caps the kernel at 128 registers (more pairs per SM, more spills)
#endif
__global__ void __launch_bounds__(kThreads, IAM_RANSAC_MIN_BLOCKS)
ransac_kernel(const RansacArgs A) {
  extern __shared__ float s_pts[];
  __shared__ float s_cand[kRound * kMaxCand * 9];
  __shared__ int s_ncand[kRound];
  __shared__ int s_score[kRound * kMaxCand];

  const int p = blockIdx.x;
  const int lane = threadIdx.x;
  const int model = A.model;
  const bool gather = A.pts1 == nullptr;
  const int o = gather ? 0 : A.off[p];
  const int n = gather ? A.count[p] : A.off[p + 1] - o;
  const int msize = (model == 0) ? 5 : (model == 2) ? 7 : (model == 3) ? 2 : 4;
  if (gather && n < A.min_pairs) {  // matcher.py:99-101: too few matches, the list is cleared
    if (lane == 0) {
      if (A.compact) A.count[p] = 0;
      A.out_inliers[p] = 0;
      for (int e = 0; e < 9; ++e) A.out_model[p * 9 + e] = 0.f;
    }
    if (A.out_mask)
      for (int i = lane; i < n; i += 32) A.out_mask[p * A.cap + i] = 0;
    return;
  }
  float* sx1 = s_pts;
  float* sy1 = sx1 + n;
  float* sx2 = sy1 + n;
  float* sy2 = sx2 + n;
  const int2* tab = gather ? reinterpret_cast<const int2*>(A.table) + static_cast<size_t>(p) * A.cap : nullptr;
  if (gather) {
    const RedJob jb = A.jobs[2 * p];
    const float2* k1 = A.imgs[jb.q_slot].kp_xy;
    const float2* k2 = A.imgs[jb.t_slot].kp_xy;
    for (int i = lane; i < n; i += 32) {
      const int2 qt = tab[i];
      const float2 a = k1[qt.x], b = k2[qt.y];
      sx1[i] = a.x; sy1[i] = a.y; sx2[i] = b.x; sy2[i] = b.y;
    }
  } else {
    for (int i = lane; i < n; i += 32) {
      sx1[i] = A.pts1[2 * (o + i)]; sy1[i] = A.pts1[2 * (o + i) + 1];
      sx2[i] = A.pts2[2 * (o + i)]; sy2[i] = A.pts2[2 * (o + i) + 1];
    }
  }
  __syncwarp();
  // normalisation
  PairXform X;
  if (model == 0) {
    // K^-1 with fx, fy, cx, cy; threshold / mean focal (how OpenCV's findEssentialMat treats pixel thresholds)
    X.a1x = X.a2x = A.cx;
    X.a1y = X.a2y = A.cy;
    X.s1x = X.s2x = 1.0f / A.fx;
    X.s1y = X.s2y = 1.0f / A.fy;
    const double t = double(A.thr_px) / ((double(A.fx) + double(A.fy)) / 2.0);
    X.thr2 = float(t * t);
    X.thr_ratio = 1.0f;
  } else {
    // Hartley normalisation; the transfer error lives in image 2, so its isotropic scale carries the threshold
    double m1x = 0, m1y = 0, m2x = 0, m2y = 0;
    for (int i = lane; i < n; i += 32) {
      m1x += sx1[i]; m1y += sy1[i]; m2x += sx2[i]; m2y += sy2[i];
    }
    const double inv = n > 0 ? 1.0 / n : 0.0;
    m1x = warp_sum(m1x) * inv; m1y = warp_sum(m1y) * inv; m2x = warp_sum(m2x) * inv; m2y = warp_sum(m2y) * inv;
    double d1 = 0, d2 = 0;
    for (int i = lane; i < n; i += 32) {
      d1 += hypot(double(sx1[i]) - m1x, double(sy1[i]) - m1y);
      d2 += hypot(double(sx2[i]) - m2x, double(sy2[i]) - m2y);
    }
    d1 = warp_sum(d1);
    d2 = warp_sum(d2);
    const double s1 = d1 > 0 ? sqrt(2.0) * n / d1 : 1.0;
    const double s2 = d2 > 0 ? sqrt(2.0) * n / d2 : 1.0;
    X.a1x = float(m1x); X.a1y = float(m1y); X.s1x = X.s1y = float(s1);
    X.a2x = float(m2x); X.a2y = float(m2y); X.s2x = X.s2y = float(s2);
    const double t = double(A.thr_px) * double(X.s2x), tb = double(A.thr_px) * double(X.s1x);
    X.thr2 = float(t * t);
    X.thr_ratio = tb > 0 ? float((t * t) / (tb * tb)) : 1.0f;
  }
  for (int i = lane; i < n; i += 32) {
    sx1[i] = (sx1[i] - X.a1x) * X.s1x;
    sy1[i] = (sy1[i] - X.a1y) * X.s1y;
    sx2[i] = (sx2[i] - X.a2x) * X.s2x;
    sy2[i] = (sy2[i] - X.a2y) * X.s2y;
  }
  __syncwarp();

  float best[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // warp-uniform
  int best_count = 0, niters = A.max_iters;
  if (n >= msize) {
    for (int base = 0; base < A.max_iters && base < niters; base += kRound) {
      int nc = 0;
      const int it = base + lane;
      if (it < niters) {
        int s[7];
        uint32_t h = hash32(A.seed ^ hash32(uint32_t(p) * 0x9e3779b9u + uint32_t(it)));
        for (int k = 0; k < msize; ++k) {
          for (int tries = 0; tries < 64; ++tries) {
            h = hash32(h + 0x6d2b79f5u);
            const int cand = int(h % uint32_t(n));
            bool dup = false;
            for (int q = 0; q < k; ++q) dup = dup || (s[q] == cand);
            if (!dup) {
              s[k] = cand;
              break;
            }
            if (tries == 63) s[k] = (s[k > 0 ? k - 1 : 0] + 1 + k) % n;
          }
        }
        nc = (model == 0)   ? five_point(sx1, sy1, sx2, sy2, s, &s_cand[lane * kMaxCand * 9])
             : (model == 2) ? seven_point(sx1, sy1, sx2, sy2, s, &s_cand[lane * kMaxCand * 9])
             : (model == 3) ? two_point(sx1, sy1, sx2, sy2, s, &s_cand[lane * kMaxCand * 9])
                            : four_point(sx1, sy1, sx2, sy2, s, &s_cand[lane * kMaxCand * 9]);
      }
      s_ncand[lane] = nc;
      __syncwarp();
      // score every candidate of the round: lanes stride the points, warp-shuffle reduction of the counts
      for (int c = 0; c < kRound * kMaxCand; ++c) {
        const int smp = c / kMaxCand, r = c % kMaxCand;
        if (r >= s_ncand[smp]) {
          c += kMaxCand - 1 - r;   // the rest of this sample's slots are empty too
          continue;
        }
        float M[9];
        for (int e = 0; e < 9; ++e) M[e] = s_cand[c * 9 + e];
        int cnt = 0;
        for (int i = lane; i < n; i += 32) cnt += model_error(model, M, sx1[i], sy1[i], sx2[i], sy2[i], X.thr_ratio) <= X.thr2;
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        if (lane == 0) s_score[c] = cnt;
      }
      __syncwarp();
      // every lane walks the scores identically: the best model and the iteration bound stay warp-uniform
      for (int smp = 0; smp < kRound && base + smp < niters; ++smp) {
        for (int r = 0; r < s_ncand[smp]; ++r) {
          const int cnt = s_score[smp * kMaxCand + r];
          if (cnt > max(best_count, msize - 1)) {
            best_count = cnt;
            for (int e = 0; e < 9; ++e) best[e] = s_cand[(smp * kMaxCand + r) * 9 + e];
            // adaptive iteration count (the standard RANSAC bound, OpenCV's RANSACUpdateNumIters)
            const double ep = fmin(fmax(1.0 - double(cnt) / double(n), 0.0), 1.0);
            const double num = log(fmax(1.0 - A.prob, 1e-300));
            const double den = 1.0 - pow(1.0 - ep, double(msize));
            int ni = A.max_iters;
            if (den < 1e-300)
              ni = 0;
            else {
              const double lden = log(den);
              if (!(lden >= 0.0) && -num < double(A.max_iters) * (-lden)) ni = int(rint(num / lden));
            }
            niters = min(niters, ni);
          }
        }
      }
      __syncwarp();
    }
  }
  // final mask with the best model
  const bool have = best_count > 0;
  const int mask_base = gather ? p * A.cap : o;
  int kept = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    const bool in = i < n && have && model_error(model, best, sx1[i], sy1[i], sx2[i], sy2[i], X.thr_ratio) <= X.thr2;
    if (A.out_mask && i < n) A.out_mask[mask_base + i] = in ? 1 : 0;
    if (gather && A.compact) {  // order-preserving compaction of the table rows (matcher.py:134-141)
      const unsigned bal = __ballot_sync(0xffffffffu, in);
      int2 row = make_int2(0, 0);
      if (in) row = tab[i];
      __syncwarp();
      if (in) const_cast<int2*>(tab)[kept + __popc(bal & ((1u << lane) - 1u))] = row;
      kept += __popc(bal);
      __syncwarp();
    }
  }
  if (gather && A.compact && lane == 0) A.count[p] = kept;
  if (model == 3 && have) {
    // cv2.estimateAffinePartial2D polishes the RANSAC model on its inliers (Levenberg-Marquardt, refineIters = 10); the
    // model is linear in (a, b, tx, ty), so the optimum it converges to is the closed-form least-squares fit.  The
    // mask stays the RANSAC model's, as in OpenCV.
    double sp_x = 0, sp_y = 0, sq_x = 0, sq_y = 0, cnt = 0;
    for (int i = lane; i < n; i += 32)
      if (model_error(model, best, sx1[i], sy1[i], sx2[i], sy2[i], X.thr_ratio) <= X.thr2) {
        sp_x += sx1[i]; sp_y += sy1[i]; sq_x += sx2[i]; sq_y += sy2[i]; cnt += 1;
      }
    sp_x = warp_sum(sp_x); sp_y = warp_sum(sp_y); sq_x = warp_sum(sq_x); sq_y = warp_sum(sq_y); cnt = warp_sum(cnt);
    if (cnt >= 2) {
      const double mpx = sp_x / cnt, mpy = sp_y / cnt, mqx = sq_x / cnt, mqy = sq_y / cnt;
      double num_a = 0, num_b = 0, den = 0;
      for (int i = lane; i < n; i += 32)
        if (model_error(model, best, sx1[i], sy1[i], sx2[i], sy2[i], X.thr_ratio) <= X.thr2) {
          const double px = sx1[i] - mpx, py = sy1[i] - mpy, qx = sx2[i] - mqx, qy = sy2[i] - mqy;
          num_a += px * qx + py * qy;
          num_b += px * qy - py * qx;
          den += px * px + py * py;
        }
      num_a = warp_sum(num_a); num_b = warp_sum(num_b); den = warp_sum(den);
      if (den > 1e-24) {
        const double a = num_a / den, b = num_b / den;
        best[0] = float(a); best[1] = float(-b); best[2] = float(mqx - (a * mpx - b * mpy));
        best[3] = float(b); best[4] = float(a); best[5] = float(mqy - (b * mpx + a * mpy));
      }
    }
  }
  if (lane == 0) {
    // back to caller units: E is reported for normalised image coordinates like cv2.findEssentialMat does
    // (x2n^T E x1n = 0); H and F are mapped back to pixels and scaled to a unit last element.
    double o9[9];
    if (!have) {
      for (int e = 0; e < 9; ++e) o9[e] = 0.0;
    } else if (model == 0) {
      for (int e = 0; e < 9; ++e) o9[e] = best[e];
    } else {
      const double T1[9] = {X.s1x, 0, -double(X.s1x) * X.a1x, 0, X.s1y, -double(X.s1y) * X.a1y, 0, 0, 1};
      double L[9];  // homography: T2^-1; fundamental: T2^T
      if (model == 1 || model == 3) {
        const double l[9] = {1.0 / X.s2x, 0, X.a2x, 0, 1.0 / X.s2y, X.a2y, 0, 0, 1};
        for (int e = 0; e < 9; ++e) L[e] = l[e];
      } else {
        const double l[9] = {X.s2x, 0, 0, 0, X.s2y, 0, -double(X.s2x) * X.a2x, -double(X.s2y) * X.a2y, 1};
        for (int e = 0; e < 9; ++e) L[e] = l[e];
      }
      double t[9];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          t[i * 3 + j] = 0;
          for (int k = 0; k < 3; ++k) t[i * 3 + j] += double(best[i * 3 + k]) * T1[k * 3 + j];
        }
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          o9[i * 3 + j] = 0;
          for (int k = 0; k < 3; ++k) o9[i * 3 + j] += L[i * 3 + k] * t[k * 3 + j];
        }
      const double sc = fabs(o9[8]) > 1e-300 ? 1.0 / o9[8] : 1.0;
      for (int e = 0; e < 9; ++e) o9[e] *= sc;
    }
    for (int e = 0; e < 9; ++e) A.out_model[p * 9 + e] = static_cast<float>(o9[e]);
    A.out_inliers[p] = have ? best_count : 0;
  }
}

cudaError_t launch(const RansacArgs& a, int n_pairs, int max_n, cudaStream_t stream) {
  const size_t smem = size_t(std::max(max_n, 1)) * 4 * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  ransac_kernel<<<n_pairs, kThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

void fill_common(RansacArgs& a, int model, const double* K, double threshold_px, double prob, int max_iters, uint32_t seed) {
  a.model = model;
  a.max_iters = max_iters <= 0 ? 1000 : max_iters;
  a.prob = (prob > 0.0 && prob < 1.0) ? prob : 0.999;
  a.seed = seed;
  a.thr_px = float(threshold_px);
  a.fx = K ? float(K[0]) : 1.f;
  a.fy = K ? float(K[4]) : 1.f;
  a.cx = K ? float(K[2]) : 0.f;
  a.cy = K ? float(K[5]) : 0.f;
}

}  // namespace

int debug_minimal_solver(int model, const float* x1, const float* y1, const float* x2, const float* y2, float* out) {
  const int s[7] = {0, 1, 2, 3, 4, 5, 6};
  return model == 0 ? five_point(x1, y1, x2, y2, s, out) : model == 2 ? seven_point(x1, y1, x2, y2, s, out)
         : model == 3 ? two_point(x1, y1, x2, y2, s, out) : four_point(x1, y1, x2, y2, s, out);
}

RansacScratch::~RansacScratch() {
  if (buf) cudaFree(buf);
}

int ransac_pairs(int model, const float* pts1, const float* pts2, const int32_t* off, int n_pairs, const double* K,
                 double threshold_px, double prob, int max_iters, uint32_t seed, uint8_t* out_mask, double* out_model,
                 int32_t* out_inliers, RansacScratch* scratch, cudaStream_t stream, std::string* err) {
  auto bad = [&](int code, const char* what, cudaError_t e) {
    if (err) *err = std::string(what) + ": " + cudaGetErrorString(e);
    return code;
  };
  if (n_pairs == 0) return 0;
  const int total = off[n_pairs];
  int max_n = 0;
  for (int p = 0; p < n_pairs; ++p) {
    if (off[p + 1] < off[p]) {
      if (err) *err = "offsets must be non-decreasing";
      return -1;
    }
    max_n = std::max(max_n, off[p + 1] - off[p]);
  }
  if (max_n > 12000) {
    if (err) *err = "more than 12000 correspondences in one pair";
    return -5;
  }
  // one device block, kept between calls: [pts1][pts2][off][model][inliers][mask]
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t pb = up(size_t(std::max(total, 1)) * 2 * sizeof(float));
  const size_t o_p2 = pb, o_off = 2 * pb, o_model = o_off + up(size_t(n_pairs + 1) * sizeof(int));
  const size_t o_inl = o_model + up(size_t(n_pairs) * 9 * sizeof(float)), o_mask = o_inl + up(size_t(n_pairs) * sizeof(int));
  const size_t bytes = o_mask + up(size_t(std::max(total, 1)));
  cudaError_t e;
  RansacScratch local;
  RansacScratch* sc = scratch ? scratch : &local;
  if (sc->cap < bytes) {
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return bad(-2, "sync", e);
    if (sc->buf) cudaFree(sc->buf);
    sc->buf = nullptr;
    sc->cap = 0;
    if ((e = cudaMalloc(&sc->buf, bytes + bytes / 4)) != cudaSuccess) return bad(-3, "cudaMalloc", e);
    sc->cap = bytes + bytes / 4;
  }
  uint8_t* d = static_cast<uint8_t*>(sc->buf);
#define RA(call) \
  if ((e = (call)) != cudaSuccess) return bad(-2, #call, e);
  if (total > 0) {
    RA(cudaMemcpyAsync(d, pts1, size_t(total) * 2 * sizeof(float), cudaMemcpyHostToDevice, stream));
    RA(cudaMemcpyAsync(d + o_p2, pts2, size_t(total) * 2 * sizeof(float), cudaMemcpyHostToDevice, stream));
  }
  RA(cudaMemcpyAsync(d + o_off, off, size_t(n_pairs + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
  RansacArgs a{};
  fill_common(a, model, K, threshold_px, prob, max_iters, seed);
  a.pts1 = reinterpret_cast<const float*>(d);
  a.pts2 = reinterpret_cast<const float*>(d + o_p2);
  a.off = reinterpret_cast<const int*>(d + o_off);
  a.out_mask = d + o_mask;
  a.out_model = reinterpret_cast<float*>(d + o_model);
  a.out_inliers = reinterpret_cast<int*>(d + o_inl);
  RA(launch(a, n_pairs, max_n, stream));
  std::vector<float> h_model(size_t(n_pairs) * 9);
  if (total > 0) RA(cudaMemcpyAsync(out_mask, d + o_mask, size_t(total), cudaMemcpyDeviceToHost, stream));
  RA(cudaMemcpyAsync(h_model.data(), d + o_model, h_model.size() * sizeof(float), cudaMemcpyDeviceToHost, stream));
  RA(cudaMemcpyAsync(out_inliers, d + o_inl, size_t(n_pairs) * sizeof(int), cudaMemcpyDeviceToHost, stream));
  RA(cudaStreamSynchronize(stream));
#undef RA
  for (size_t i = 0; i < h_model.size(); ++i) out_model[i] = h_model[i];
  return 0;
}

int ransac_tables(int model, int* d_table, int* d_count, int cap, int n_pairs, const RedJob* d_jobs, const ImgDev* d_imgs,
                  const double* K, double threshold_px, double prob, int max_iters, uint32_t seed, int min_pairs,
                  bool compact, uint8_t* d_mask, float* d_model, int* d_inliers, cudaStream_t stream, std::string* err) {
  if (n_pairs == 0) return 0;
  RansacArgs a{};
  fill_common(a, model, K, threshold_px, prob, max_iters, seed);
  a.table = d_table;
  a.count = d_count;
  a.cap = cap;
  a.compact = compact ? 1 : 0;
  a.min_pairs = min_pairs;
  a.jobs = d_jobs;
  a.imgs = d_imgs;
  a.out_mask = d_mask;
  a.out_model = d_model;
  a.out_inliers = d_inliers;
  const cudaError_t e = launch(a, n_pairs, cap, stream);
  if (e != cudaSuccess) {
    if (err) *err = std::string("ransac launch: ") + cudaGetErrorString(e);
    return -2;
  }
  return 0;
}

}  // namespace iam
