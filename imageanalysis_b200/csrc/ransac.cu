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

#include <cmath>
#include <vector>

#include "ransac.h"

namespace iam {
namespace {

constexpr int kThreads = 256;
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
  // 6. roots of the degree-10 polynomial: Durand-Kerner on the monic form
  double mx = 0.0;
  for (int i = 0; i < 11; ++i) mx = fmax(mx, fabs(poly[i]));
  if (!(mx > 0.0) || !isfinite(mx)) return 0;
  int deg = 10;
  while (deg > 0 && fabs(poly[deg]) < 1e-13 * mx) --deg;
  if (deg < 1) return 0;
  double a[11];
  for (int i = 0; i <= deg; ++i) a[i] = poly[i] / poly[deg];
  double rad = 0.0;
  for (int i = 0; i < deg; ++i) rad = fmax(rad, fabs(a[i]));
  rad = fmin(1.0 + rad, 1e6);
  Cplx root[10];
  {
    Cplx w = {1.0, 0.0};
    const Cplx step = {0.4, 0.9};
    for (int i = 0; i < deg; ++i) {
      root[i] = {w.re * rad * 0.5, w.im * rad * 0.5};
      w = cmul(w, step);
      const double n = sqrt(w.re * w.re + w.im * w.im);
      w.re /= n;
      w.im /= n;
      w = cmul(w, Cplx{cos(0.37 * (i + 1)), sin(0.37 * (i + 1))});
    }
  }
  for (int it = 0; it < 120; ++it) {
    double delta = 0.0;
    for (int i = 0; i < deg; ++i) {
      Cplx p = {1.0, 0.0};
      for (int k = deg - 1; k >= 0; --k) {
        p = cmul(p, root[i]);
        p.re += a[k];
      }
      Cplx den = {1.0, 0.0};
      for (int j = 0; j < deg; ++j)
        if (j != i) den = cmul(den, csub(root[i], root[j]));
      if (den.re * den.re + den.im * den.im < 1e-300) continue;
      const Cplx d = cdiv(p, den);
      root[i] = csub(root[i], d);
      delta = fmax(delta, fabs(d.re) + fabs(d.im));
    }
    if (delta < 1e-13) break;
  }
  // 7. back-substitute every real root
  int n_out = 0;
  for (int i = 0; i < deg && n_out < kMaxCand; ++i) {
    if (fabs(root[i].im) > 1e-7 * (1.0 + fabs(root[i].re))) continue;
    double z = root[i].re;
    for (int nw = 0; nw < 3; ++nw) {  // Newton polish on the real polynomial
      double p = poly[deg], dp = 0.0;
      for (int k = deg - 1; k >= 0; --k) {
        dp = dp * z + p;
        p = p * z + poly[k];
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

struct PairXform {  // per-pair point normalisation (host computed)
  float a1x, a1y, s1x, s1y;  // image 1: xn = (x - a1x) * s1x
  float a2x, a2y, s2x, s2y;
  float thr2;                // squared threshold in normalised units (errors measured in image 2)
  float thr2b;               // fundamental: squared threshold for the error measured in image 1
  int pad[2];
};

__device__ __forceinline__ float model_error(int model, const float* M, float x1, float y1, float x2, float y2,
                                             float thr_ratio = 1.0f) {
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
    // distance in image 1 is rescaled to image-2 units (thr_ratio = thr2 / thr2b) and one comparison serves both.
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

__global__ void __launch_bounds__(kThreads)
ransac_kernel(int model, const float* __restrict__ pts1, const float* __restrict__ pts2, const int* __restrict__ off,
              const PairXform* __restrict__ xf, double prob, int max_iters, uint32_t seed,
              unsigned char* __restrict__ out_mask, float* __restrict__ out_model, int* __restrict__ out_inliers) {
  extern __shared__ float s_pts[];
  __shared__ float s_cand[kRound * kMaxCand * 9];
  __shared__ int s_ncand[kRound];
  __shared__ int s_score[kRound * kMaxCand];
  __shared__ float s_best[9];
  __shared__ int s_best_count, s_niters, s_done;

  const int p = blockIdx.x;
  const int o = off[p];
  const int n = off[p + 1] - o;
  const int msize = (model == 0) ? 5 : (model == 2) ? 7 : 4;
  const PairXform X = xf[p];
  const float thr_ratio = (model == 2 && X.thr2b > 0.f) ? X.thr2 / X.thr2b : 1.0f;
  float* sx1 = s_pts;
  float* sy1 = sx1 + n;
  float* sx2 = sy1 + n;
  float* sy2 = sx2 + n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    sx1[i] = (pts1[2 * (o + i)] - X.a1x) * X.s1x;
    sy1[i] = (pts1[2 * (o + i) + 1] - X.a1y) * X.s1y;
    sx2[i] = (pts2[2 * (o + i)] - X.a2x) * X.s2x;
    sy2[i] = (pts2[2 * (o + i) + 1] - X.a2y) * X.s2y;
  }
  if (threadIdx.x == 0) {
    s_best_count = 0;
    s_niters = max_iters;
    s_done = 0;
    for (int e = 0; e < 9; ++e) s_best[e] = 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (n >= msize) {
    for (int base = 0; base < max_iters; base += kRound) {
      if (warp == 0) {
        int nc = 0;
        const int it = base + lane;
        if (it < s_niters) {
          int s[7];
          uint32_t h = hash32(seed ^ hash32(uint32_t(p) * 0x9e3779b9u + uint32_t(it)));
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
                              : four_point(sx1, sy1, sx2, sy2, s, &s_cand[lane * kMaxCand * 9]);
        }
        s_ncand[lane] = nc;
      }
      __syncthreads();
      // score all candidates: one warp per candidate, lanes stride the points
      for (int c = warp; c < kRound * kMaxCand; c += kThreads / 32) {
        const int smp = c / kMaxCand, r = c % kMaxCand;
        if (r >= s_ncand[smp]) continue;
        float M[9];
        for (int e = 0; e < 9; ++e) M[e] = s_cand[c * 9 + e];
        int cnt = 0;
        for (int i = lane; i < n; i += 32) cnt += model_error(model, M, sx1[i], sy1[i], sx2[i], sy2[i], thr_ratio) <= X.thr2;
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        if (lane == 0) s_score[c] = cnt;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int smp = 0; smp < kRound && base + smp < s_niters; ++smp) {
          for (int r = 0; r < s_ncand[smp]; ++r) {
            const int cnt = s_score[smp * kMaxCand + r];
            if (cnt > max(s_best_count, msize - 1)) {
              s_best_count = cnt;
              for (int e = 0; e < 9; ++e) s_best[e] = s_cand[(smp * kMaxCand + r) * 9 + e];
              // adaptive iteration count (the standard RANSAC bound)
              const double ep = fmin(fmax(1.0 - double(cnt) / double(n), 0.0), 1.0);
              const double num = log(fmax(1.0 - prob, 1e-300));
              const double den = 1.0 - pow(1.0 - ep, double(msize));
              int ni = max_iters;
              if (den < 1e-300)
                ni = 0;
              else {
                const double lden = log(den);
                if (!(lden >= 0.0) && -num < double(max_iters) * (-lden)) ni = int(rint(num / lden));
              }
              s_niters = min(s_niters, ni);
            }
          }
        }
        s_done = (base + kRound >= s_niters);
      }
      __syncthreads();
      if (s_done) break;
    }
  }
  // final mask with the best model
  float M[9];
  for (int e = 0; e < 9; ++e) M[e] = s_best[e];
  const bool have = s_best_count > 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    out_mask[o + i] = (have && model_error(model, M, sx1[i], sy1[i], sx2[i], sy2[i], thr_ratio) <= X.thr2) ? 1 : 0;
  if (threadIdx.x < 9) out_model[p * 9 + threadIdx.x] = have ? M[threadIdx.x] : 0.f;
  if (threadIdx.x == 0) out_inliers[p] = have ? s_best_count : 0;
}

}  // namespace

int debug_minimal_solver(int model, const float* x1, const float* y1, const float* x2, const float* y2, float* out) {
  const int s[7] = {0, 1, 2, 3, 4, 5, 6};
  return model == 0 ? five_point(x1, y1, x2, y2, s, out) : model == 2 ? seven_point(x1, y1, x2, y2, s, out)
                                                                        : four_point(x1, y1, x2, y2, s, out);
}

int ransac_pairs(int model, const float* pts1, const float* pts2, const int32_t* off, int n_pairs, const double* K,
                 double threshold_px, double prob, int max_iters, uint32_t seed, uint8_t* out_mask, double* out_model,
                 int32_t* out_inliers, cudaStream_t stream, std::string* err) {
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
  if (max_iters <= 0) max_iters = 1000;
  if (!(prob > 0.0 && prob < 1.0)) prob = 0.999;

  // per-pair normalisation
  std::vector<PairXform> xf(n_pairs);
  for (int p = 0; p < n_pairs; ++p) {
    PairXform& X = xf[p];
    const int o = off[p], n = off[p + 1] - o;
    if (model == 0) {
      // K^-1 with fx, fy, cx, cy; threshold / mean focal (how OpenCV's findEssentialMat treats pixel thresholds)
      const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
      X.a1x = X.a2x = float(cx);
      X.a1y = X.a2y = float(cy);
      X.s1x = X.s2x = float(1.0 / fx);
      X.s1y = X.s2y = float(1.0 / fy);
      const double t = threshold_px / ((fx + fy) / 2.0);
      X.thr2 = float(t * t);
    } else {
      // Hartley normalisation; the transfer error lives in image 2, so its isotropic scale carries the threshold
      double m1x = 0, m1y = 0, m2x = 0, m2y = 0;
      for (int i = 0; i < n; ++i) {
        m1x += pts1[2 * (o + i)]; m1y += pts1[2 * (o + i) + 1];
        m2x += pts2[2 * (o + i)]; m2y += pts2[2 * (o + i) + 1];
      }
      const double inv = n > 0 ? 1.0 / n : 0.0;
      m1x *= inv; m1y *= inv; m2x *= inv; m2y *= inv;
      double d1 = 0, d2 = 0;
      for (int i = 0; i < n; ++i) {
        d1 += std::hypot(pts1[2 * (o + i)] - m1x, pts1[2 * (o + i) + 1] - m1y);
        d2 += std::hypot(pts2[2 * (o + i)] - m2x, pts2[2 * (o + i) + 1] - m2y);
      }
      const double s1 = d1 > 0 ? std::sqrt(2.0) * n / d1 : 1.0;
      const double s2 = d2 > 0 ? std::sqrt(2.0) * n / d2 : 1.0;
      X.a1x = float(m1x); X.a1y = float(m1y); X.s1x = X.s1y = float(s1);
      X.a2x = float(m2x); X.a2y = float(m2y); X.s2x = X.s2y = float(s2);
      const double t = threshold_px * double(X.s2x);
      X.thr2 = float(t * t);
      const double tb = threshold_px * double(X.s1x);
      X.thr2b = float(tb * tb);
    }
  }

  float *d_p1 = nullptr, *d_p2 = nullptr, *d_model = nullptr;
  int *d_off = nullptr, *d_inl = nullptr;
  unsigned char* d_mask = nullptr;
  PairXform* d_xf = nullptr;
  cudaError_t e;
  const size_t pb = size_t(std::max(total, 1)) * 2 * sizeof(float);
#define RA(call)                                  \
  if ((e = (call)) != cudaSuccess) {              \
    cudaFree(d_p1); cudaFree(d_p2); cudaFree(d_model); cudaFree(d_off); cudaFree(d_inl); cudaFree(d_mask); cudaFree(d_xf); \
    return bad(-2, #call, e);                     \
  }
  RA(cudaMalloc(&d_p1, pb));
  RA(cudaMalloc(&d_p2, pb));
  RA(cudaMalloc(&d_model, size_t(n_pairs) * 9 * sizeof(float)));
  RA(cudaMalloc(&d_off, size_t(n_pairs + 1) * sizeof(int)));
  RA(cudaMalloc(&d_inl, size_t(n_pairs) * sizeof(int)));
  RA(cudaMalloc(&d_mask, size_t(std::max(total, 1))));
  RA(cudaMalloc(&d_xf, size_t(n_pairs) * sizeof(PairXform)));
  if (total > 0) {
    RA(cudaMemcpyAsync(d_p1, pts1, size_t(total) * 2 * sizeof(float), cudaMemcpyHostToDevice, stream));
    RA(cudaMemcpyAsync(d_p2, pts2, size_t(total) * 2 * sizeof(float), cudaMemcpyHostToDevice, stream));
  }
  RA(cudaMemcpyAsync(d_off, off, size_t(n_pairs + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
  RA(cudaMemcpyAsync(d_xf, xf.data(), size_t(n_pairs) * sizeof(PairXform), cudaMemcpyHostToDevice, stream));
  const size_t smem = size_t(std::max(max_n, 1)) * 4 * sizeof(float);
  if (smem > 48 * 1024) RA(cudaFuncSetAttribute(ransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ransac_kernel<<<n_pairs, kThreads, smem, stream>>>(model, d_p1, d_p2, d_off, d_xf, prob, max_iters, seed, d_mask,
                                                     d_model, d_inl);
  RA(cudaGetLastError());
  std::vector<float> h_model(size_t(n_pairs) * 9);
  if (total > 0) RA(cudaMemcpyAsync(out_mask, d_mask, size_t(total), cudaMemcpyDeviceToHost, stream));
  RA(cudaMemcpyAsync(h_model.data(), d_model, h_model.size() * sizeof(float), cudaMemcpyDeviceToHost, stream));
  RA(cudaMemcpyAsync(out_inliers, d_inl, size_t(n_pairs) * sizeof(int), cudaMemcpyDeviceToHost, stream));
  RA(cudaStreamSynchronize(stream));
#undef RA
  cudaFree(d_p1); cudaFree(d_p2); cudaFree(d_model); cudaFree(d_off); cudaFree(d_inl); cudaFree(d_mask); cudaFree(d_xf);

  // back to caller units: E is reported for normalised image coordinates like
  // cv2.findEssentialMat does (x2n^T E x1n = 0); H is mapped back to pixels.
  for (int p = 0; p < n_pairs; ++p) {
    const float* m = &h_model[size_t(p) * 9];
    double* o = out_model + size_t(p) * 9;
    if (model == 0) {
      for (int i = 0; i < 9; ++i) o[i] = m[i];
    } else if (model == 2) {
      // F_pixel = T2^T Fn T1, scaled to unit f33 like cv2.findFundamentalMat reports it
      const PairXform& X = xf[p];
      const double T1[9] = {X.s1x, 0, -X.s1x * X.a1x, 0, X.s1y, -X.s1y * X.a1y, 0, 0, 1};
      const double T2[9] = {X.s2x, 0, -X.s2x * X.a2x, 0, X.s2y, -X.s2y * X.a2y, 0, 0, 1};
      double t[9], r[9];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          t[i * 3 + j] = 0;
          for (int k = 0; k < 3; ++k) t[i * 3 + j] += double(m[i * 3 + k]) * T1[k * 3 + j];
        }
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          r[i * 3 + j] = 0;
          for (int k = 0; k < 3; ++k) r[i * 3 + j] += T2[k * 3 + i] * t[k * 3 + j];
        }
      const double sc = std::fabs(r[8]) > 1e-300 ? 1.0 / r[8] : 1.0;
      for (int i = 0; i < 9; ++i) o[i] = r[i] * sc;
    } else {
      const PairXform& X = xf[p];
      // Hp = T2^-1 Hn T1, T = [[s,0,-s*a],[0,s,-s*b],[0,0,1]]
      const double T1[9] = {X.s1x, 0, -X.s1x * X.a1x, 0, X.s1y, -X.s1y * X.a1y, 0, 0, 1};
      const double T2i[9] = {1.0 / X.s2x, 0, X.a2x, 0, 1.0 / X.s2y, X.a2y, 0, 0, 1};
      double t[9], r[9];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          t[i * 3 + j] = 0;
          for (int k = 0; k < 3; ++k) t[i * 3 + j] += double(m[i * 3 + k]) * T1[k * 3 + j];
        }
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          r[i * 3 + j] = 0;
          for (int k = 0; k < 3; ++k) r[i * 3 + j] += T2i[i * 3 + k] * t[k * 3 + j];
        }
      const double s = std::fabs(r[8]) > 1e-300 ? 1.0 / r[8] : 1.0;
      for (int i = 0; i < 9; ++i) o[i] = r[i] * s;
    }
  }
  return 0;
}

}  // namespace iam
