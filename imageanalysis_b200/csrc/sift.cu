// sift.cu — SIFT detect + describe on the GPU, standing in for
//   detector = cv2.SIFT_create(); detector.detectAndCompute(scaled, None)
// (reference scripts/lib/image.py:236-237, :324; the default detector of the pipeline).  OpenCV's defaults: 3 layers
// per octave, sigma 1.6, contrast threshold 0.04, edge threshold 10, first octave -1 (the image is doubled).
// The method is Lowe's (IJCV 2004) as OpenCV implements it; the CPU checker of the test suite restates it stage by
// stage and is pinned against live cv2:
//   base image    grey -> float, x2 bilinear, Gaussian blur to sigma 1.6
//   pyramid       per octave 6 Gaussian layers (incremental separable float blurs, reflect-101 borders), next octave =
//                 every second pixel of layer 3; 5 difference-of-Gaussian layers
//   extrema       |D| > 1 and >= / <= all 26 neighbours; up to 5 Newton steps of the 3-D quadratic fit (Cramer's rule
//                 in float, as cv::Matx33f::solve), contrast and edge tests
//   orientation   36-bin histogram of gradient directions (OpenCV's polynomial fastAtan2, Gaussian weights), smoothed,
//                 every peak >= 0.8 max becomes a key point (parabolic peak position)
//   descriptor    4 x 4 x 8 histogram with tri-linear interpolation, clipped at 0.2, scaled by 512, saturated to 8 bits
// Histograms are accumulated by ONE thread per key point in OpenCV's loop order, and every float expression is written
// with explicit round-to-nearest intrinsics (no FMA contraction), so the result does not depend on scheduling.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>

#include "sift.h"

namespace iam {
namespace {

constexpr int kNL = 3;            // nOctaveLayers
constexpr int kImgBorder = 5;     // SIFT_IMG_BORDER
constexpr int kMaxSteps = 5;      // SIFT_MAX_INTERP_STEPS
constexpr int kOriBins = 36;
constexpr int kMaxOctaves = 16;
constexpr int kMaxKernel = 64;
#ifndef IAM_SIFT_GENERIC_BLUR
#define IAM_SIFT_GENERIC_BLUR 0
#endif
#ifndef IAM_SIFT_SEQ_ORI
#define IAM_SIFT_SEQ_ORI 0
#endif
#ifndef IAM_SIFT_SEQ_DESC
#define IAM_SIFT_SEQ_DESC 0
#endif

struct Octave {
  int w, h;
  size_t gauss[kNL + 3];   // float offsets into the pyramid block
  size_t dog[kNL + 2];
};
struct Pyramid {
  int n_oct;
  Octave oct[kMaxOctaves];
};
struct BlurKernel {
  int ksize;
  float k[kMaxKernel];
};

__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}

// grey u8 -> float, doubled with bilinear interpolation (cv::resize INTER_LINEAR on the float image)
__global__ void up2_kernel(const uint8_t* __restrict__ src, int w, int h, float* __restrict__ dst) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y * blockDim.y + threadIdx.y;
  if (dx >= 2 * w || dy >= 2 * h) return;
  auto coef = [](int d, int sn, int& s, float& a) {
    const float f = (d + 0.5f) * 0.5f - 0.5f;   // exact in float for these sizes
    int si = (int)floorf(f);
    float fr = f - (float)si;
    if (si < 0) {
      si = 0;
      fr = 0.f;
    }
    if (si >= sn - 1) {
      si = sn - 1;
      fr = 0.f;
    }
    s = si;
    a = fr;
  };
  int xi, yi;
  float xa, ya;
  coef(dx, w, xi, xa);
  coef(dy, h, yi, ya);
  const int x1 = min(xi + 1, w - 1), y1 = min(yi + 1, h - 1);
  const float a0 = __fsub_rn(1.f, xa), b0 = __fsub_rn(1.f, ya);
  const float h0 = __fadd_rn(__fmul_rn((float)src[(size_t)yi * w + xi], a0), __fmul_rn((float)src[(size_t)yi * w + x1], xa));
  const float h1 = __fadd_rn(__fmul_rn((float)src[(size_t)y1 * w + xi], a0), __fmul_rn((float)src[(size_t)y1 * w + x1], xa));
  dst[(size_t)dy * (2 * w) + dx] = __fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, ya));
}

// separable float Gaussian, taps accumulated in order (reflect-101 borders)
__global__ void blur_rows_kernel(const float* __restrict__ src, int w, int h, BlurKernel K, float* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const int r = K.ksize >> 1;
  const float* row = src + (size_t)y * w;
  float s = 0.f;
  if (x >= r && x < w - r) {
    for (int i = 0; i < K.ksize; ++i) s = __fadd_rn(s, __fmul_rn(K.k[i], row[x - r + i]));
  } else {
    for (int i = 0; i < K.ksize; ++i) s = __fadd_rn(s, __fmul_rn(K.k[i], row[reflect101(x - r + i, w)]));
  }
  dst[(size_t)y * w + x] = s;
}
// column pass; with `prev` also writes the difference-of-Gaussian layer dog = dst - prev
__global__ void blur_cols_kernel(const float* __restrict__ src, int w, int h, BlurKernel K, float* __restrict__ dst,
                                 const float* __restrict__ prev, float* __restrict__ dog) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const int r = K.ksize >> 1;
  float s = 0.f;
  if (y >= r && y < h - r) {
    for (int i = 0; i < K.ksize; ++i) s = __fadd_rn(s, __fmul_rn(K.k[i], src[(size_t)(y - r + i) * w + x]));
  } else {
    for (int i = 0; i < K.ksize; ++i) s = __fadd_rn(s, __fmul_rn(K.k[i], src[(size_t)reflect101(y - r + i, h) * w + x]));
  }
  const size_t at = (size_t)y * w + x;
  dst[at] = s;
  if (dog) dog[at] = __fsub_rn(s, prev[at]);
}
// The same two passes for the radii SIFT's six sigmas produce (5, 6, 8, 10, 13), unrolled: a row block stages 512 + halo
// pixels in shared memory and every thread sums 4 neighbouring outputs from one register window (aligned 128-bit
// reads); a column thread keeps 8 running sums so that every loaded pixel feeds 8 outputs.  Tap order per output is
// unchanged (i = 0 .. 2R), so the results are bit-identical to the generic kernels above.
constexpr int kRowTile = 512, kRowHalo = 16, kColRows = 8;
template <int R>
__global__ void __launch_bounds__(kRowTile / 4) blur_rows_t(const float* __restrict__ src, int w, int h, BlurKernel K,
                                                          float* __restrict__ dst) {
  static_assert(R <= kRowHalo, "halo too small");
  __shared__ __align__(16) float tile[kRowTile + 2 * kRowHalo];
  const int y = blockIdx.y, x0 = blockIdx.x * kRowTile, t = threadIdx.x;
  const float* row = src + (size_t)y * w;
  for (int s = t; s < kRowTile + 2 * kRowHalo; s += kRowTile / 4) {
    const int x = x0 - kRowHalo + s;
    tile[s] = (x >= 0 && x < w) ? row[x] : row[reflect101(min(x, 2 * w + kRowHalo), w)];
  }
  __syncthreads();
  constexpr int kLead = (kRowHalo - R) & ~3, kMis = (kRowHalo - R) & 3, kVec = (kMis + 2 * R + 4 + 3) / 4;
  float v[4 * kVec];
#pragma unroll
  for (int j = 0; j < kVec; ++j) {
    const float4 q = *reinterpret_cast<const float4*>(&tile[4 * t + kLead + 4 * j]);
    v[4 * j] = q.x;
    v[4 * j + 1] = q.y;
    v[4 * j + 2] = q.z;
    v[4 * j + 3] = q.w;
  }
  float out[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i <= 2 * R; ++i) a = __fadd_rn(a, __fmul_rn(K.k[i], v[kMis + q + i]));
    out[q] = a;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) tile[4 * t + q] = out[q];
  __syncthreads();
  float* orow = dst + (size_t)y * w;
  for (int s = t; s < kRowTile; s += kRowTile / 4)
    if (x0 + s < w) orow[x0 + s] = tile[s];
}

template <int R>
__global__ void __launch_bounds__(128) blur_cols_t(const float* __restrict__ src, int w, int h, BlurKernel K,
                                                   float* __restrict__ dst, const float* __restrict__ prev,
                                                   float* __restrict__ dog) {
  const int x = blockIdx.x * 128 + threadIdx.x, y0 = blockIdx.y * kColRows;
  if (x >= w) return;
  float acc[kColRows];
#pragma unroll
  for (int j = 0; j < kColRows; ++j) acc[j] = 0.f;
  const bool inside = y0 - R >= 0 && y0 + kColRows - 1 + R < h;
#pragma unroll
  for (int i = 0; i < 2 * R + kColRows; ++i) {
    const int yy = y0 - R + i;
    const float val = src[(size_t)(inside ? yy : reflect101(min(yy, 2 * h + R), h)) * w + x];
#pragma unroll
    for (int j = 0; j < kColRows; ++j) {
      const int tap = i - j;
      if (tap >= 0 && tap <= 2 * R) acc[j] = __fadd_rn(acc[j], __fmul_rn(K.k[tap], val));
    }
  }
#pragma unroll
  for (int j = 0; j < kColRows; ++j) {
    const int y = y0 + j;
    if (y < h) {
      const size_t at = (size_t)y * w + x;
      dst[at] = acc[j];
      if (dog) dog[at] = __fsub_rn(acc[j], prev[at]);
    }
  }
}

template <int R>
void launch_blur_t(const float* src, float* tmp, float* dst, int W, int H, const BlurKernel& K, float* dog, cudaStream_t stream) {
  blur_rows_t<R><<<dim3((W + kRowTile - 1) / kRowTile, H), kRowTile / 4, 0, stream>>>(src, W, H, K, tmp);
  blur_cols_t<R><<<dim3((W + 127) / 128, (H + kColRows - 1) / kColRows), 128, 0, stream>>>(tmp, W, H, K, dst, src, dog);
}

__global__ void half_kernel(const float* __restrict__ src, int sw, float* __restrict__ dst, int w, int h) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x < w && y < h) dst[(size_t)y * w + x] = src[(size_t)(2 * y) * sw + 2 * x];
}

struct Cand {   // a refined scale-space extremum
  float x, y, size, response;   // in base-image units (not yet halved)
  int octave_field;             // octave + (layer << 8) + (xi code << 16)
  short o, layer;
  int r, c;
};
struct KeyOut {
  float x, y, size, angle, response;
  int octave_field;
  short o, layer;
};

__device__ __forceinline__ float at(const float* img, int w, int r, int c) { return img[(size_t)r * w + c]; }

// adjustLocalExtrema of sift.simd.hpp: Newton steps on the 3-D quadratic, contrast / edge rejection
__device__ bool adjust_extremum(const float* base, const Octave& O, int o, int layer, int r, int c, Cand& out) {
  const float img_scale = 1.f / 255.f;
  const float ds = __fmul_rn(img_scale, 0.5f), ss = img_scale, cs = __fmul_rn(img_scale, 0.25f);
  const int w = O.w, h = O.h;
  float xi = 0.f, xr = 0.f, xc = 0.f;
  int i = 0;
  for (; i < kMaxSteps; ++i) {
    const float* img = base + O.dog[layer];
    const float* prv = base + O.dog[layer - 1];
    const float* nxt = base + O.dog[layer + 1];
    const float dD0 = __fmul_rn(__fsub_rn(at(img, w, r, c + 1), at(img, w, r, c - 1)), ds);
    const float dD1 = __fmul_rn(__fsub_rn(at(img, w, r + 1, c), at(img, w, r - 1, c)), ds);
    const float dD2 = __fmul_rn(__fsub_rn(at(nxt, w, r, c), at(prv, w, r, c)), ds);
    const float v2 = __fmul_rn(at(img, w, r, c), 2.f);
    const float dxx = __fmul_rn(__fsub_rn(__fadd_rn(at(img, w, r, c + 1), at(img, w, r, c - 1)), v2), ss);
    const float dyy = __fmul_rn(__fsub_rn(__fadd_rn(at(img, w, r + 1, c), at(img, w, r - 1, c)), v2), ss);
    const float dss = __fmul_rn(__fsub_rn(__fadd_rn(at(nxt, w, r, c), at(prv, w, r, c)), v2), ss);
    const float dxy = __fmul_rn(__fadd_rn(__fsub_rn(__fsub_rn(at(img, w, r + 1, c + 1), at(img, w, r + 1, c - 1)), at(img, w, r - 1, c + 1)), at(img, w, r - 1, c - 1)), cs);
    const float dxs = __fmul_rn(__fadd_rn(__fsub_rn(__fsub_rn(at(nxt, w, r, c + 1), at(nxt, w, r, c - 1)), at(prv, w, r, c + 1)), at(prv, w, r, c - 1)), cs);
    const float dys = __fmul_rn(__fadd_rn(__fsub_rn(__fsub_rn(at(nxt, w, r + 1, c), at(nxt, w, r - 1, c)), at(prv, w, r + 1, c)), at(prv, w, r - 1, c)), cs);
    // X = H^-1 dD by Cramer's rule in float (cv::Matx33f::solve, DECOMP_LU resolves to the closed form for 3 x 3)
    const float a00 = dxx, a01 = dxy, a02 = dxs, a10 = dxy, a11 = dyy, a12 = dys, a20 = dxs, a21 = dys, a22 = dss;
    auto m2 = [](float p, float q, float s, float t) { return __fsub_rn(__fmul_rn(p, q), __fmul_rn(s, t)); };
    float det = __fadd_rn(__fsub_rn(__fmul_rn(a00, m2(a11, a22, a21, a12)), __fmul_rn(a01, m2(a10, a22, a20, a12))),
                          __fmul_rn(a02, m2(a10, a21, a20, a11)));
    float X0 = 0.f, X1 = 0.f, X2 = 0.f;
    if (det != 0.f) {
      const float d = __fdiv_rn(1.f, det);
      X0 = __fmul_rn(d, __fadd_rn(__fsub_rn(__fmul_rn(dD0, m2(a11, a22, a12, a21)), __fmul_rn(a01, m2(dD1, a22, a12, dD2))),
                                  __fmul_rn(a02, m2(dD1, a21, a11, dD2))));
      X1 = __fmul_rn(d, __fadd_rn(__fsub_rn(__fmul_rn(a00, m2(dD1, a22, a12, dD2)), __fmul_rn(dD0, m2(a10, a22, a12, a20))),
                                  __fmul_rn(a02, m2(a10, dD2, dD1, a20))));
      X2 = __fmul_rn(d, __fadd_rn(__fsub_rn(__fmul_rn(a00, m2(a11, dD2, dD1, a21)), __fmul_rn(a01, m2(a10, dD2, dD1, a20))),
                                  __fmul_rn(dD0, m2(a10, a21, a11, a20))));
    }
    xi = -X2;
    xr = -X1;
    xc = -X0;
    if (fabsf(xi) < 0.5f && fabsf(xr) < 0.5f && fabsf(xc) < 0.5f) break;
    if (fabsf(xi) > 715827882.f || fabsf(xr) > 715827882.f || fabsf(xc) > 715827882.f) return false;
    c += __float2int_rn(xc);
    r += __float2int_rn(xr);
    layer += __float2int_rn(xi);
    if (layer < 1 || layer > kNL || c < kImgBorder || c >= w - kImgBorder || r < kImgBorder || r >= h - kImgBorder) return false;
  }
  if (i >= kMaxSteps) return false;
  const float* img = base + O.dog[layer];
  const float* prv = base + O.dog[layer - 1];
  const float* nxt = base + O.dog[layer + 1];
  const float dD0 = __fmul_rn(__fsub_rn(at(img, w, r, c + 1), at(img, w, r, c - 1)), ds);
  const float dD1 = __fmul_rn(__fsub_rn(at(img, w, r + 1, c), at(img, w, r - 1, c)), ds);
  const float dD2 = __fmul_rn(__fsub_rn(at(nxt, w, r, c), at(prv, w, r, c)), ds);
  const float t = __fadd_rn(__fadd_rn(__fmul_rn(dD0, xc), __fmul_rn(dD1, xr)), __fmul_rn(dD2, xi));
  const float contr = __fadd_rn(__fmul_rn(at(img, w, r, c), img_scale), __fmul_rn(t, 0.5f));
  if (__fmul_rn(fabsf(contr), (float)kNL) < 0.04f) return false;
  const float v2 = __fmul_rn(at(img, w, r, c), 2.f);
  const float dxx = __fmul_rn(__fsub_rn(__fadd_rn(at(img, w, r, c + 1), at(img, w, r, c - 1)), v2), ss);
  const float dyy = __fmul_rn(__fsub_rn(__fadd_rn(at(img, w, r + 1, c), at(img, w, r - 1, c)), v2), ss);
  const float dxy = __fmul_rn(__fadd_rn(__fsub_rn(__fsub_rn(at(img, w, r + 1, c + 1), at(img, w, r + 1, c - 1)), at(img, w, r - 1, c + 1)), at(img, w, r - 1, c - 1)), cs);
  const float tr = __fadd_rn(dxx, dyy);
  const float det = __fsub_rn(__fmul_rn(dxx, dyy), __fmul_rn(dxy, dxy));
  if (det <= 0.f || __fmul_rn(__fmul_rn(tr, tr), 10.f) >= __fmul_rn(121.f, det)) return false;
  const float sc = (float)(1 << o);
  out.x = __fmul_rn(__fadd_rn((float)c, xc), sc);
  out.y = __fmul_rn(__fadd_rn((float)r, xr), sc);
  out.octave_field = o + (layer << 8) + (__double2int_rn(((double)xi + 0.5) * 255.0) << 16);
  out.size = __fmul_rn(__fmul_rn(__fmul_rn(1.6f, (float)pow(2.0, (double)__fdiv_rn(__fadd_rn((float)layer, xi), (float)kNL))), sc), 2.f);
  out.response = fabsf(contr);
  out.o = (short)o;
  out.layer = (short)layer;
  out.r = r;
  out.c = c;
  return true;
}

__global__ void extrema_kernel(const float* __restrict__ base, Pyramid P, int o, Cand* __restrict__ cand,
                               int* __restrict__ n_cand, int cap) {
  const Octave O = P.oct[o];
  const int layer = blockIdx.z + 1;   // the three inner DoG layers of the octave in one launch
  const int c = blockIdx.x * blockDim.x + threadIdx.x + kImgBorder, r = blockIdx.y * blockDim.y + threadIdx.y + kImgBorder;
  if (c >= O.w - kImgBorder || r >= O.h - kImgBorder) return;
  const int w = O.w;
  const float* img = base + O.dog[layer];
  const float v = at(img, w, r, c);
  if (!(fabsf(v) > 1.0f)) return;       // threshold = floor(0.5 * 0.04 / 3 * 255) = 1
  const float* L[3] = {base + O.dog[layer - 1], img, base + O.dog[layer + 1]};
  bool is_max = v > 0.f, is_min = v < 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
      for (int dc = -1; dc <= 1; ++dc) {
        const float nb = at(L[k], w, r + dr, c + dc);
        is_max = is_max && v >= nb;
        is_min = is_min && v <= nb;
      }
  if (!is_max && !is_min) return;
  Cand cd;
  if (!adjust_extremum(base, O, o, layer, r, c, cd)) return;
  const int pos = atomicAdd(n_cand, 1);
  if (pos < cap) cand[pos] = cd;
}

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float p1 = 0.9997878412794807f * 57.29577951308232f, p3 = -0.3258083974640975f * 57.29577951308232f;
  const float p5 = 0.1555786518463281f * 57.29577951308232f, p7 = -0.04432655554792128f * 57.29577951308232f;
  const float eps = 2.220446049250313e-16f;
  const float ax = fabsf(x), ay = fabsf(y);
  float a;
  if (ax >= ay) {
    const float c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    const float c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    const float c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    const float c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.0f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0.f) a = __fsub_rn(180.0f, a);
  if (y < 0.f) a = __fsub_rn(360.0f, a);
  return a;
}

// calcOrientationHist + the peak loop of findScaleSpaceExtrema, one thread per refined extremum in OpenCV's loop order:
// the A/B reference of orientation_kernel (-DIAM_SIFT_SEQ_ORI=1 selects it)
__global__ void orientation_seq_kernel(const float* __restrict__ base, Pyramid P, const Cand* __restrict__ cand,
                                   const int* __restrict__ n_cand, int cand_cap, KeyOut* __restrict__ keys,
                                   int* __restrict__ n_keys, int key_cap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= min(*n_cand, cand_cap)) return;
  const Cand cd = cand[i];
  const Octave O = P.oct[cd.o];
  const float* img = base + O.gauss[cd.layer];
  const int w = O.w, h = O.h, n = kOriBins;
  const float scl_octv = __fdiv_rn(__fmul_rn(cd.size, 0.5f), (float)(1 << cd.o));
  const int radius = __float2int_rn(__fmul_rn(4.5f, scl_octv));
  const float sigma = __fmul_rn(1.5f, scl_octv);
  const float expf_scale = __fdiv_rn(-1.f, __fmul_rn(__fmul_rn(2.f, sigma), sigma));
  float temphist[kOriBins];
  for (int k = 0; k < n; ++k) temphist[k] = 0.f;
  for (int di = -radius; di <= radius; ++di) {
    const int y = cd.r + di;
    if (y <= 0 || y >= h - 1) continue;
    for (int dj = -radius; dj <= radius; ++dj) {
      const int x = cd.c + dj;
      if (x <= 0 || x >= w - 1) continue;
      const float dx = __fsub_rn(at(img, w, y, x + 1), at(img, w, y, x - 1));
      const float dy = __fsub_rn(at(img, w, y - 1, x), at(img, w, y + 1, x));
      const float wgt = expf(__fmul_rn((float)(di * di + dj * dj), expf_scale));
      const float ori = fast_atan2_deg(dy, dx);
      const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      int bin = __float2int_rn(__fmul_rn(0.1f, ori));   // (n / 360.f) * Ori
      if (bin >= n) bin -= n;
      if (bin < 0) bin += n;
      temphist[bin] = __fadd_rn(temphist[bin], __fmul_rn(wgt, mag));
    }
  }
  float hist[kOriBins];
  float omax = 0.f;
  for (int k = 0; k < n; ++k) {
    const float m2 = temphist[(k + n - 2) % n], m1 = temphist[(k + n - 1) % n], p1 = temphist[(k + 1) % n], p2 = temphist[(k + 2) % n];
    hist[k] = __fadd_rn(__fadd_rn(__fmul_rn(__fadd_rn(m2, p2), 1.f / 16.f), __fmul_rn(__fadd_rn(m1, p1), 4.f / 16.f)),
                        __fmul_rn(temphist[k], 6.f / 16.f));
    omax = k == 0 ? hist[0] : fmaxf(omax, hist[k]);
  }
  const float mag_thr = __fmul_rn(omax, 0.8f);
  for (int j = 0; j < n; ++j) {
    const int l = j > 0 ? j - 1 : n - 1, r2 = j < n - 1 ? j + 1 : 0;
    if (hist[j] > hist[l] && hist[j] > hist[r2] && hist[j] >= mag_thr) {
      float bin = __fadd_rn((float)j, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(hist[l], hist[r2])),
                                                __fadd_rn(__fsub_rn(hist[l], __fmul_rn(2.f, hist[j])), hist[r2])));
      bin = bin < 0.f ? __fadd_rn((float)n, bin) : (bin >= (float)n ? __fsub_rn(bin, (float)n) : bin);
      float ang = __fsub_rn(360.f, __fmul_rn(10.f, bin));
      if (fabsf(__fsub_rn(ang, 360.f)) < 1.1920929e-07f) ang = 0.f;
      const int pos = atomicAdd(n_keys, 1);
      if (pos < key_cap) {
        KeyOut k;
        k.x = cd.x; k.y = cd.y; k.size = cd.size; k.angle = ang; k.response = cd.response;
        k.octave_field = cd.octave_field; k.o = cd.o; k.layer = cd.layer;
        keys[pos] = k;
      }
    }
  }
}

// calcOrientationHist + the peak loop of findScaleSpaceExtrema, one WARP per refined extremum (persistent warps stride
// the candidate list, whose length only the device knows).  The lanes share the (2 radius + 1)^2 window; weighted
// magnitudes are added to the 36-bin histogram as 64-bit fixed point (2^-40), so the sums are exact and do not depend
// on the order the lanes arrive in; lane 0 smooths the histogram and emits one key point per peak.
constexpr int kOriWarps = 4;
__global__ void __launch_bounds__(kOriWarps * 32) orientation_kernel(const float* __restrict__ base, Pyramid P,
                                                                    const Cand* __restrict__ cand, const int* __restrict__ n_cand,
                                                                    int cand_cap, KeyOut* __restrict__ keys,
                                                                    int* __restrict__ n_keys, int key_cap) {
  __shared__ unsigned long long hist_s[kOriWarps][kOriBins];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = min(*n_cand, cand_cap);
  unsigned long long* hq = hist_s[warp];
  for (int i = blockIdx.x * kOriWarps + warp; i < total; i += gridDim.x * kOriWarps) {
    for (int k = lane; k < kOriBins; k += 32) hq[k] = 0ull;
    __syncwarp();
    const Cand cd = cand[i];
    const Octave O = P.oct[cd.o];
    const float* img = base + O.gauss[cd.layer];
    const int w = O.w, h = O.h, n = kOriBins;
    const float scl_octv = __fdiv_rn(__fmul_rn(cd.size, 0.5f), (float)(1 << cd.o));
    const int radius = __float2int_rn(__fmul_rn(4.5f, scl_octv));
    const float sigma = __fmul_rn(1.5f, scl_octv);
    const float expf_scale = __fdiv_rn(-1.f, __fmul_rn(__fmul_rn(2.f, sigma), sigma));
    const int side = 2 * radius + 1;
    for (int s = lane; s < side * side; s += 32) {
      const int di = s / side - radius, dj = s % side - radius;
      const int y = cd.r + di, x = cd.c + dj;
      if (y <= 0 || y >= h - 1 || x <= 0 || x >= w - 1) continue;
      const float dx = __fsub_rn(at(img, w, y, x + 1), at(img, w, y, x - 1));
      const float dy = __fsub_rn(at(img, w, y - 1, x), at(img, w, y + 1, x));
      const float wgt = expf(__fmul_rn((float)(di * di + dj * dj), expf_scale));
      const float ori = fast_atan2_deg(dy, dx);
      const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      int bin = __float2int_rn(__fmul_rn(0.1f, ori));   // (n / 360.f) * Ori
      if (bin >= n) bin -= n;
      if (bin < 0) bin += n;
      atomicAdd(&hq[bin], (unsigned long long)__float2ll_rn(__fmul_rn(__fmul_rn(wgt, mag), 1099511627776.f)));
    }
    __syncwarp();
    if (lane == 0) {
      float temphist[kOriBins], hist[kOriBins];
      for (int k = 0; k < n; ++k) temphist[k] = __fmul_rn(__ll2float_rn((long long)hq[k]), 9.094947017729282e-13f);
      float omax = 0.f;
      for (int k = 0; k < n; ++k) {
        const float m2 = temphist[(k + n - 2) % n], m1 = temphist[(k + n - 1) % n], p1 = temphist[(k + 1) % n], p2 = temphist[(k + 2) % n];
        hist[k] = __fadd_rn(__fadd_rn(__fmul_rn(__fadd_rn(m2, p2), 1.f / 16.f), __fmul_rn(__fadd_rn(m1, p1), 4.f / 16.f)),
                            __fmul_rn(temphist[k], 6.f / 16.f));
        omax = k == 0 ? hist[0] : fmaxf(omax, hist[k]);
      }
      const float mag_thr = __fmul_rn(omax, 0.8f);
      for (int j = 0; j < n; ++j) {
        const int l = j > 0 ? j - 1 : n - 1, r2 = j < n - 1 ? j + 1 : 0;
        if (hist[j] > hist[l] && hist[j] > hist[r2] && hist[j] >= mag_thr) {
          float bin = __fadd_rn((float)j, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(hist[l], hist[r2])),
                                                    __fadd_rn(__fsub_rn(hist[l], __fmul_rn(2.f, hist[j])), hist[r2])));
          bin = bin < 0.f ? __fadd_rn((float)n, bin) : (bin >= (float)n ? __fsub_rn(bin, (float)n) : bin);
          float ang = __fsub_rn(360.f, __fmul_rn(10.f, bin));
          if (fabsf(__fsub_rn(ang, 360.f)) < 1.1920929e-07f) ang = 0.f;
          const int pos = atomicAdd(n_keys, 1);
          if (pos < key_cap) {
            KeyOut k;
            k.x = cd.x; k.y = cd.y; k.size = cd.size; k.angle = ang; k.response = cd.response;
            k.octave_field = cd.octave_field; k.o = cd.o; k.layer = cd.layer;
            keys[pos] = k;
          }
        }
      }
    }
    __syncwarp();
  }
}

// calcSIFTDescriptor, one thread per key point in OpenCV's loop order: the A/B reference of descriptor_kernel
// (-DIAM_SIFT_SEQ_DESC=1 selects it; bit-identical to the CPU restatement, about 4 x slower)
__global__ void descriptor_seq_kernel(const float* __restrict__ base, Pyramid P, const KeyOut* __restrict__ keys, int n_keys,
                                  uint8_t* __restrict__ des) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_keys) return;
  const KeyOut kp = keys[i];
  const Octave O = P.oct[kp.o];
  const float* img = base + O.gauss[kp.layer];
  const int w = O.w, h = O.h;
  constexpr int d = 4, n = 8;
  // published key point = (x, y, size) * 0.5 with octave field - 1; the descriptor works in the octave's own frame
  const float pub_scale = 0.5f;
  const int oct_pub = kp.o - 1;
  const float sc = oct_pub >= 0 ? __fdiv_rn(1.f, (float)(1 << oct_pub)) : (float)(1 << -oct_pub);
  const float ptx = __fmul_rn(__fmul_rn(kp.x, pub_scale), sc), pty = __fmul_rn(__fmul_rn(kp.y, pub_scale), sc);
  const float scl = __fmul_rn(__fmul_rn(__fmul_rn(kp.size, pub_scale), sc), 0.5f);
  float ori = __fsub_rn(360.f, kp.angle);
  if (fabsf(__fsub_rn(ori, 360.f)) < 1.1920929e-07f) ori = 0.f;
  const int px = __float2int_rn(ptx), py = __float2int_rn(pty);
  const float rad = __fmul_rn(ori, 0.017453292519943295f);
  float cos_t = (float)cos((double)rad), sin_t = (float)sin((double)rad);
  const float bins_per_rad = 8.f / 360.f;
  const float exp_scale = -1.f / (d * d * 0.5f);
  const float hist_width = __fmul_rn(3.f, scl);
  int radius = __float2int_rn(__fmul_rn(__fmul_rn(__fmul_rn(hist_width, 1.4142135623730951f), (float)(d + 1)), 0.5f));
  radius = min(radius, (int)sqrt((double)w * w + (double)h * h));
  cos_t = __fdiv_rn(cos_t, hist_width);
  sin_t = __fdiv_rn(sin_t, hist_width);
  float hist[(d + 2) * (d + 2) * (n + 2)];
  for (int k = 0; k < (d + 2) * (d + 2) * (n + 2); ++k) hist[k] = 0.f;
  for (int di = -radius; di <= radius; ++di)
    for (int dj = -radius; dj <= radius; ++dj) {
      const float c_rot = __fsub_rn(__fmul_rn((float)dj, cos_t), __fmul_rn((float)di, sin_t));
      const float r_rot = __fadd_rn(__fmul_rn((float)dj, sin_t), __fmul_rn((float)di, cos_t));
      float rbin = __fsub_rn(__fadd_rn(r_rot, (float)(d / 2)), 0.5f);
      float cbin = __fsub_rn(__fadd_rn(c_rot, (float)(d / 2)), 0.5f);
      const int r = py + di, c = px + dj;
      if (rbin > -1.f && rbin < (float)d && cbin > -1.f && cbin < (float)d && r > 0 && r < h - 1 && c > 0 && c < w - 1) {
        const float dx = __fsub_rn(at(img, w, r, c + 1), at(img, w, r, c - 1));
        const float dy = __fsub_rn(at(img, w, r - 1, c), at(img, w, r + 1, c));
        const float wgt = expf(__fmul_rn(__fadd_rn(__fmul_rn(c_rot, c_rot), __fmul_rn(r_rot, r_rot)), exp_scale));
        const float o_deg = fast_atan2_deg(dy, dx);
        const float mag0 = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        float obin = __fmul_rn(__fsub_rn(o_deg, ori), bins_per_rad);
        const float mag = __fmul_rn(mag0, wgt);
        const int r0 = (int)floorf(rbin), c0 = (int)floorf(cbin);
        int o0 = (int)floorf(obin);
        rbin = __fsub_rn(rbin, (float)r0);
        cbin = __fsub_rn(cbin, (float)c0);
        obin = __fsub_rn(obin, (float)o0);
        if (o0 < 0) o0 += n;
        if (o0 >= n) o0 -= n;
        const float v_r1 = __fmul_rn(mag, rbin), v_r0 = __fsub_rn(mag, v_r1);
        const float v_rc11 = __fmul_rn(v_r1, cbin), v_rc10 = __fsub_rn(v_r1, v_rc11);
        const float v_rc01 = __fmul_rn(v_r0, cbin), v_rc00 = __fsub_rn(v_r0, v_rc01);
        const float v111 = __fmul_rn(v_rc11, obin), v110 = __fsub_rn(v_rc11, v111);
        const float v101 = __fmul_rn(v_rc10, obin), v100 = __fsub_rn(v_rc10, v101);
        const float v011 = __fmul_rn(v_rc01, obin), v010 = __fsub_rn(v_rc01, v011);
        const float v001 = __fmul_rn(v_rc00, obin), v000 = __fsub_rn(v_rc00, v001);
        const int idx = ((r0 + 1) * (d + 2) + c0 + 1) * (n + 2) + o0;
        hist[idx] = __fadd_rn(hist[idx], v000);
        hist[idx + 1] = __fadd_rn(hist[idx + 1], v001);
        hist[idx + (n + 2)] = __fadd_rn(hist[idx + (n + 2)], v010);
        hist[idx + (n + 3)] = __fadd_rn(hist[idx + (n + 3)], v011);
        hist[idx + (d + 2) * (n + 2)] = __fadd_rn(hist[idx + (d + 2) * (n + 2)], v100);
        hist[idx + (d + 2) * (n + 2) + 1] = __fadd_rn(hist[idx + (d + 2) * (n + 2) + 1], v101);
        hist[idx + (d + 3) * (n + 2)] = __fadd_rn(hist[idx + (d + 3) * (n + 2)], v110);
        hist[idx + (d + 3) * (n + 2) + 1] = __fadd_rn(hist[idx + (d + 3) * (n + 2) + 1], v111);
      }
    }
  // the orientation histograms are circular; then normalise, clip at 0.2, normalise to 512 and saturate
  float nrm2 = 0.f;
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) {
      const int idx = ((a + 1) * (d + 2) + (b + 1)) * (n + 2);
      hist[idx] = __fadd_rn(hist[idx], hist[idx + n]);
      hist[idx + 1] = __fadd_rn(hist[idx + 1], hist[idx + n + 1]);
      for (int k = 0; k < n; ++k) nrm2 = __fadd_rn(nrm2, __fmul_rn(hist[idx + k], hist[idx + k]));
    }
  const float thr = __fmul_rn(__fsqrt_rn(nrm2), 0.2f);
  nrm2 = 0.f;
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) {
      const int idx = ((a + 1) * (d + 2) + (b + 1)) * (n + 2);
      for (int k = 0; k < n; ++k) {
        const float val = fminf(hist[idx + k], thr);
        hist[idx + k] = val;
        nrm2 = __fadd_rn(nrm2, __fmul_rn(val, val));
      }
    }
  const float s = __fdiv_rn(512.f, fmaxf(__fsqrt_rn(nrm2), 1.1920929e-07f));
  uint8_t* out = des + (size_t)i * 128;
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) {
      const int idx = ((a + 1) * (d + 2) + (b + 1)) * (n + 2);
      for (int k = 0; k < n; ++k) out[(a * d + b) * n + k] = (uint8_t)min(255, max(0, __float2int_rn(__fmul_rn(hist[idx + k], s))));
    }
}

// calcSIFTDescriptor, one WARP per key point.  The lanes share the (2 radius + 1)^2 window; the 8 tri-linear
// contributions of every sample are added to the warp's histogram in shared memory as fixed point (sixteenths in one
// 32-bit counter, the exact remainder in units of 2^-24 in a second one: native 32-bit shared atomics, no CAS loop):
// integer addition does not depend on the order the lanes arrive in, so the result is deterministic, and it is closer
// to the true sum than OpenCV's sequential float sum (which carries ~1e-6 of rounding noise -- why descriptors agree
// with cv2 to +-1 rather than bit for bit).  Lane 0 finishes with OpenCV's normalise / clip / quantise.
constexpr int kDescWarps = 4;
constexpr int kHistLen = 6 * 6 * 10;
__global__ void __launch_bounds__(kDescWarps * 32) descriptor_kernel(const float* __restrict__ base, Pyramid P,
                                                                     const KeyOut* __restrict__ keys, int n_keys,
                                                                     uint8_t* __restrict__ des) {
  __shared__ uint32_t hist_s[kDescWarps][2 * kHistLen];   // per bin: sixteenths, and the remainder in units of 2^-24
  __shared__ float fin_s[kDescWarps][kHistLen];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kDescWarps + warp;
  if (i >= n_keys) return;          // warp-uniform; no block-wide barrier below
  uint32_t* hq = hist_s[warp];
  for (int k = lane; k < 2 * kHistLen; k += 32) hq[k] = 0u;
  __syncwarp();
  const KeyOut kp = keys[i];
  const Octave O = P.oct[kp.o];
  const float* img = base + O.gauss[kp.layer];
  const int w = O.w, h = O.h;
  constexpr int d = 4, n = 8;
  const int oct_pub = kp.o - 1;
  const float sc = oct_pub >= 0 ? __fdiv_rn(1.f, (float)(1 << oct_pub)) : (float)(1 << -oct_pub);
  const float ptx = __fmul_rn(__fmul_rn(kp.x, 0.5f), sc), pty = __fmul_rn(__fmul_rn(kp.y, 0.5f), sc);
  const float scl = __fmul_rn(__fmul_rn(__fmul_rn(kp.size, 0.5f), sc), 0.5f);
  float ori = __fsub_rn(360.f, kp.angle);
  if (fabsf(__fsub_rn(ori, 360.f)) < 1.1920929e-07f) ori = 0.f;
  const int px = __float2int_rn(ptx), py = __float2int_rn(pty);
  const float rad = __fmul_rn(ori, 0.017453292519943295f);
  float cos_t = (float)cos((double)rad), sin_t = (float)sin((double)rad);
  const float bins_per_rad = 8.f / 360.f;
  const float exp_scale = -1.f / (d * d * 0.5f);
  const float hist_width = __fmul_rn(3.f, scl);
  int radius = __float2int_rn(__fmul_rn(__fmul_rn(__fmul_rn(hist_width, 1.4142135623730951f), (float)(d + 1)), 0.5f));
  radius = min(radius, (int)sqrt((double)w * w + (double)h * h));
  cos_t = __fdiv_rn(cos_t, hist_width);
  sin_t = __fdiv_rn(sin_t, hist_width);
  const int side = 2 * radius + 1;
  // rows of the window that lie inside the image (r > 0 && r < h - 1), then lanes stride over columns
  const int di_lo = max(-radius, 1 - py), di_hi = min(radius, h - 2 - py);
  const int dj_lo = max(-radius, 1 - px), dj_hi = min(radius, w - 2 - px);
  const int ncol = dj_hi - dj_lo + 1;
  (void)side;
  if (ncol > 0) {
    for (int di = di_lo; di <= di_hi; ++di) {
      const float* row = img + (size_t)(py + di) * w;
      for (int dj = dj_lo + lane; dj <= dj_hi; dj += 32) {
        const float c_rot = __fsub_rn(__fmul_rn((float)dj, cos_t), __fmul_rn((float)di, sin_t));
        const float r_rot = __fadd_rn(__fmul_rn((float)dj, sin_t), __fmul_rn((float)di, cos_t));
        float rbin = __fsub_rn(__fadd_rn(r_rot, (float)(d / 2)), 0.5f);
        float cbin = __fsub_rn(__fadd_rn(c_rot, (float)(d / 2)), 0.5f);
        if (!(rbin > -1.f && rbin < (float)d && cbin > -1.f && cbin < (float)d)) continue;
        const int c = px + dj;
        const float dx = __fsub_rn(row[c + 1], row[c - 1]);
        const float dy = __fsub_rn(row[c - w], row[c + w]);
        const float wgt = expf(__fmul_rn(__fadd_rn(__fmul_rn(c_rot, c_rot), __fmul_rn(r_rot, r_rot)), exp_scale));
        const float o_deg = fast_atan2_deg(dy, dx);
        const float mag0 = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        float obin = __fmul_rn(__fsub_rn(o_deg, ori), bins_per_rad);
        const float mag = __fmul_rn(mag0, wgt);
        const int r0 = (int)floorf(rbin), c0 = (int)floorf(cbin);
        int o0 = (int)floorf(obin);
        rbin = __fsub_rn(rbin, (float)r0);
        cbin = __fsub_rn(cbin, (float)c0);
        obin = __fsub_rn(obin, (float)o0);
        if (o0 < 0) o0 += n;
        if (o0 >= n) o0 -= n;
        const float v_r1 = __fmul_rn(mag, rbin), v_r0 = __fsub_rn(mag, v_r1);
        const float v_rc11 = __fmul_rn(v_r1, cbin), v_rc10 = __fsub_rn(v_r1, v_rc11);
        const float v_rc01 = __fmul_rn(v_r0, cbin), v_rc00 = __fsub_rn(v_r0, v_rc01);
        const float v111 = __fmul_rn(v_rc11, obin), v110 = __fsub_rn(v_rc11, v111);
        const float v101 = __fmul_rn(v_rc10, obin), v100 = __fsub_rn(v_rc10, v101);
        const float v011 = __fmul_rn(v_rc01, obin), v010 = __fsub_rn(v_rc01, v011);
        const float v001 = __fmul_rn(v_rc00, obin), v000 = __fsub_rn(v_rc00, v001);
        const int idx = ((r0 + 1) * (d + 2) + c0 + 1) * (n + 2) + o0;
        auto add = [&](int at, float v) {   // v = hi / 16 + rem exactly; both parts are native 32-bit shared atomics
          const uint32_t hi = __float2uint_rd(__fmul_rn(v, 16.f));
          const float rem = __fsub_rn(v, __fmul_rn((float)hi, 0.0625f));
          atomicAdd(&hq[2 * at], hi);
          atomicAdd(&hq[2 * at + 1], __float2uint_rn(__fmul_rn(rem, 16777216.f)));
        };
        add(idx, v000);
        add(idx + 1, v001);
        add(idx + (n + 2), v010);
        add(idx + (n + 3), v011);
        add(idx + (d + 2) * (n + 2), v100);
        add(idx + (d + 2) * (n + 2) + 1, v101);
        add(idx + (d + 3) * (n + 2), v110);
        add(idx + (d + 3) * (n + 2) + 1, v111);
      }
    }
  }
  __syncwarp();
  float* hist = fin_s[warp];
  for (int k = lane; k < kHistLen; k += 32) hist[k] = (float)((double)hq[2 * k] * 0.0625 + (double)hq[2 * k + 1] * 5.9604644775390625e-08);
  __syncwarp();
  if (lane == 0) {
    float nrm2 = 0.f;
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < d; ++b) {
        const int idx = ((a + 1) * (d + 2) + (b + 1)) * (n + 2);
        hist[idx] = __fadd_rn(hist[idx], hist[idx + n]);
        hist[idx + 1] = __fadd_rn(hist[idx + 1], hist[idx + n + 1]);
        for (int k = 0; k < n; ++k) nrm2 = __fadd_rn(nrm2, __fmul_rn(hist[idx + k], hist[idx + k]));
      }
    const float thr = __fmul_rn(__fsqrt_rn(nrm2), 0.2f);
    nrm2 = 0.f;
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < d; ++b) {
        const int idx = ((a + 1) * (d + 2) + (b + 1)) * (n + 2);
        for (int k = 0; k < n; ++k) {
          const float val = fminf(hist[idx + k], thr);
          hist[idx + k] = val;
          nrm2 = __fadd_rn(nrm2, __fmul_rn(val, val));
        }
      }
    hist[0] = __fdiv_rn(512.f, fmaxf(__fsqrt_rn(nrm2), 1.1920929e-07f));   // slot 0 is a border bin: free
  }
  __syncwarp();
  const float s = hist[0];
  uint8_t* out = des + (size_t)i * 128;
  for (int e = lane; e < 128; e += 32) {
    const int cell = e >> 3, k = e & 7;
    const int idx = (((cell >> 2) + 1) * (d + 2) + ((cell & 3) + 1)) * (n + 2) + k;
    out[e] = (uint8_t)min(255, max(0, __float2int_rn(__fmul_rn(hist[idx], s))));
  }
}

// descriptor rows (128 bytes = one warp of words) into their final order
__global__ void gather_rows_kernel(const uint32_t* __restrict__ src, const int* __restrict__ order, int n, uint32_t* __restrict__ dst) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row < n) dst[(size_t)row * 32 + lane] = src[(size_t)order[row] * 32 + lane];
}

BlurKernel make_kernel(double sigma) {
  BlurKernel K{};
  int ks = (int)std::nearbyint(sigma * 4 * 2 + 1) | 1;
  ks = std::min(ks, kMaxKernel - 1);
  K.ksize = ks;
  double sum = 0, v[kMaxKernel];
  for (int i = 0; i < ks; ++i) {
    const double x = i - (ks - 1) * 0.5;
    v[i] = std::exp(-(x * x) / (2.0 * sigma * sigma));
    sum += v[i];
  }
  for (int i = 0; i < ks; ++i) K.k[i] = (float)(v[i] / sum);
  return K;
}

}  // namespace

SiftScratch::~SiftScratch() {
  if (buf) cudaFree(buf);
  if (pinned) cudaFreeHost(pinned);
}

int sift_detect(const uint8_t* gray, int w, int h, int max_out, float* out_kp5, int32_t* out_octave, uint8_t* out_des,
                int* out_n, SiftScratch* scratch, cudaStream_t stream, std::string* err) {
  auto fail = [&](int code, const std::string& what) {
    if (err) *err = what;
    return code;
  };
  *out_n = 0;
  if (w < 2 || h < 2 || max_out < 0) return fail(-1, "bad image size");
  static const bool trace = std::getenv("IAM_SIFT_TRACE") != nullptr;   // stage times on stderr (debug aid)
  auto t_prev = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!trace) return;
    const auto t = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[sift] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
    t_prev = t;
  };
  const int bw = 2 * w, bh = 2 * h;
  Pyramid P{};
  P.n_oct = (int)std::nearbyint(std::log((double)std::min(bw, bh)) / std::log(2.0) - 2.0) + 1;
  P.n_oct = std::max(1, std::min(P.n_oct, kMaxOctaves));
  size_t off = 0;
  for (int o = 0; o < P.n_oct; ++o) {
    Octave& O = P.oct[o];
    O.w = o == 0 ? bw : P.oct[o - 1].w / 2;
    O.h = o == 0 ? bh : P.oct[o - 1].h / 2;
    if (O.w < 1 || O.h < 1) {
      P.n_oct = o;
      break;
    }
    const size_t px = ((size_t)O.w * O.h + 63) / 64 * 64;
    for (int i = 0; i < kNL + 3; ++i) {
      O.gauss[i] = off;
      off += px;
    }
    for (int i = 0; i < kNL + 2; ++i) {
      O.dog[i] = off;
      off += px;
    }
  }
  const size_t base_px = (size_t)bw * bh;
  const int cand_cap = (int)std::max<size_t>(65536, base_px / 16);
  const int key_cap = std::max(max_out, 1);
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t o_tmp = up(off * sizeof(float)), o_src = o_tmp + up(base_px * sizeof(float));
  const size_t o_cand = o_src + up((size_t)w * h), o_keys = o_cand + up((size_t)cand_cap * sizeof(Cand));
  const size_t o_des = o_keys + up((size_t)key_cap * sizeof(KeyOut)), o_des2 = o_des + up((size_t)key_cap * 128);
  const size_t o_order = o_des2 + up((size_t)key_cap * 128);
  const size_t o_cnt = o_order + up((size_t)key_cap * sizeof(int));
  const size_t total = o_cnt + 256;
  cudaError_t e;
#define SC(call) \
  if ((e = (call)) != cudaSuccess) return fail(-2, std::string(#call) + ": " + cudaGetErrorString(e));
  SiftScratch local;
  SiftScratch* sc = scratch ? scratch : &local;
  if (sc->cap < total) {
    SC(cudaStreamSynchronize(stream));
    if (sc->buf) cudaFree(sc->buf);
    sc->buf = nullptr;
    sc->cap = 0;
    if ((e = cudaMalloc(&sc->buf, total)) != cudaSuccess) return fail(-3, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    sc->cap = total;
  }
  uint8_t* blk = static_cast<uint8_t*>(sc->buf);
  float* base = reinterpret_cast<float*>(blk);
  float* tmp = reinterpret_cast<float*>(blk + o_tmp);
  uint8_t* d_src = blk + o_src;
  Cand* d_cand = reinterpret_cast<Cand*>(blk + o_cand);
  KeyOut* d_keys = reinterpret_cast<KeyOut*>(blk + o_keys);
  uint8_t* d_des = blk + o_des;
  uint8_t* d_des2 = blk + o_des2;
  int* d_cnt = reinterpret_cast<int*>(blk + o_cnt);
  SC(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(int), stream));
  SC(cudaMemcpyAsync(d_src, gray, (size_t)w * h, cudaMemcpyHostToDevice, stream));
  const dim3 blk2(32, 8);
  auto grid = [&](int W, int H) { return dim3((W + 31) / 32, (H + 7) / 8); };
  int launches = 0;
  auto blur = [&](const float* src, float* dst, int W, int H, double sigma, float* dog) {
    const BlurKernel K = make_kernel(sigma);
    launches += 2;
    if (W >= 64 && H >= 64 && !IAM_SIFT_GENERIC_BLUR) {
      switch (K.ksize >> 1) {
        case 5: return launch_blur_t<5>(src, tmp, dst, W, H, K, dog, stream);
        case 6: return launch_blur_t<6>(src, tmp, dst, W, H, K, dog, stream);
        case 8: return launch_blur_t<8>(src, tmp, dst, W, H, K, dog, stream);
        case 10: return launch_blur_t<10>(src, tmp, dst, W, H, K, dog, stream);
        case 13: return launch_blur_t<13>(src, tmp, dst, W, H, K, dog, stream);
        default: break;
      }
    }
    blur_rows_kernel<<<grid(W, H), blk2, 0, stream>>>(src, W, H, K, tmp);
    blur_cols_kernel<<<grid(W, H), blk2, 0, stream>>>(tmp, W, H, K, dst, src, dog);
  };
  // base image: doubled, blurred from the assumed 2 * 0.5 to sigma 1.6
  const double sigma0 = 1.6;
  float* dbl = base + P.oct[0].dog[kNL + 1];   // scratch: overwritten by the last DoG layer later
  up2_kernel<<<grid(bw, bh), blk2, 0, stream>>>(d_src, w, h, dbl);
  ++launches;
  blur(dbl, base + P.oct[0].gauss[0], bw, bh, std::sqrt(std::max(sigma0 * sigma0 - 0.5 * 0.5 * 4, 0.01)), nullptr);
  double sig[kNL + 3];
  sig[0] = sigma0;
  const double k = std::pow(2.0, 1.0 / kNL);
  for (int i = 1; i < kNL + 3; ++i) {
    const double sp = std::pow(k, (double)(i - 1)) * sigma0, st = sp * k;
    sig[i] = std::sqrt(st * st - sp * sp);
  }
  for (int o = 0; o < P.n_oct; ++o) {
    const Octave& O = P.oct[o];
    for (int i = 0; i < kNL + 3; ++i) {
      float* dst = base + O.gauss[i];
      if (o == 0 && i == 0) continue;
      if (i == 0) {
        half_kernel<<<grid(O.w, O.h), blk2, 0, stream>>>(base + P.oct[o - 1].gauss[kNL], P.oct[o - 1].w, dst, O.w, O.h);
        ++launches;
      } else {
        blur(base + O.gauss[i - 1], dst, O.w, O.h, sig[i], base + O.dog[i - 1]);   // DoG layer i-1 = layer i - layer i-1
      }
    }
  }
  for (int o = 0; o < P.n_oct; ++o) {
    const Octave& O = P.oct[o];
    if (O.w <= 2 * kImgBorder || O.h <= 2 * kImgBorder) continue;
    dim3 g3 = grid(O.w - 2 * kImgBorder, O.h - 2 * kImgBorder);
    g3.z = kNL;
    extrema_kernel<<<g3, blk2, 0, stream>>>(base, P, o, d_cand, d_cnt, cand_cap);
    ++launches;
  }
#if IAM_SIFT_SEQ_ORI
  orientation_seq_kernel<<<(cand_cap + 127) / 128, 128, 0, stream>>>(base, P, d_cand, d_cnt, cand_cap, d_keys, d_cnt + 1, key_cap);
#else
  orientation_kernel<<<148 * 8, kOriWarps * 32, 0, stream>>>(base, P, d_cand, d_cnt, cand_cap, d_keys, d_cnt + 1, key_cap);
#endif
  ++launches;
  mark("enqueue pyramid..orientation");
  int h_cnt[2];
  SC(cudaMemcpyAsync(h_cnt, d_cnt, sizeof h_cnt, cudaMemcpyDeviceToHost, stream));
  SC(cudaStreamSynchronize(stream));
  mark("wait for the key point count");
  if (h_cnt[0] > cand_cap) return fail(-5, "more scale-space extrema than the candidate buffer holds");
  if (h_cnt[1] > key_cap) return fail(-5, "more key points than the output buffers hold");
  const int nk = h_cnt[1];
  if (nk > 0) {
    // page-locked staging owned by the scratch block: key points down, final order up
    const size_t pin_need = (size_t)nk * (sizeof(KeyOut) + sizeof(int));
    if (sc->pinned_cap < pin_need) {
      if (sc->pinned) cudaFreeHost(sc->pinned);
      sc->pinned = nullptr;
      sc->pinned_cap = 0;
      const size_t want = pin_need + pin_need / 2;
      if ((e = cudaHostAlloc(&sc->pinned, want, cudaHostAllocDefault)) != cudaSuccess)
        return fail(-3, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
      sc->pinned_cap = want;
    }
    KeyOut* hk = static_cast<KeyOut*>(sc->pinned);
    int* h_order = reinterpret_cast<int*>(hk + nk);
    SC(cudaMemcpyAsync(hk, d_keys, (size_t)nk * sizeof(KeyOut), cudaMemcpyDeviceToHost, stream));
    SC(cudaStreamSynchronize(stream));
    mark("download key points");
    // descriptors are computed in device order while the host sorts
#if IAM_SIFT_SEQ_DESC
    descriptor_seq_kernel<<<(nk + 63) / 64, 64, 0, stream>>>(base, P, d_keys, nk, d_des);
#else
    descriptor_kernel<<<(nk + kDescWarps - 1) / kDescWarps, kDescWarps * 32, 0, stream>>>(base, P, d_keys, nk, d_des);
#endif
    ++launches;
    // KeyPointsFilter::removeDuplicatedSorted: order by (x, y, size desc, angle, response desc, octave desc), drop
    // key points that repeat (x, y, size, angle); then first octave -1: halve coordinates and size.
    // x and y are positive floats, so their bit patterns order like the values: one 64-bit key decides all but ties.
    struct SortKey {
      uint64_t xy;
      int idx;
    };
    std::vector<SortKey> sk(nk);
    for (int i = 0; i < nk; ++i) {
      uint32_t bx, by;
      std::memcpy(&bx, &hk[i].x, 4);
      std::memcpy(&by, &hk[i].y, 4);
      sk[i] = {(uint64_t)bx << 32 | by, i};
    }
    {
      // LSD radix sort on the 64-bit key (four 16-bit digits, stable), then the few runs of equal (x, y) -- the extra
      // orientations of one extremum -- by the remaining fields
      std::vector<SortKey> tmp_sk(nk);
      std::vector<uint32_t> hist(65536);
      SortKey *src = sk.data(), *dst = tmp_sk.data();
      for (int pass = 0; pass < 4; ++pass) {
        const int shift = 16 * pass;
        std::fill(hist.begin(), hist.end(), 0u);
        for (int i = 0; i < nk; ++i) ++hist[(src[i].xy >> shift) & 0xFFFF];
        uint32_t run = 0;
        for (auto& hcount : hist) {
          const uint32_t cnt = hcount;
          hcount = run;
          run += cnt;
        }
        for (int i = 0; i < nk; ++i) dst[hist[(src[i].xy >> shift) & 0xFFFF]++] = src[i];
        std::swap(src, dst);
      }                                     // four passes: the sorted list is back in sk
      auto tie_less = [&](const SortKey& a, const SortKey& b) {
        const KeyOut &p = hk[a.idx], &q = hk[b.idx];
        if (p.size != q.size) return p.size > q.size;
        if (p.angle != q.angle) return p.angle < q.angle;
        if (p.response != q.response) return p.response > q.response;
        if (p.octave_field != q.octave_field) return p.octave_field > q.octave_field;
        return a.idx < b.idx;
      };
      for (int i = 0; i < nk;) {
        int j = i + 1;
        while (j < nk && sk[j].xy == sk[i].xy) ++j;
        if (j - i > 1) std::sort(sk.begin() + i, sk.begin() + j, tie_less);
        i = j;
      }
    }
    int n_out = 0;
    const KeyOut* last = nullptr;
    for (int i = 0; i < nk; ++i) {
      const KeyOut& kpt = hk[sk[i].idx];
      if (last && last->x == kpt.x && last->y == kpt.y && last->size == kpt.size && last->angle == kpt.angle) continue;
      last = &kpt;
      float* o5 = out_kp5 + (size_t)n_out * 5;
      o5[0] = kpt.x * 0.5f;
      o5[1] = kpt.y * 0.5f;
      o5[2] = kpt.size * 0.5f;
      o5[3] = kpt.angle;
      o5[4] = kpt.response;
      out_octave[n_out] = (kpt.octave_field & ~255) | ((kpt.octave_field - 1) & 255);
      h_order[n_out++] = sk[i].idx;
    }
    mark("host sort + key point output");
    // descriptors leave the device already in the final order, straight into the caller's array
    int* d_order = reinterpret_cast<int*>(blk + o_order);
    SC(cudaMemcpyAsync(d_order, h_order, (size_t)n_out * sizeof(int), cudaMemcpyHostToDevice, stream));
    gather_rows_kernel<<<(n_out + 3) / 4, 128, 0, stream>>>(reinterpret_cast<const uint32_t*>(d_des), d_order, n_out,
                                                           reinterpret_cast<uint32_t*>(d_des2));
    ++launches;
    SC(cudaMemcpyAsync(out_des, d_des2, (size_t)n_out * 128, cudaMemcpyDeviceToHost, stream));
    SC(cudaStreamSynchronize(stream));
    *out_n = n_out;
    mark("descriptors in order to host");
  }
#undef SC
  return launches;
}

}  // namespace iam
