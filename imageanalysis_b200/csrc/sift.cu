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
#include <cmath>
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
__global__ void blur_cols_kernel(const float* __restrict__ src, int w, int h, BlurKernel K, float* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const int r = K.ksize >> 1;
  float s = 0.f;
  if (y >= r && y < h - r) {
    for (int i = 0; i < K.ksize; ++i) s = __fadd_rn(s, __fmul_rn(K.k[i], src[(size_t)(y - r + i) * w + x]));
  } else {
    for (int i = 0; i < K.ksize; ++i) s = __fadd_rn(s, __fmul_rn(K.k[i], src[(size_t)reflect101(y - r + i, h) * w + x]));
  }
  dst[(size_t)y * w + x] = s;
}
__global__ void half_kernel(const float* __restrict__ src, int sw, float* __restrict__ dst, int w, int h) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x < w && y < h) dst[(size_t)y * w + x] = src[(size_t)(2 * y) * sw + 2 * x];
}
__global__ void sub_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, float* __restrict__ d) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = __fsub_rn(a[i], b[i]);
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

__global__ void extrema_kernel(const float* __restrict__ base, Pyramid P, int o, int layer, Cand* __restrict__ cand,
                               int* __restrict__ n_cand, int cap) {
  const Octave O = P.oct[o];
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

// calcOrientationHist + the peak loop of findScaleSpaceExtrema: one thread per refined extremum
__global__ void orientation_kernel(const float* __restrict__ base, Pyramid P, const Cand* __restrict__ cand,
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

// calcSIFTDescriptor: one thread per key point, OpenCV's loop order
__global__ void descriptor_kernel(const float* __restrict__ base, Pyramid P, const KeyOut* __restrict__ keys, int n_keys,
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
}

int sift_detect(const uint8_t* gray, int w, int h, int max_out, float* out_kp5, int32_t* out_octave, uint8_t* out_des,
                int* out_n, SiftScratch* scratch, cudaStream_t stream, std::string* err) {
  auto fail = [&](int code, const std::string& what) {
    if (err) *err = what;
    return code;
  };
  *out_n = 0;
  if (w < 2 || h < 2 || max_out < 0) return fail(-1, "bad image size");
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
  const size_t o_des = o_keys + up((size_t)key_cap * sizeof(KeyOut)), o_cnt = o_des + up((size_t)key_cap * 128);
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
  int* d_cnt = reinterpret_cast<int*>(blk + o_cnt);
  SC(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(int), stream));
  SC(cudaMemcpyAsync(d_src, gray, (size_t)w * h, cudaMemcpyHostToDevice, stream));
  const dim3 blk2(32, 8);
  auto grid = [&](int W, int H) { return dim3((W + 31) / 32, (H + 7) / 8); };
  int launches = 0;
  auto blur = [&](const float* src, float* dst, int W, int H, double sigma) {
    const BlurKernel K = make_kernel(sigma);
    blur_rows_kernel<<<grid(W, H), blk2, 0, stream>>>(src, W, H, K, tmp);
    blur_cols_kernel<<<grid(W, H), blk2, 0, stream>>>(tmp, W, H, K, dst);
    launches += 2;
  };
  // base image: doubled, blurred from the assumed 2 * 0.5 to sigma 1.6
  const double sigma0 = 1.6;
  float* dbl = base + P.oct[0].dog[0];   // scratch: overwritten by the DoG later
  up2_kernel<<<grid(bw, bh), blk2, 0, stream>>>(d_src, w, h, dbl);
  ++launches;
  blur(dbl, base + P.oct[0].gauss[0], bw, bh, std::sqrt(std::max(sigma0 * sigma0 - 0.5 * 0.5 * 4, 0.01)));
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
        blur(base + O.gauss[i - 1], dst, O.w, O.h, sig[i]);
      }
    }
    const size_t npx = (size_t)O.w * O.h;
    for (int i = 0; i < kNL + 2; ++i) {
      sub_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, stream>>>(base + O.gauss[i + 1], base + O.gauss[i], npx, base + O.dog[i]);
      ++launches;
    }
  }
  for (int o = 0; o < P.n_oct; ++o) {
    const Octave& O = P.oct[o];
    if (O.w <= 2 * kImgBorder || O.h <= 2 * kImgBorder) continue;
    for (int i = 1; i <= kNL; ++i) {
      extrema_kernel<<<grid(O.w - 2 * kImgBorder, O.h - 2 * kImgBorder), blk2, 0, stream>>>(base, P, o, i, d_cand, d_cnt, cand_cap);
      ++launches;
    }
  }
  orientation_kernel<<<(cand_cap + 127) / 128, 128, 0, stream>>>(base, P, d_cand, d_cnt, cand_cap, d_keys, d_cnt + 1, key_cap);
  ++launches;
  int h_cnt[2];
  SC(cudaMemcpyAsync(h_cnt, d_cnt, sizeof h_cnt, cudaMemcpyDeviceToHost, stream));
  SC(cudaStreamSynchronize(stream));
  if (h_cnt[0] > cand_cap) return fail(-5, "more scale-space extrema than the candidate buffer holds");
  if (h_cnt[1] > key_cap) return fail(-5, "more key points than the output buffers hold");
  const int nk = h_cnt[1];
  if (nk > 0) {
    descriptor_kernel<<<(nk + 63) / 64, 64, 0, stream>>>(base, P, d_keys, nk, d_des);
    ++launches;
    std::vector<KeyOut> hk(nk);
    std::vector<uint8_t> hd((size_t)nk * 128);
    SC(cudaMemcpyAsync(hk.data(), d_keys, (size_t)nk * sizeof(KeyOut), cudaMemcpyDeviceToHost, stream));
    SC(cudaMemcpyAsync(hd.data(), d_des, (size_t)nk * 128, cudaMemcpyDeviceToHost, stream));
    SC(cudaStreamSynchronize(stream));
    // KeyPointsFilter::removeDuplicatedSorted: order by (x, y, size desc, angle, response desc, octave desc), drop
    // key points that repeat (x, y, size, angle); then first octave -1: halve coordinates and size
    std::vector<int> order(nk);
    for (int i = 0; i < nk; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
      const KeyOut &p = hk[a], &q = hk[b];
      if (p.x != q.x) return p.x < q.x;
      if (p.y != q.y) return p.y < q.y;
      if (p.size != q.size) return p.size > q.size;
      if (p.angle != q.angle) return p.angle < q.angle;
      if (p.response != q.response) return p.response > q.response;
      if (p.octave_field != q.octave_field) return p.octave_field > q.octave_field;
      return a < b;
    });
    int n_out = 0;
    const KeyOut* last = nullptr;
    for (int i = 0; i < nk; ++i) {
      const KeyOut& kpt = hk[order[i]];
      if (last && last->x == kpt.x && last->y == kpt.y && last->size == kpt.size && last->angle == kpt.angle) continue;
      last = &kpt;
      float* o5 = out_kp5 + (size_t)n_out * 5;
      o5[0] = kpt.x * 0.5f;
      o5[1] = kpt.y * 0.5f;
      o5[2] = kpt.size * 0.5f;
      o5[3] = kpt.angle;
      o5[4] = kpt.response;
      out_octave[n_out] = (kpt.octave_field & ~255) | ((kpt.octave_field - 1) & 255);
      std::copy(hd.begin() + (size_t)order[i] * 128, hd.begin() + (size_t)order[i] * 128 + 128, out_des + (size_t)n_out * 128);
      ++n_out;
    }
    *out_n = n_out;
  }
#undef SC
  return launches;
}

}  // namespace iam
