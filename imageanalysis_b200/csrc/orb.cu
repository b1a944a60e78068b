// orb.cu — ORB detect + describe on the GPU, standing in for
//   detector = cv2.ORB_create(max_features); detector.detectAndCompute(scaled, None)
// (reference scripts/lib/image.py:243-245, :324).  OpenCV's defaults: 8 pyramid levels, scale 1.2, edge threshold 31,
// patch 31, FAST threshold 20, Harris score, WTA_K 2.  Stage by stage the arithmetic is OpenCV's (the CPU checker of
// the test suite restates it and is pinned against live cv2):
//   pyramid      each level from the previous one, bilinear with 8.8 fixed-point weights (INTER_LINEAR_EXACT), plus a
//                32-pixel reflect-101 border so that every later stage reads without bounds tests
//   FAST-9/16    score = largest threshold at which the pixel is still a corner; 3x3 non-maximum suppression
//   retainBest   histogram / radix select of the n-th largest response; ties survive, as in KeyPointsFilter::retainBest
//   Harris       7x7 block of Sobel products -> float32 response, evaluated in OpenCV's order with no FMA contraction
//   orientation  intensity centroid over the circular patch (radius 15), OpenCV's polynomial fastAtan2
//   descriptor   256 steered BRIEF tests on the level blurred by the separable float 7-tap Gaussian (sigma 2)
// All images of the call stay in HBM; only the grey input goes up and the key points / descriptors come back.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "orb.h"

namespace iam {
namespace {

constexpr int kLevels = 8;
constexpr int kEdge = 31;
constexpr int kBorder = 32;       // max(edgeThreshold, ceil(15 sqrt 2), 4) + 1, orb.cpp detectAndCompute
constexpr int kHalfPatch = 15;
constexpr int kFastThr = 20;

struct Level {
  int w, h;           // level size
  int pitch;          // padded row pitch = w + 2 * kBorder
  size_t img_off;     // offset of the padded image in the pyramid block (bytes)
  size_t blur_off;    // ... of the blurred padded image
  size_t score_off;   // u8 FAST score map [h][w]
  int cand_off;       // first candidate slot of this level
  int cand_cap;
  int n_want;         // nfeaturesPerLevel
  float scale;
};

struct Params {
  Level lv[kLevels];
};

__constant__ int c_umax[kHalfPatch + 2];
__constant__ signed char c_pattern[256 * 4];
__constant__ float c_gauss[7];

__device__ __forceinline__ uint8_t* lvl_px(uint8_t* base, const Level& L, int x, int y) {
  return base + (size_t)(y + kBorder) * L.pitch + (x + kBorder);
}

// ---- pyramid -------------------------------------------------------------------------------------------------
__global__ void resize_kernel(const uint8_t* __restrict__ src_img, Level S, uint8_t* __restrict__ dst_img, Level D) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y * blockDim.y + threadIdx.y;
  if (dx >= D.w || dy >= D.h) return;
  auto coef = [](int d, int dn, int sn, int& i, int& a) {
    const double scale = (double)sn / (double)dn;
    const double f = ((double)d + 0.5) * scale - 0.5;
    int ii = (int)floor(f);
    double aa = f - (double)ii;
    if (ii < 0) {
      ii = 0;
      aa = 0.0;
    }
    if (ii >= sn - 1) {
      ii = sn - 1;
      aa = 0.0;
    }
    i = ii;
    a = (int)rint(aa * 256.0);
  };
  int xi, xa, yi, ya;
  coef(dx, D.w, S.w, xi, xa);
  coef(dy, D.h, S.h, yi, ya);
  const int x1 = min(xi + 1, S.w - 1), y1 = min(yi + 1, S.h - 1);
  const uint8_t* r0 = src_img + (size_t)(yi + kBorder) * S.pitch + kBorder;
  const uint8_t* r1 = src_img + (size_t)(y1 + kBorder) * S.pitch + kBorder;
  const int h0 = r0[xi] * (256 - xa) + r0[x1] * xa;
  const int h1 = r1[xi] * (256 - xa) + r1[x1] * xa;
  const int v = h0 * (256 - ya) + h1 * ya;
  dst_img[(size_t)(dy + kBorder) * D.pitch + dx + kBorder] = (uint8_t)((v + (1 << 15)) >> 16);
}

__global__ void upload_kernel(const uint8_t* __restrict__ src, int w, int h, uint8_t* __restrict__ dst, int pitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x < w && y < h) dst[(size_t)(y + kBorder) * pitch + x + kBorder] = src[(size_t)y * w + x];
}

// reflect-101 border of width kBorder around the w x h interior (also used for the float blur buffer's source)
__global__ void border_kernel(uint8_t* __restrict__ img, int w, int h, int pitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  const int W = w + 2 * kBorder, H = h + 2 * kBorder;
  if (x >= W || y >= H) return;
  const int ix = x - kBorder, iy = y - kBorder;
  if (ix >= 0 && ix < w && iy >= 0 && iy < h) return;
  auto refl = [](int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
  };
  img[(size_t)y * pitch + x] = img[(size_t)(refl(iy, h) + kBorder) * pitch + refl(ix, w) + kBorder];
}

// ---- FAST-9/16 -----------------------------------------------------------------------------------------------
__global__ void fast_score_kernel(const uint8_t* __restrict__ img, Level L, uint8_t* __restrict__ score) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= L.w || y >= L.h) return;
  int out = 0;
  if (x >= 3 && y >= 3 && x < L.w - 3 && y < L.h - 3) {
    constexpr int cx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    constexpr int cy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    const uint8_t* p = img + (size_t)(y + kBorder) * L.pitch + x + kBorder;
    const int v = p[0];
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = v - (int)p[cy[k] * L.pitch + cx[k]];
    // Is the pixel a corner at threshold T?  Bit k of `b` / `dk`: ring pixel k darker / brighter than the centre by
    // more than T; nine contiguous set bits on the 16-ring (the mask doubled to 32 bits makes the ring linear).
    auto corner = [&](int T) {
      unsigned b = 0, dk = 0;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        b |= (unsigned)(d[k] > T) << k;
        dk |= (unsigned)(d[k] < -T) << k;
      }
      b |= b << 16;
      dk |= dk << 16;
      unsigned m = b & (b >> 1);
      m &= m >> 2;
      m &= m >> 4;
      m &= b >> 8;
      unsigned n = dk & (dk >> 1);
      n &= n >> 2;
      n &= n >> 4;
      n &= dk >> 8;
      return ((m | n) & 0xffffu) != 0u;
    };
    // cheap rejection: a 9-arc always contains one of every pair of opposite pixels (0, 8) and (4, 12)
    const bool may = ((d[0] > kFastThr || d[8] > kFastThr) && (d[4] > kFastThr || d[12] > kFastThr)) ||
                     ((d[0] < -kFastThr || d[8] < -kFastThr) && (d[4] < -kFastThr || d[12] < -kFastThr));
    if (may && corner(kFastThr)) {
      // score = the largest threshold at which it still is one (fast.cpp cornerScore): bisection on T in [20, 254]
      int lo = kFastThr, hi = 255;          // corner(lo) holds, corner(hi) cannot (|d| <= 255)
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (corner(mid)) lo = mid; else hi = mid;
      }
      out = lo;
    }
  }
  score[(size_t)y * L.w + x] = (uint8_t)out;
}

struct Cand {
  unsigned short x, y;
  float resp;        // FAST score, later the Harris response
};

__global__ void nms_collect_kernel(const uint8_t* __restrict__ score, Level L, int level, Cand* __restrict__ cand,
                                   int* __restrict__ cand_count, int* __restrict__ hist) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x < kEdge || y < kEdge || x >= L.w - kEdge || y >= L.h - kEdge) return;   // KeyPointsFilter::runByImageBorder
  const uint8_t* s = score + (size_t)y * L.w + x;
  const int c = s[0];
  if (c == 0) return;
  const int w = L.w;
  if (c > s[-1] && c > s[1] && c > s[-w - 1] && c > s[-w] && c > s[-w + 1] && c > s[w - 1] && c > s[w] && c > s[w + 1]) {
    const int pos = atomicAdd(&cand_count[level], 1);
    if (pos < L.cand_cap) cand[L.cand_off + pos] = Cand{(unsigned short)x, (unsigned short)y, (float)c};
    atomicAdd(&hist[level * 256 + c], 1);
  }
}

// retainBest(2 n) on the FAST scores: the smallest score that is still among the 2 n best (ties survive)
__global__ void fast_threshold_kernel(const int* __restrict__ hist, const int* __restrict__ cand_count, Params P,
                                      int* __restrict__ fast_thr) {
  const int l = threadIdx.x;
  if (l >= kLevels) return;
  const int total = min(cand_count[l], P.lv[l].cand_cap), want = 2 * P.lv[l].n_want;
  int thr = 0;
  if (want <= 0) thr = 1 << 20;
  else if (total > want) {
    int cum = 0;
    for (int s = 255; s >= 0; --s) {
      cum += hist[l * 256 + s];
      if (cum >= want) {
        thr = s;
        break;
      }
    }
  }
  fast_thr[l] = thr;
}

// ---- Harris response (orb.cpp HarrisResponses, blockSize 7) ---------------------------------------------------
__global__ void harris_kernel(const uint8_t* __restrict__ pyr, Params P, const int* __restrict__ cand_count,
                              const int* __restrict__ fast_thr, Cand* __restrict__ cand) {
  const int l = blockIdx.y;
  const Level L = P.lv[l];
  const int n = min(cand_count[l], L.cand_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
  Cand c = cand[L.cand_off + i];
  if (c.resp < (float)fast_thr[l]) {
    cand[L.cand_off + i].resp = -INFINITY;   // dropped by the first retainBest
    continue;
  }
  const uint8_t* img = pyr + L.img_off;
  const int st = L.pitch;
  int a = 0, b = 0, cc = 0;
  for (int dy = -3; dy <= 3; ++dy) {
    const uint8_t* p = img + (size_t)(c.y + dy + kBorder) * st + c.x - 3 + kBorder;
    for (int dx = 0; dx < 7; ++dx, ++p) {
      const int Ix = ((int)p[1] - (int)p[-1]) * 2 + ((int)p[-st + 1] - (int)p[-st - 1]) + ((int)p[st + 1] - (int)p[st - 1]);
      const int Iy = ((int)p[st] - (int)p[-st]) * 2 + ((int)p[st - 1] - (int)p[-st - 1]) + ((int)p[st + 1] - (int)p[-st + 1]);
      a += Ix * Ix;
      b += Iy * Iy;
      cc += Ix * Iy;
    }
  }
  const float scale = __fdiv_rn(1.0f, __fmul_rn(28.0f, 255.0f));      // 1 / ((1 << 2) * blockSize * 255)
  const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
  const float fa = (float)a, fb = (float)b, fc = (float)cc;
  const float ab = __fadd_rn(fa, fb);
  const float r = __fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(0.04f, ab), ab));
  cand[L.cand_off + i].resp = __fmul_rn(r, s4);
  }
}

__global__ void kp_base_kernel(Params P, const int* __restrict__ out_count, int* __restrict__ kp_base) {
  int total = 0;
  for (int l = 0; l < kLevels; ++l) {
    kp_base[l] = total;
    total += min(out_count[l], P.lv[l].cand_cap);
  }
  kp_base[kLevels] = total;
}

__device__ __forceinline__ uint32_t order_key(float f) {   // larger float -> larger key
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// retainBest(n) on the Harris responses: radix select of the n-th largest key, one CTA per level; then the survivors
// (key >= that key: ties survive) are appended to the level's output list
__global__ void __launch_bounds__(1024)
harris_select_kernel(Params P, const int* __restrict__ cand_count, const Cand* __restrict__ cand, int* __restrict__ out_count,
                     Cand* __restrict__ out) {
  __shared__ int s_hist[256];
  __shared__ uint32_t s_prefix, s_mask;
  __shared__ int s_want, s_alive;
  const int l = blockIdx.x;
  const Level L = P.lv[l];
  const int n = min(cand_count[l], L.cand_cap);
  const Cand* c = cand + L.cand_off;
  // candidates that survived the FAST-score cut
  if (threadIdx.x == 0) s_alive = 0;
  __syncthreads();
  int mine = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mine += c[i].resp != -INFINITY;
  atomicAdd(&s_alive, mine);
  __syncthreads();
  const int alive = s_alive, want = L.n_want;
  uint32_t kth = 0;          // keep everything with key >= kth
  if (want <= 0) {
    kth = 0xffffffffu;
  } else if (alive > want) {
    if (threadIdx.x == 0) {
      s_prefix = 0;
      s_mask = 0;
      s_want = want;
    }
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix, mask = s_mask;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (c[i].resp == -INFINITY) continue;
        const uint32_t k = order_key(c[i].resp);
        if ((k & mask) == prefix) atomicAdd(&s_hist[(k >> shift) & 255], 1);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int cum = 0, w = s_want;
        for (int bkt = 255; bkt >= 0; --bkt) {
          if (cum + s_hist[bkt] >= w) {
            s_prefix = prefix | ((uint32_t)bkt << shift);
            s_mask = mask | (0xffu << shift);
            s_want = w - cum;
            break;
          }
          cum += s_hist[bkt];
        }
      }
      __syncthreads();
    }
    kth = s_prefix;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (c[i].resp == -INFINITY) continue;
    if (order_key(c[i].resp) >= kth) {
      const int pos = atomicAdd(&out_count[l], 1);
      if (pos < L.cand_cap) out[L.cand_off + pos] = c[i];
    }
  }
}

// ---- orientation (orb.cpp ICAngles) + fastAtan2 ----------------------------------------------------------------
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

struct KpOut {
  float x, y, size, angle, response;
  int octave;
};

// one warp per key point: orientation, then the 32 descriptor bytes (lane = byte)
__global__ void __launch_bounds__(256)
describe_kernel(const uint8_t* __restrict__ pyr, Params P, const int* __restrict__ out_count, const int* __restrict__ kp_base,
                const Cand* __restrict__ kps, KpOut* __restrict__ out_kp, uint8_t* __restrict__ out_des, int max_out) {
  const int l = blockIdx.y;
  const Level L = P.lv[l];
  const int n = min(out_count[l], L.cand_cap);
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); i < n; i += gridDim.x * (blockDim.x / 32)) {
  if (kp_base[l] + i >= max_out) return;
  const Cand c = kps[L.cand_off + i];
  const int st = L.pitch;
  const uint8_t* center = pyr + L.img_off + (size_t)(c.y + kBorder) * st + c.x + kBorder;
  // moments: lane v handles rows +-v (v = 0..15)
  int m10 = 0, m01 = 0;
  if (lane == 0) {
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * center[u];
  } else if (lane <= kHalfPatch) {
    const int v = lane, d = c_umax[v];
    int vsum = 0;
    for (int u = -d; u <= d; ++u) {
      const int plus = center[u + v * st], minus = center[u - v * st];
      vsum += plus - minus;
      m10 += u * (plus + minus);
    }
    m01 = v * vsum;
  }
  for (int o = 16; o > 0; o >>= 1) {
    m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    m01 += __shfl_xor_sync(0xffffffffu, m01, o);
  }
  const float angle = fast_atan2_deg((float)m01, (float)m10);
  // steered BRIEF on the blurred level
  const uint8_t* bc = pyr + L.blur_off + (size_t)(c.y + kBorder) * st + c.x + kBorder;
  const float rad = __fmul_rn(angle, 0.017453292519943295f);
  const float a = (float)cos((double)rad), b = (float)sin((double)rad);
  unsigned byte = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const signed char* pt = c_pattern + (lane * 8 + k) * 4;
    const float x0 = pt[0], y0 = pt[1], x1 = pt[2], y1 = pt[3];
    const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
    const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
    const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
    const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
    const int t0 = bc[iy0 * st + ix0], t1 = bc[iy1 * st + ix1];
    byte |= (unsigned)(t0 < t1) << k;
  }
  const int o = kp_base[l] + i;
  out_des[(size_t)o * 32 + lane] = (uint8_t)byte;
  if (lane == 0) {
    KpOut k;
    k.x = __fmul_rn((float)c.x, L.scale);
    k.y = __fmul_rn((float)c.y, L.scale);
    k.size = __fmul_rn(31.0f, L.scale);
    k.angle = angle;
    k.response = c.resp;
    k.octave = l;
    out_kp[o] = k;
  }
  }
}

// ---- the blur ORB applies before the descriptors: separable float 7-tap Gaussian ------------------------------
__global__ void blur_rows_kernel(const uint8_t* __restrict__ img, int W, int H, int pitch, float* __restrict__ tmp) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  float s = 0.f;
  if (x >= 3 && x < W - 3) {
    const uint8_t* p = img + (size_t)y * pitch + x - 3;
#pragma unroll
    for (int t = 0; t < 7; ++t) s = __fadd_rn(s, __fmul_rn(c_gauss[t], (float)p[t]));
  }
  tmp[(size_t)y * pitch + x] = s;
}
__global__ void blur_cols_kernel(const float* __restrict__ tmp, int W, int H, int pitch, uint8_t* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  float s = 0.f;
  if (y >= 3 && y < H - 3) {
#pragma unroll
    for (int t = 0; t < 7; ++t) s = __fadd_rn(s, __fmul_rn(c_gauss[t], tmp[(size_t)(y - 3 + t) * pitch + x]));
  }
  out[(size_t)y * pitch + x] = (uint8_t)min(255, max(0, __float2int_rn(s)));
}

int cv_round(double x) { return (int)std::nearbyint(x); }

const int8_t kPattern[1024] = {
#include "orb_pattern.inc"
};

}  // namespace

const int8_t* orb_pattern() { return kPattern; }

int orb_detect(const uint8_t* gray, int w, int h, int nfeatures, const int8_t* pattern256x4, int max_out, float* out_kp6,
               uint8_t* out_des, int* out_n, OrbScratch* scratch, cudaStream_t stream, std::string* err) {
  auto fail = [&](int code, const std::string& what) {
    if (err) *err = what;
    return code;
  };
  *out_n = 0;
  if (w <= 0 || h <= 0 || nfeatures < 0) return fail(-1, "bad image size / feature count");
  // level geometry and per-level feature budget (orb.cpp: float arithmetic; scaleFactor is the float 1.2f in a double)
  const double scale_factor = (double)1.2f;
  Params P{};
  {
    const float factor = (float)(1.0 / scale_factor);
    float nd = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)kLevels));
    int sum = 0;
    for (int l = 0; l < kLevels - 1; ++l) {
      P.lv[l].n_want = cv_round(nd);
      sum += P.lv[l].n_want;
      nd *= factor;
    }
    P.lv[kLevels - 1].n_want = std::max(nfeatures - sum, 0);
  }
  size_t off = 0;
  int cand_total = 0;
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  for (int l = 0; l < kLevels; ++l) {
    Level& L = P.lv[l];
    L.scale = (float)std::pow(scale_factor, (double)l);
    L.w = l == 0 ? w : cv_round((float)w / L.scale);
    L.h = l == 0 ? h : cv_round((float)h / L.scale);
    L.pitch = L.w + 2 * kBorder;
    const size_t padded = (size_t)L.pitch * (L.h + 2 * kBorder);
    L.img_off = off;
    off += up(padded);
    L.blur_off = off;
    off += up(padded);
    L.score_off = off;
    off += up((size_t)std::max(L.w, 1) * std::max(L.h, 1));
    L.cand_off = cand_total;
    L.cand_cap = std::max(64, (std::max(L.w, 1) * std::max(L.h, 1)) / 4);
    cand_total += L.cand_cap;
  }
  const size_t pyr_bytes = off;
  const size_t tmp_floats = (size_t)P.lv[0].pitch * (P.lv[0].h + 2 * kBorder);
  // one device block kept between calls (grow only): [pyramid][source][float rows][candidates][survivors][ints][kp][des]
  auto up2 = [](size_t x) { return (x + 255) / 256 * 256; };
  const int n_ints = 8 + 8 * 256 + 8 + 8 + 16;
  const size_t o_src = up2(pyr_bytes), o_tmp = o_src + up2((size_t)w * h), o_cand = o_tmp + up2(tmp_floats * sizeof(float));
  const size_t o_keep = o_cand + up2((size_t)cand_total * sizeof(Cand)), o_ints = o_keep + up2((size_t)cand_total * sizeof(Cand));
  const size_t o_kp = o_ints + up2(n_ints * sizeof(int)), o_des = o_kp + up2((size_t)std::max(max_out, 1) * sizeof(KpOut));
  const size_t total_bytes = o_des + up2((size_t)std::max(max_out, 1) * 32);
  cudaError_t e = cudaSuccess;
#define OC(call)                                                          \
  if ((e = (call)) != cudaSuccess) return fail(-2, std::string(#call) + ": " + cudaGetErrorString(e));
  OrbScratch local;
  OrbScratch* sc = scratch ? scratch : &local;
  if (sc->cap < total_bytes) {
    OC(cudaStreamSynchronize(stream));
    if (sc->buf) cudaFree(sc->buf);
    sc->buf = nullptr;
    sc->cap = 0;
    OC(cudaMalloc(&sc->buf, total_bytes));
    sc->cap = total_bytes;
  }
  uint8_t* base_p = static_cast<uint8_t*>(sc->buf);
  uint8_t* d_pyr = base_p;
  uint8_t* d_src = base_p + o_src;
  float* d_tmp = reinterpret_cast<float*>(base_p + o_tmp);
  Cand* d_cand = reinterpret_cast<Cand*>(base_p + o_cand);
  Cand* d_keep = reinterpret_cast<Cand*>(base_p + o_keep);
  int* d_ints = reinterpret_cast<int*>(base_p + o_ints);   // [cand_count 8][hist 8*256][fast_thr 8][out_count 8][kp_base 9 ..]
  KpOut* d_kp = reinterpret_cast<KpOut*>(base_p + o_kp);
  uint8_t* d_des = base_p + o_des;
  OC(cudaMemsetAsync(d_ints, 0, n_ints * sizeof(int), stream));
  int* d_count = d_ints;
  int* d_hist = d_ints + 8;
  int* d_thr = d_hist + 8 * 256;
  int* d_out_count = d_thr + 8;
  int* d_kp_base = d_out_count + 8;
  // constants
  {
    int um[kHalfPatch + 2] = {0};
    const int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1), vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
    for (int v = 0; v <= vmax; ++v) um[v] = cv_round(std::sqrt((double)kHalfPatch * kHalfPatch - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
      while (um[v0] == um[v0 + 1]) ++v0;
      um[v] = v0;
      ++v0;
    }
    OC(cudaMemcpyToSymbolAsync(c_umax, um, sizeof um, 0, cudaMemcpyHostToDevice, stream));
    double kd[7], sum = 0;
    for (int t = 0; t < 7; ++t) {
      kd[t] = std::exp(-((t - 3) * (t - 3)) / 8.0);   // sigma = 2
      sum += kd[t];
    }
    float kf[7];
    for (int t = 0; t < 7; ++t) kf[t] = (float)(kd[t] / sum);
    OC(cudaMemcpyToSymbolAsync(c_gauss, kf, sizeof kf, 0, cudaMemcpyHostToDevice, stream));
    OC(cudaMemcpyToSymbolAsync(c_pattern, pattern256x4, 1024, 0, cudaMemcpyHostToDevice, stream));
  }
  OC(cudaMemcpyAsync(d_src, gray, (size_t)w * h, cudaMemcpyHostToDevice, stream));
  const dim3 blk(32, 8);
  auto grid = [&](int W, int H) { return dim3((W + 31) / 32, (H + 7) / 8); };
  int launches = 0, max_cap = 1;
  for (int l = 0; l < kLevels; ++l) {
    const Level& L = P.lv[l];
    if (L.w < 1 || L.h < 1) continue;
    max_cap = std::max(max_cap, L.cand_cap);
    uint8_t* img = d_pyr + L.img_off;
    if (l == 0)
      upload_kernel<<<grid(L.w, L.h), blk, 0, stream>>>(d_src, w, h, img, L.pitch);
    else
      resize_kernel<<<grid(L.w, L.h), blk, 0, stream>>>(d_pyr + P.lv[l - 1].img_off, P.lv[l - 1], img, L);
    border_kernel<<<grid(L.pitch, L.h + 2 * kBorder), blk, 0, stream>>>(img, L.w, L.h, L.pitch);
    fast_score_kernel<<<grid(L.w, L.h), blk, 0, stream>>>(img, L, d_pyr + L.score_off);
    nms_collect_kernel<<<grid(L.w, L.h), blk, 0, stream>>>(d_pyr + L.score_off, L, l, d_cand, d_count, d_hist);
    blur_rows_kernel<<<grid(L.pitch, L.h + 2 * kBorder), blk, 0, stream>>>(img, L.pitch, L.h + 2 * kBorder, L.pitch, d_tmp);
    blur_cols_kernel<<<grid(L.pitch, L.h + 2 * kBorder), blk, 0, stream>>>(d_tmp, L.pitch, L.h + 2 * kBorder, L.pitch,
                                                                        d_pyr + L.blur_off);
    launches += 6;
  }
  // No host round trip between the stages: grids are sized by the capacities, kernels read the counts on the device
  fast_threshold_kernel<<<1, 32, 0, stream>>>(d_hist, d_count, P, d_thr);
  const int harris_blocks = std::min((max_cap + 255) / 256, 4096);   // more FAST survivors than 1 M per level cannot pass the first cut
  harris_kernel<<<dim3(harris_blocks, kLevels), 256, 0, stream>>>(d_pyr, P, d_count, d_thr, d_cand);
  harris_select_kernel<<<kLevels, 1024, 0, stream>>>(P, d_count, d_cand, d_out_count, d_keep);
  kp_base_kernel<<<1, 1, 0, stream>>>(P, d_out_count, d_kp_base);
  const int desc_blocks = (std::min(max_cap, std::max(max_out, 1)) + 7) / 8;
  describe_kernel<<<dim3(desc_blocks, kLevels), 256, 0, stream>>>(d_pyr, P, d_out_count, d_kp_base, d_keep, d_kp, d_des, max_out);
  launches += 5;
  int h_ints[8 + 8 + 9];
  int h_count[8];
  OC(cudaMemcpyAsync(h_count, d_count, sizeof h_count, cudaMemcpyDeviceToHost, stream));
  OC(cudaMemcpyAsync(h_ints, d_out_count, (8 + 9) * sizeof(int), cudaMemcpyDeviceToHost, stream));
  std::vector<KpOut> hk(std::max(max_out, 1));
  std::vector<uint8_t> hd((size_t)std::max(max_out, 1) * 32);
  if (max_out > 0) {
    OC(cudaMemcpyAsync(hk.data(), d_kp, (size_t)max_out * sizeof(KpOut), cudaMemcpyDeviceToHost, stream));
    OC(cudaMemcpyAsync(hd.data(), d_des, (size_t)max_out * 32, cudaMemcpyDeviceToHost, stream));
  }
  OC(cudaStreamSynchronize(stream));
#undef OC
  for (int l = 0; l < kLevels; ++l)
    if (h_count[l] > P.lv[l].cand_cap) return fail(-5, "more FAST corners than the candidate buffer holds");
  const int total = h_ints[8 + 8];   // kp_base[8]
  if (total > max_out) return fail(-5, "more key points than the output buffers hold");
  if (total > 0) {
    // deterministic order: level, then raster order of the level coordinates (the append order on the device is not)
    std::vector<int> order(total);
    for (int i = 0; i < total; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
      if (hk[a].octave != hk[b].octave) return hk[a].octave < hk[b].octave;
      if (hk[a].y != hk[b].y) return hk[a].y < hk[b].y;
      return hk[a].x < hk[b].x;
    });
    for (int i = 0; i < total; ++i) {
      const KpOut& k = hk[order[i]];
      float* o = out_kp6 + (size_t)i * 6;
      o[0] = k.x; o[1] = k.y; o[2] = k.size; o[3] = k.angle; o[4] = k.response; o[5] = (float)k.octave;
      std::copy(hd.begin() + (size_t)order[i] * 32, hd.begin() + (size_t)order[i] * 32 + 32, out_des + (size_t)i * 32);
    }
  }
  *out_n = total;
  return launches;
}

OrbScratch::~OrbScratch() {
  if (buf) cudaFree(buf);
}

// Debug aid: the FAST-9/16 score map of the grey image itself (pyramid level 0), no suppression.
int orb_debug_fast_scores(const uint8_t* gray, int w, int h, uint8_t* out_score, cudaStream_t stream, std::string* err) {
  Level L{};
  L.w = w;
  L.h = h;
  L.pitch = w + 2 * kBorder;
  uint8_t *d_src = nullptr, *d_img = nullptr, *d_score = nullptr;
  cudaError_t e;
  auto done = [&](int rc, const char* what) {
    if (rc && err) *err = std::string(what) + ": " + cudaGetErrorString(e);
    cudaFree(d_src); cudaFree(d_img); cudaFree(d_score);
    return rc;
  };
  if ((e = cudaMalloc(&d_src, (size_t)w * h)) != cudaSuccess) return done(-2, "cudaMalloc");
  if ((e = cudaMalloc(&d_img, (size_t)L.pitch * (h + 2 * kBorder))) != cudaSuccess) return done(-2, "cudaMalloc");
  if ((e = cudaMalloc(&d_score, (size_t)w * h)) != cudaSuccess) return done(-2, "cudaMalloc");
  if ((e = cudaMemcpyAsync(d_src, gray, (size_t)w * h, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return done(-2, "copy");
  const dim3 blk(32, 8), grd((w + 31) / 32, (h + 7) / 8);
  upload_kernel<<<grd, blk, 0, stream>>>(d_src, w, h, d_img, L.pitch);
  border_kernel<<<dim3((L.pitch + 31) / 32, (h + 2 * kBorder + 7) / 8), blk, 0, stream>>>(d_img, w, h, L.pitch);
  fast_score_kernel<<<grd, blk, 0, stream>>>(d_img, L, d_score);
  if ((e = cudaMemcpyAsync(out_score, d_score, (size_t)w * h, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return done(-2, "copy");
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return done(-2, "sync");
  return done(0, "");
}

}  // namespace iam
