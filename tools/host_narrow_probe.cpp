#include <immintrin.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <chrono>
#include <atomic>
__attribute__((target("avx2"))) int narrow_avx2(const float* s, uint8_t* d, size_t n) {
  __m256i bad = _mm256_setzero_si256();
  const __m256i perm = _mm256_setr_epi32(0,4,1,5,2,6,3,7);
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    __m256 f0 = _mm256_loadu_ps(s+i), f1 = _mm256_loadu_ps(s+i+8), f2 = _mm256_loadu_ps(s+i+16), f3 = _mm256_loadu_ps(s+i+24);
    __m256i i0 = _mm256_cvtps_epi32(f0), i1 = _mm256_cvtps_epi32(f1), i2 = _mm256_cvtps_epi32(f2), i3 = _mm256_cvtps_epi32(f3);
    __m256 e0 = _mm256_cmp_ps(_mm256_cvtepi32_ps(i0), f0, _CMP_NEQ_UQ);
    __m256 e1 = _mm256_cmp_ps(_mm256_cvtepi32_ps(i1), f1, _CMP_NEQ_UQ);
    __m256 e2 = _mm256_cmp_ps(_mm256_cvtepi32_ps(i2), f2, _CMP_NEQ_UQ);
    __m256 e3 = _mm256_cmp_ps(_mm256_cvtepi32_ps(i3), f3, _CMP_NEQ_UQ);
    __m256i o = _mm256_or_si256(_mm256_or_si256(i0,i1), _mm256_or_si256(i2,i3));
    bad = _mm256_or_si256(bad, _mm256_andnot_si256(_mm256_set1_epi32(255), o));
    bad = _mm256_or_si256(bad, _mm256_castps_si256(_mm256_or_ps(_mm256_or_ps(e0,e1), _mm256_or_ps(e2,e3))));
    __m256i p01 = _mm256_packus_epi32(i0, i1), p23 = _mm256_packus_epi32(i2, i3);
    __m256i p = _mm256_packus_epi16(p01, p23);
    p = _mm256_permutevar8x32_epi32(p, perm);
    _mm256_storeu_si256((__m256i*)(d+i), p);
  }
  int b = !_mm256_testz_si256(bad, bad);
  for (; i < n; ++i) { float f = s[i]; int v = (int)f; if ((float)v != f || v < 0 || v > 255) b = 1; d[i] = (uint8_t)v; }
  return b;
}
int main(int argc, char** argv) {
  int T = argc > 1 ? atoi(argv[1]) : 8;
  size_t n = 500ull*5000*128;
  float* s = (float*)aligned_alloc(64, n*4); uint8_t* d = (uint8_t*)aligned_alloc(64, n);
  for (size_t i = 0; i < n; ++i) s[i] = (float)((uint32_t)(i*2654435761u) >> 24);
  for (int rep = 0; rep < 3; ++rep) {
    auto t0 = std::chrono::steady_clock::now();
    std::atomic<size_t> next{0}; std::atomic<int> bad{0};
    const size_t chunk = 512*128;
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back([&]{ for(;;){ size_t c = next.fetch_add(1); size_t o = c*chunk; if (o >= n) break; size_t m = std::min(chunk, n-o); if (narrow_avx2(s+o, d+o, m)) bad = 1; } });
    for (auto& x : th) x.join();
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now()-t0).count();
    printf("T=%d %.2f ms  %.1f GB/s read  bad=%d\n", T, ms, n*4/ms/1e6, bad.load());
  }
  // verify
  for (size_t i = 0; i < n; i += 977) if (d[i] != (uint8_t)s[i]) { printf("MISMATCH %zu\n", i); return 1; }
  s[12345] = 3.5f; printf("bad detect: %d\n", narrow_avx2(s+12288, d+12288, 128));
  s[12345] = 256.f; printf("bad detect: %d\n", narrow_avx2(s+12288, d+12288, 128));
  s[12345] = -1.f; printf("bad detect: %d\n", narrow_avx2(s+12288, d+12288, 128));
  return 0;
}
