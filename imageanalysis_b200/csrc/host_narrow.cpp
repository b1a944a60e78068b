// host_narrow.cpp — see host_narrow.h.  Plain C++ (no CUDA): compiled by the host compiler.
#include "host_narrow.h"

#include <immintrin.h>

#include <algorithm>
#include <cstdlib>

namespace iam {

namespace {

int narrow_scalar(const float* s, uint8_t* d, size_t n) {
  int bad = 0;
  for (size_t i = 0; i < n; ++i) {
    const float f = s[i];
    const int v = (f >= 0.0f && f <= 255.0f) ? static_cast<int>(f) : -1;
    if (v < 0 || static_cast<float>(v) != f) bad = 1;
    d[i] = static_cast<uint8_t>(v);
  }
  return bad;
}

__attribute__((target("avx2"))) int narrow_avx2(const float* s, uint8_t* d, size_t n) {
  __m256i bad = _mm256_setzero_si256();
  const __m256i lanes = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
  const __m256i hi = _mm256_set1_epi32(~255);
  const bool aligned = (reinterpret_cast<uintptr_t>(d) & 31) == 0;
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    const __m256 f0 = _mm256_loadu_ps(s + i), f1 = _mm256_loadu_ps(s + i + 8);
    const __m256 f2 = _mm256_loadu_ps(s + i + 16), f3 = _mm256_loadu_ps(s + i + 24);
    const __m256i i0 = _mm256_cvtps_epi32(f0), i1 = _mm256_cvtps_epi32(f1);
    const __m256i i2 = _mm256_cvtps_epi32(f2), i3 = _mm256_cvtps_epi32(f3);
    // not an integer (or NaN): converting back does not give the input
    const __m256 e0 = _mm256_cmp_ps(_mm256_cvtepi32_ps(i0), f0, _CMP_NEQ_UQ);
    const __m256 e1 = _mm256_cmp_ps(_mm256_cvtepi32_ps(i1), f1, _CMP_NEQ_UQ);
    const __m256 e2 = _mm256_cmp_ps(_mm256_cvtepi32_ps(i2), f2, _CMP_NEQ_UQ);
    const __m256 e3 = _mm256_cmp_ps(_mm256_cvtepi32_ps(i3), f3, _CMP_NEQ_UQ);
    bad = _mm256_or_si256(bad, _mm256_castps_si256(_mm256_or_ps(_mm256_or_ps(e0, e1), _mm256_or_ps(e2, e3))));
    // outside 0..255: any bit above the low byte (negative values have the sign bit)
    bad = _mm256_or_si256(bad, _mm256_and_si256(hi, _mm256_or_si256(_mm256_or_si256(i0, i1), _mm256_or_si256(i2, i3))));
    const __m256i p = _mm256_packus_epi16(_mm256_packus_epi32(i0, i1), _mm256_packus_epi32(i2, i3));
    const __m256i o = _mm256_permutevar8x32_epi32(p, lanes);
    // the bytes are read next by the DMA engine, not by this core: streaming stores skip the read-for-ownership
    if (aligned) _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), o);
    else _mm256_storeu_si256(reinterpret_cast<__m256i*>(d + i), o);
  }
  if (aligned) _mm_sfence();
  int b = !_mm256_testz_si256(bad, bad);
  if (i < n) b |= narrow_scalar(s + i, d + i, n - i);
  return b;
}

}  // namespace

int narrow_f32_to_u8(const float* src, uint8_t* dst, size_t n) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  return avx2 ? narrow_avx2(src, dst, n) : narrow_scalar(src, dst, n);
}

int narrow_default_threads() {
  if (const char* e = getenv("IAM_HOST_THREADS")) return std::max(0, atoi(e));
  int hw = static_cast<int>(std::thread::hardware_concurrency());
  int ranks = 1;
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
  return std::max(0, std::min(16, hw / ranks) - 1);
}

NarrowPool::NarrowPool(int threads) {
  // in order only when this process has the host to itself: several ranks narrowing at once share the memory
  // bandwidth, and the from-the-end scheme is the one that cannot lose against the plain float32 upload
  int ranks = 1;
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
  backward_ = threads < 7 || ranks > 2;
  if (const char* e = getenv("IAM_NARROW_ORDER")) backward_ = e[0] == 'b';  // A/B aid: "fwd" / "bwd"
  for (int t = 0; t < threads; ++t) workers_.emplace_back([this] { run(); });
}

NarrowPool::~NarrowPool() {
  {
    std::lock_guard<std::mutex> g(m_);
    stop_ = true;
  }
  wake_.notify_all();
  for (auto& w : workers_) w.join();
}

void NarrowPool::start(NarrowJob* jobs, int n_jobs) {
  {
    std::lock_guard<std::mutex> g(m_);
    jobs_ = jobs;
    n_jobs_ = n_jobs;
    cursor_.store(0, std::memory_order_relaxed);
    active_ = static_cast<int>(workers_.size());
    ++generation_;
  }
  wake_.notify_all();
}

void NarrowPool::finish() {
  std::unique_lock<std::mutex> g(m_);
  idle_.wait(g, [this] { return active_ == 0; });
  jobs_ = nullptr;
  n_jobs_ = 0;
}

void NarrowPool::run() {
  uint64_t seen = 0;
  for (;;) {
    NarrowJob* jobs;
    int n;
    bool backward;
    {
      std::unique_lock<std::mutex> g(m_);
      wake_.wait(g, [&] { return stop_ || generation_ != seen; });
      if (stop_) return;
      seen = generation_;
      jobs = jobs_;
      n = n_jobs_;
      backward = backward_;
    }
    for (;;) {
      // Enough workers to outrun the bus (measured: >= 7 on a 16-core host narrow faster than the GPU matches): in
      // order, just ahead of the owner.  Fewer: LAST job first -- the owner walks the list from the front, sending
      // float32 rows of images nobody has narrowed yet, and the two meet where the bus and the workers finish
      // together, so a slow host never makes the call slower than the plain float32 upload.
      const int k = cursor_.fetch_add(1, std::memory_order_relaxed);
      if (k >= n) break;
      const int i = backward ? n - 1 - k : k;
      int expect = NarrowJob::kFree;
      if (!jobs[i].state.compare_exchange_strong(expect, NarrowJob::kBusy, std::memory_order_acq_rel)) continue;
      const int bad = narrow_f32_to_u8(jobs[i].src, jobs[i].dst, jobs[i].n);
      jobs[i].state.store(bad ? NarrowJob::kBad : NarrowJob::kDone, std::memory_order_release);
    }
    {
      std::lock_guard<std::mutex> g(m_);
      if (--active_ == 0) idle_.notify_all();
    }
  }
}

}  // namespace iam
