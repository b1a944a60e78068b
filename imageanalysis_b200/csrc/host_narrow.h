// host_narrow.h — host-side float32 -> uint8 narrowing of integer-valued descriptors, in worker threads, so that
// iam_match_images moves a quarter of the bytes across PCIe.  The reference keeps SIFT descriptors as float32
// numpy arrays (scripts/lib/image.py:160-180) although every component is an integer in 0..255;
// uploading them as they are makes the end-to-end path PCIe-bound (1.28 GB per 500 frames).  Narrowing is
// transport only: the distance arithmetic stays on the GPU, and a descriptor that is not an integer in 0..255
// is reported so that the caller sends the original float32 rows instead.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

namespace iam {

// Returns 0 when every src[i] is an integer in 0..255 (dst[i] = that integer), 1 otherwise (dst undefined).
int narrow_f32_to_u8(const float* src, uint8_t* dst, size_t n);

struct NarrowJob {
  enum State : int { kFree = 0, kBusy = 1, kDone = 2, kTaken = 3, kBad = 4 };
  const float* src = nullptr;
  uint8_t* dst = nullptr;
  size_t n = 0;                   // elements
  std::atomic<int> state{kFree};  // kTaken: the owner sends the float32 rows itself; kBad: not narrowable
};

// A fixed set of worker threads that walk a job list (in order, or from its end: see run()).  One list at a time:
// start() -> finish().
class NarrowPool {
 public:
  explicit NarrowPool(int threads);
  ~NarrowPool();
  NarrowPool(const NarrowPool&) = delete;
  NarrowPool& operator=(const NarrowPool&) = delete;
  int threads() const { return static_cast<int>(workers_.size()); }
  bool backward() const { return backward_; }
  void start(NarrowJob* jobs, int n_jobs);
  void finish();  // returns when no worker touches the job list any more

 private:
  void run();
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable wake_, idle_;
  NarrowJob* jobs_ = nullptr;
  int n_jobs_ = 0;
  std::atomic<int> cursor_{0};
  int active_ = 0;
  uint64_t generation_ = 0;
  bool stop_ = false;
  bool backward_ = false;  // order in which the workers walk a job list (run())
};

// Threads to use for one context: IAM_HOST_THREADS, else the hardware threads divided among the local ranks
// (LOCAL_WORLD_SIZE, set by torchrun), at most 16, minus the caller's own thread.  0 disables narrowing.
int narrow_default_threads();

}  // namespace iam
