// iamatch.cu — the C-ABI of libiamatch.so (declared in include/iamatch.h):
// context, descriptor residency, work-list construction and kernel
// sequencing.  Device code lives in knn_umma.cu / knn_simt.cu / convert.cu /
// reduce.cu / ransac.cu.  No CPU compute path exists here by design: every
// entry point either runs CUDA kernels or fails with an error code.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/iamatch.h"
#include "ba.h"
#include "gms.h"
#include <memory>
#include <thread>

#include "host_narrow.h"
#include "knn.h"
#include "layout.h"
#include "orb.h"
#include "sift.h"
#include "triangulate.h"
#include "ransac.h"
#include "reduce.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return fail(e__ == cudaErrorMemoryAllocation ? IAM_E_NOMEM : IAM_E_CUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString(e__), __FILE__, __LINE__);                               \
  } while (0)

struct Buffer {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

struct Image {
  int* keys = nullptr;
  float2* kp = nullptr;       // keypoint pixel coordinates (GMS), kp_n entries
  int kp_n = 0;
  uint8_t* block = nullptr;   // [256 B meta words][raw][a_form][b_form] (+ L2: [byte form][perm][rowc][norms][even mask])
  size_t block_bytes = 0;
  int n = -1;
  int n_pad = 0;
  int exact = 1;              // -1: not read back from the device yet
  int i8ok = 0;               // byte layout usable (integer valued, norms within capacity); -1: not read back yet
  bool has_wide = false;      // a_form / b_form hold this descriptor set
  bool has_i8 = false;        // i8_form / perm / rowc hold this descriptor set
  cudaEvent_t ready = nullptr;  // recorded on the upload stream after the layout conversion
  uint64_t seq = 0;           // upload order; a later seq completing implies every earlier one did
  iam::ImgDev dev{};
};

struct Plan {
  std::vector<iam::KnnUnit> units;
  std::vector<iam::RedJob> jobs;
  std::vector<int> chunk_pair_begin;  // pair index where each chunk starts (+ sentinel)
  std::vector<int> chunk_unit_begin;
  std::vector<size_t> chunk_rows;
  std::vector<int32_t> pairs;         // the list this plan was built for (a hash hit is confirmed against it)
  int max_n = 0;
  size_t max_chunk_rows = 0;
};

}  // namespace

struct iam_ctx {
  int device = 0;
  int norm = 0;
  int desc_bytes = 0;
  int num_sms = 0;
  int engine = IAM_ENGINE_AUTO;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;     // compute stream (own_stream or the caller's)
  cudaStream_t up_stream = nullptr;  // H2D + layout conversion; overlaps with matching of earlier pair chunks
  cudaStream_t up_stream2 = nullptr; // second upload lane (iam_match_images alternates images between the two, so the
                                     // copy of one image overlaps the conversion of the previous one)
  Buffer stage2;
  cudaEvent_t lane2_ev = nullptr;
  cudaStream_t dl_stream = nullptr;  // iam_match_images: a wave's tables go back to the host while later waves still compute
  std::vector<cudaEvent_t> done_ev;  // one per wave: its rows of out_table / out_count are final
  unsigned up_rr = 0;
  cudaEvent_t compute_done = nullptr;
  bool compute_pending = false;
  uint64_t up_seq = 0;
  std::vector<Image> images;
  bool imgs_dirty = true;
  Buffer d_imgs, stage;
  Buffer units, jobs, knn_idx, knn_dist, cand_metric, cand_qt, job_table, job_count, out_table, out_count, packed_i,
      packed_d, csr_rows, csr_off;
  int last_pairs = 0, last_cap = 0;
  Plan plan;                 // cached work list: rebuilt only when the pair list or an image's size changes
  uint64_t plan_key = 0;
  bool plan_valid = false;
  uint64_t shape_epoch = 1;
  bool timing_pending = false;
  bool profiling = false;
  cudaEvent_t ev[6] = {};
  std::vector<cudaEvent_t> chunk_ev;  // profiling: (kNN start, kNN end, reductions end) per chunk of the last match call
  int n_prof_chunks = 0;
  std::vector<cudaEvent_t> wave_ev;   // one event per wave of uploads in iam_match_images
  int reserve_sms = 0;                // SMs left free for conversion kernels while uploads are in flight
  bool feed_mode = false;             // inside iam_match_images: no per-image memset / event
  bool feed_wide = false;             // ... and its uploads build the wide (fp16) forms instead of the byte layout
  bool l2_wide_sticky = false;        // a feed-mode call met descriptors the byte layout cannot hold: stay on fp16 operands
  int* d_ctx_flag = nullptr;          // device word, cleared by a conversion that met such descriptors
  int last_kind = -1;
  // bundle-adjustment problem (iam_ba_*): structure resident, parameters re-uploaded per evaluation
  Buffer ba_params, ba_cam_idx, ba_pt_idx, ba_obs, ba_res, ba_jac, ba_jac_cal;
  // robust fits (iam_ransac_*): device block kept between calls, outputs of the table form
  iam::RansacScratch ransac_scratch;
  iam::OrbScratch orb_scratch;
  iam::SiftScratch sift_scratch;
  Buffer ransac_mask, ransac_model, ransac_inl;
  Buffer tri_in, tri_out;
  int ba_n_cam = 0, ba_n_pts = 0, ba_n_obs = -1;
  // iam_match_images, float32 L2 descriptors: worker threads narrow them to bytes (host_narrow.h) into a pinned arena
  std::unique_ptr<iam::NarrowPool> narrow_pool;
  bool narrow_pool_tried = false;
  uint8_t* narrow_arena = nullptr;
  size_t narrow_cap = 0;
  std::unique_ptr<iam::NarrowJob[]> narrow_jobs;
  int narrow_jobs_cap = 0;
  unsigned long long h2d_bytes = 0;   // descriptor bytes copied host -> device by the current iam_match_images call
  int narrowed_images = 0;
  cudaEvent_t probe_ev[2] = {};  // last upload enqueued on each lane by iam_match_images
  // batched conversion of a wave's uint8 images on the compute stream (iam_match_images)
  Buffer conv_jobs_d;
  iam::ConvJob* conv_jobs_h = nullptr;   // page-locked
  int conv_jobs_cap = 0;
  cudaEvent_t span[4] = {};   // upload first/last, compute first/last of the last iam_match_images call
  bool span_pending = false;
  iam_timing timing{};
};

namespace {

int bind(iam_ctx* c) {
  if (!c) return fail(IAM_E_ARG, "null context");
  CU(cudaSetDevice(c->device));
  return IAM_OK;
}

int raw_row_bytes(const iam_ctx* c) { return c->desc_bytes; }

bool umma_capable(const iam_ctx* c) {
  return (c->norm == IAM_NORM_L2 && c->desc_bytes <= 128 && c->desc_bytes % 4 == 0) ||
         (c->norm == IAM_NORM_HAMMING && c->desc_bytes <= 32);
}
bool simt_capable(const iam_ctx* c) { return c->desc_bytes == 32 || c->desc_bytes == 64 || c->desc_bytes == 128; }

int sync_imgs(iam_ctx* c) {
  if (!c->imgs_dirty) return IAM_OK;
  const size_t n = c->images.size();
  if (n == 0) return IAM_OK;
  std::vector<iam::ImgDev> host(n);
  for (size_t i = 0; i < n; ++i) host[i] = c->images[i].dev;
  CU(c->d_imgs.ensure(n * sizeof(iam::ImgDev)));
  CU(cudaMemcpyAsync(c->d_imgs.p, host.data(), n * sizeof(iam::ImgDev), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));  // `host` goes out of scope
  c->imgs_dirty = false;
  return IAM_OK;
}

int check_image(const iam_ctx* c, int id) {
  if (id < 0 || id >= (int)c->images.size() || c->images[id].n < 0)
    return fail(IAM_E_STATE, "image %d has no descriptors uploaded", id);
  return IAM_OK;
}

// Offsets of the parts of an image's device block.
struct BlockLayout {
  size_t raw, a_form, b_form, i8_form, perm, rowc, nrm, mask, total;
};
BlockLayout block_layout(const iam_ctx* c, int n_pad) {
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  BlockLayout b{};
  b.raw = 256;
  b.a_form = b.raw + up(size_t(n_pad) * c->desc_bytes);
  b.b_form = b.a_form + iam::form_bytes(n_pad);
  b.i8_form = b.b_form + iam::form_bytes(n_pad);
  if (c->norm == IAM_NORM_L2) {
    b.perm = b.i8_form + iam::i8_form_bytes(n_pad);
    b.rowc = b.perm + up(size_t(n_pad) * 4);
    b.nrm = b.rowc + up(size_t(n_pad) * 4);
    b.mask = b.nrm + up(size_t(n_pad) * 4);
    b.total = b.mask + up(size_t(n_pad) / 8);
  } else {
    b.perm = b.rowc = b.nrm = b.mask = b.total = b.i8_form;
  }
  return b;
}

// Read an image's flags (integer valued? byte layout usable?) back from the device on first use.
int resolve_flags(iam_ctx* c, int id) {
  Image& im = c->images[id];
  if (im.exact < 0 || im.i8ok < 0) {
    int flag[2] = {0, 0};
    if (cudaMemcpyAsync(flag, im.block, sizeof flag, cudaMemcpyDeviceToHost, c->up_stream) != cudaSuccess ||
        cudaStreamSynchronize(c->up_stream) != cudaSuccess)
      return -1;
    if (im.exact < 0) im.exact = flag[iam::kMetaExact] != 0;
    if (im.i8ok < 0) im.i8ok = flag[iam::kMetaI8Ok] != 0;
  }
  return 0;
}
int resolve_exact(iam_ctx* c, int id) {
  if (resolve_flags(c, id) != 0) return -1;
  return c->images[id].exact;
}

// Make the compute stream wait for the uploads of every image a chunk of pairs touches.
int wait_uploads(iam_ctx* c, const int32_t* pairs, int p0, int p1) {
  uint64_t best = 0;
  int best_id = -1;
  for (int i = 2 * p0; i < 2 * p1; ++i) {
    const Image& im = c->images[pairs[i]];
    if (im.seq > best) {
      best = im.seq;
      best_id = pairs[i];
    }
  }
  if (best_id >= 0 && c->images[best_id].ready) CU(cudaStreamWaitEvent(c->stream, c->images[best_id].ready, 0));
  return IAM_OK;
}

// After the last kernel of a call: later uploads must not overwrite operands still being read.
int mark_compute(iam_ctx* c) {
  CU(cudaEventRecord(c->compute_done, c->stream));
  c->compute_pending = true;
  return IAM_OK;
}

// Which tensor-core operand kind serves this pair list.  L2: the byte layout (kind::i8) when every image involved
// holds one; fp16 operands otherwise.  Inside iam_match_images the images are not uploaded yet: the call's own
// decision (feed_wide) says what its uploads build.
int pick_kind(iam_ctx* c, const int32_t* pairs, int n_pairs, int* kind) {
  if (c->norm == IAM_NORM_HAMMING) {
    *kind = iam::kKindF8;
    return IAM_OK;
  }
  bool i8 = c->engine != IAM_ENGINE_UMMA_F16;
  if (c->feed_mode) {
    i8 = !c->feed_wide;
  } else {
    for (int p = 0; p < n_pairs * 2 && i8; ++p) {
      Image& im = c->images[pairs[p]];
      if (!im.has_i8) i8 = false;
      else if (resolve_flags(c, pairs[p]) != 0) return fail(IAM_E_CUDA, "flag read-back failed");
      else if (im.i8ok != 1) i8 = false;
    }
    if (!i8)
      for (int p = 0; p < n_pairs * 2; ++p)
        if (!c->images[pairs[p]].has_wide)
          return fail(IAM_E_STATE, "image %d holds only the byte layout but this pair list needs fp16 operands: upload it again", pairs[p]);
  }
  *kind = i8 ? iam::kKindI8 : iam::kKindF16;
  return IAM_OK;
}

int pick_engine(const iam_ctx* c, const int32_t* pairs, int n_pairs, int* engine) {
  int e = c->engine;
  if (e == IAM_ENGINE_AUTO) e = umma_capable(c) ? IAM_ENGINE_UMMA : IAM_ENGINE_SIMT;
  if (e == IAM_ENGINE_UMMA_F16) e = IAM_ENGINE_UMMA;
  if (e == IAM_ENGINE_UMMA && !umma_capable(c)) return fail(IAM_E_UNSUPPORTED, "descriptor size %d has no tensor-core layout", c->desc_bytes);
  if (e == IAM_ENGINE_SIMT) {
    if (!simt_capable(c)) return fail(IAM_E_UNSUPPORTED, "SIMT engine supports 32/64/128-byte descriptors, got %d", c->desc_bytes);
    for (int p = 0; p < n_pairs * 2; ++p)
      if (resolve_exact(const_cast<iam_ctx*>(c), pairs[p]) == 0) return fail(IAM_E_UNSUPPORTED, "SIMT engine needs integer-valued descriptors (image %d)", pairs[p]);
  }
  *engine = e;
  return IAM_OK;
}

// Two directed jobs per pair (2p: i->j, 2p+1: j->i); out_base restarts per chunk.
int build_plan(const iam_ctx* c, const int32_t* pairs, int n_pairs, int k, bool both, int waves, Plan* pl) {
  const int max_pairs = waves > 1 ? std::max(64, (n_pairs + waves - 1) / waves) : (1 << 30);
  int in_chunk = 0;
  bool ramp = waves >= 8;
  if (const char* env = getenv("IAM_WAVE_RAMP")) ramp = ramp && atoi(env) != 0;  // A/B aid
  size_t budget_mb = 768;  // kNN output workspace per chunk; IAM_CHUNK_MB overrides (tests force many chunks)
  if (const char* env = getenv("IAM_CHUNK_MB")) budget_mb = std::max(1, atoi(env));
  const size_t budget_rows = (budget_mb << 20) / (size_t(k) * 8);
  size_t rows = 0;
  pl->chunk_pair_begin.push_back(0);
  pl->chunk_unit_begin.push_back(0);
  for (int p = 0; p < n_pairs; ++p) {
    const int i = pairs[2 * p], j = pairs[2 * p + 1];
    int rc;
    if ((rc = check_image(c, i)) != IAM_OK || (rc = check_image(c, j)) != IAM_OK) return rc;
    const Image& a = c->images[i];
    const Image& b = c->images[j];
    const size_t need = size_t(a.n_pad) + (both ? size_t(b.n_pad) : 0);
    // Waves (uploads interleaved with matching): the first two are short, so that the first kernel starts after
    // a handful of images has crossed the bus instead of a sixteenth of the project.
    int limit = max_pairs;
    if (ramp && pl->chunk_rows.size() == 0) limit = std::max(16, max_pairs / 8);
    if (ramp && pl->chunk_rows.size() == 1) limit = std::max(32, max_pairs / 2);
    if (rows > 0 && (rows + need > budget_rows || in_chunk >= limit)) {
      in_chunk = 0;
      pl->chunk_rows.push_back(rows);
      pl->max_chunk_rows = std::max(pl->max_chunk_rows, rows);
      pl->chunk_pair_begin.push_back(p);
      pl->chunk_unit_begin.push_back((int)pl->units.size());
      rows = 0;
    }
    ++in_chunk;
    for (int dir = 0; dir < (both ? 2 : 1); ++dir) {
      const int qs = dir ? j : i, ts = dir ? i : j;
      const Image& q = c->images[qs];
      const Image& t = c->images[ts];
      iam::RedJob jb{(int)rows, q.n, t.n, qs, ts, {0, 0, 0}};
      pl->jobs.push_back(jb);
      const int supers = (q.n + iam::kSuperRows - 1) / iam::kSuperRows;
      for (int s = 0; s < supers; ++s) pl->units.push_back(iam::KnnUnit{qs, ts, s, (int)rows});
      // the clustered kernel walks units in pairs that share the train image: keep every job even
      // (the duplicate recomputes and rewrites identical results)
      if (supers & 1) pl->units.push_back(iam::KnnUnit{qs, ts, supers - 1, (int)rows});
      rows += q.n_pad;
      pl->max_n = std::max(pl->max_n, std::max(q.n, t.n));
    }
  }
  pl->chunk_rows.push_back(rows);
  pl->max_chunk_rows = std::max(pl->max_chunk_rows, rows);
  pl->chunk_pair_begin.push_back(n_pairs);
  pl->chunk_unit_begin.push_back((int)pl->units.size());
  return IAM_OK;
}

// Uploads still in flight on the upload stream?  Then matching is cut into waves so
// that early pair chunks run while later images are still crossing PCIe.
int pick_waves(iam_ctx* c) {
  uint64_t best = 0;
  cudaEvent_t ev = nullptr;
  for (const Image& im : c->images)
    if (im.seq > best) {
      best = im.seq;
      ev = im.ready;
    }
  if (!ev) return 1;
  const cudaError_t q = cudaEventQuery(ev);
  if (q == cudaErrorNotReady) {
    cudaGetLastError();
    return 8;
  }
  return 1;
}

uint64_t plan_hash(const iam_ctx* c, const int32_t* pairs, int n_pairs, int k, bool both, int waves) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&h](uint64_t v) {
    for (int b = 0; b < 8; ++b) {
      h ^= (v >> (8 * b)) & 0xff;
      h *= 1099511628211ull;
    }
  };
  mix(c->shape_epoch);
  mix(uint64_t(n_pairs));
  mix(uint64_t(k) * 2 + (both ? 1 : 0) + 16 * uint64_t(waves));
  if (const char* env = getenv("IAM_CHUNK_MB")) mix(uint64_t(atoi(env)) + 77);
  for (int i = 0; i < 2 * n_pairs; ++i) mix(uint64_t(uint32_t(pairs[i])));
  return h | 1ull;
}

int upload_plan(iam_ctx* c, const Plan& pl);

// Build (or reuse) the work list for this pair list and make it resident.
int prepare_plan(iam_ctx* c, const int32_t* pairs, int n_pairs, int k, bool both, int waves_override = 0) {
  const int waves = waves_override > 0 ? waves_override : pick_waves(c);
  const uint64_t key = plan_hash(c, pairs, n_pairs, k, both, waves);
  if (c->plan_valid && c->plan_key == key && c->plan.pairs.size() == size_t(n_pairs) * 2 &&
      (n_pairs == 0 || memcmp(c->plan.pairs.data(), pairs, size_t(n_pairs) * 2 * sizeof(int32_t)) == 0))
    return IAM_OK;
  c->plan_valid = false;
  c->plan = Plan{};
  int rc = build_plan(c, pairs, n_pairs, k, both, waves, &c->plan);
  if (rc != IAM_OK) return rc;
  c->plan.pairs.assign(pairs, pairs + size_t(n_pairs) * 2);
  if ((rc = upload_plan(c, c->plan)) != IAM_OK) return rc;
  c->plan_key = key;
  c->plan_valid = true;
  return IAM_OK;
}

int upload_plan(iam_ctx* c, const Plan& pl) {
  CU(c->units.ensure(std::max<size_t>(1, pl.units.size()) * sizeof(iam::KnnUnit)));
  CU(c->jobs.ensure(std::max<size_t>(1, pl.jobs.size()) * sizeof(iam::RedJob)));
  if (!pl.units.empty())
    CU(cudaMemcpyAsync(c->units.p, pl.units.data(), pl.units.size() * sizeof(iam::KnnUnit), cudaMemcpyHostToDevice, c->stream));
  if (!pl.jobs.empty())
    CU(cudaMemcpyAsync(c->jobs.p, pl.jobs.data(), pl.jobs.size() * sizeof(iam::RedJob), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));  // pageable sources must outlive the copies
  return IAM_OK;
}

int launch_knn(iam_ctx* c, int engine, int kind, int k, int unit_begin, int n_units) {
  const iam::KnnUnit* u = c->units.as<iam::KnnUnit>() + unit_begin;
  cudaError_t e;
  c->timing.mma_kind = engine == IAM_ENGINE_UMMA ? kind : -1;
  if (engine == IAM_ENGINE_UMMA)
    e = iam::launch_knn_umma(kind, k, c->d_imgs.as<iam::ImgDev>(), u, n_units, c->knn_idx.as<int>(), c->knn_dist.as<float>(), std::max(2, c->num_sms - c->reserve_sms), c->stream);
  else
    e = iam::launch_knn_simt(c->norm, k, raw_row_bytes(c), c->d_imgs.as<iam::ImgDev>(), u, n_units, c->knn_idx.as<int>(), c->knn_dist.as<float>(), c->stream);
  if (e != cudaSuccess) return fail(IAM_E_CUDA, "kNN launch failed: %s", cudaGetErrorString(e));
  c->timing.knn_launches += 1;
  c->timing.total_launches += 1;
  c->timing.engine_used = engine;
  return IAM_OK;
}

__global__ void pack_knn_kernel(const iam::RedJob* jobs, int job_begin, int n_jobs, int dirs, int dir, int k, int n_stride,
                                const int* idx, const float* dist, int* out_i, float* out_d) {
  // grid.x = job (of this direction) within the chunk, threads stride over rows*k
  const int jl = blockIdx.x;
  const iam::RedJob jb = jobs[job_begin + jl * dirs + dir];
  const size_t src = size_t(jb.out_base) * k;
  const size_t dst = size_t(jl) * n_stride * k;
  const int rows = min(jb.n_q, n_stride);
  for (int e = threadIdx.x; e < rows * k; e += blockDim.x) {
    out_i[dst + e] = idx[src + e];
    out_d[dst + e] = dist[src + e];
  }
}

}  // namespace

extern "C" {

int iam_abi_version(void) { return 6; }
const char* iam_last_error(void) { return g_err.c_str(); }

int iam_create(int device, int norm, int desc_bytes, iam_ctx** out) {
  if (!out) return fail(IAM_E_ARG, "out is null");
  *out = nullptr;
  if (norm != IAM_NORM_L2 && norm != IAM_NORM_HAMMING) return fail(IAM_E_ARG, "unknown norm %d", norm);
  if (desc_bytes <= 0 || desc_bytes > 128 || (desc_bytes & 3)) return fail(IAM_E_ARG, "descriptor size %d not in 4..128 (multiple of 4)", desc_bytes);
  int count = 0;
  CU(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return fail(IAM_E_CUDA, "CUDA device %d not present (%d visible)", device, count);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(IAM_E_CUDA, "libiamatch is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
  iam_ctx* c = new (std::nothrow) iam_ctx();
  if (!c) return fail(IAM_E_NOMEM, "host allocation failed");
  c->device = device;
  c->norm = norm;
  c->desc_bytes = desc_bytes;
  c->num_sms = prop.multiProcessorCount;
  cudaError_t e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete c;
    return fail(IAM_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
  }
  c->stream = c->own_stream;
  e = cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->up_stream2, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->lane2_ev, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->dl_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->compute_done, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    delete c;
    return fail(IAM_E_CUDA, "stream/event creation: %s", cudaGetErrorString(e));
  }
  for (auto& ev : c->ev) cudaEventCreate(&ev);
  for (auto& ev : c->span) cudaEventCreate(&ev);
  for (auto& ev : c->probe_ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  if (cudaMalloc(reinterpret_cast<void**>(&c->d_ctx_flag), 256) != cudaSuccess ||
      cudaMemset(c->d_ctx_flag, 1, 256) != cudaSuccess) {
    iam_destroy(c);
    return fail(IAM_E_NOMEM, "context flag allocation failed");
  }
  *out = c;
  return IAM_OK;
}

int iam_destroy(iam_ctx* c) {
  if (!c) return IAM_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->up_stream) cudaStreamSynchronize(c->up_stream);
  if (c->up_stream2) cudaStreamSynchronize(c->up_stream2);
  for (auto& im : c->images) {
    if (im.block) cudaFree(im.block);
    if (im.keys) cudaFree(im.keys);
    if (im.kp) cudaFree(im.kp);
    if (im.ready) cudaEventDestroy(im.ready);
  }
  c->narrow_pool.reset();
  if (c->narrow_arena) cudaFreeHost(c->narrow_arena);
  if (c->conv_jobs_h) cudaFreeHost(c->conv_jobs_h);
  c->conv_jobs_d.release();
  if (c->d_ctx_flag) cudaFree(c->d_ctx_flag);
  if (c->compute_done) cudaEventDestroy(c->compute_done);
  if (c->up_stream) cudaStreamDestroy(c->up_stream);
  if (c->up_stream2) cudaStreamDestroy(c->up_stream2);
  if (c->lane2_ev) cudaEventDestroy(c->lane2_ev);
  if (c->dl_stream) {
    cudaStreamSynchronize(c->dl_stream);
    cudaStreamDestroy(c->dl_stream);
  }
  for (cudaEvent_t ev : c->done_ev)
    if (ev) cudaEventDestroy(ev);
  c->stage2.release();
  Buffer* ba_bufs[] = {&c->ba_params, &c->ba_cam_idx, &c->ba_pt_idx, &c->ba_obs, &c->ba_res, &c->ba_jac};
  for (Buffer* b : ba_bufs) b->release();
  Buffer* bufs[] = {&c->d_imgs, &c->stage, &c->units, &c->jobs, &c->knn_idx, &c->knn_dist, &c->cand_metric,
                    &c->cand_qt, &c->job_table, &c->job_count, &c->out_table, &c->out_count, &c->packed_i, &c->packed_d, &c->csr_rows, &c->csr_off};
  for (Buffer* b : bufs) b->release();
  for (auto& ev : c->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->span)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->probe_ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->wave_ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->chunk_ev)
    if (ev) cudaEventDestroy(ev);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
  return IAM_OK;
}

int iam_set_stream(iam_ctx* c, void* s) {
  int rc = bind(c);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->stream));
  c->stream = s ? static_cast<cudaStream_t>(s) : c->own_stream;
  return IAM_OK;
}

int iam_set_engine(iam_ctx* c, int engine) {
  if (!c) return fail(IAM_E_ARG, "null context");
  if (engine < IAM_ENGINE_AUTO || engine > IAM_ENGINE_UMMA_F16) return fail(IAM_E_ARG, "unknown engine %d", engine);
  c->engine = engine;
  return IAM_OK;
}

int iam_synchronize(iam_ctx* c) {
  int rc = bind(c);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->up_stream2));
  CU(cudaStreamSynchronize(c->up_stream));
  CU(cudaStreamSynchronize(c->stream));
  return IAM_OK;
}

int iam_set_profiling(iam_ctx* c, int enable) {
  if (!c) return fail(IAM_E_ARG, "null context");
  c->profiling = enable != 0;
  return IAM_OK;
}

int iam_get_timing(iam_ctx* c, iam_timing* out) {
  if (!c || !out) return fail(IAM_E_ARG, "null argument");
  if (c->profiling && c->up_seq > 0) {
    CU(cudaSetDevice(c->device));
    if (cudaEventSynchronize(c->ev[5]) == cudaSuccess) cudaEventElapsedTime(&c->timing.convert_ms, c->ev[4], c->ev[5]);
    cudaGetLastError();
  }
  if (c->span_pending) {
    CU(cudaSetDevice(c->device));
    CU(cudaEventSynchronize(c->span[1]));
    CU(cudaEventSynchronize(c->span[3]));
    cudaEventElapsedTime(&c->timing.upload_span_ms, c->span[0], c->span[1]);
    cudaEventElapsedTime(&c->timing.compute_span_ms, c->span[2], c->span[3]);
    cudaEventElapsedTime(&c->timing.total_span_ms, c->span[0], c->span[3]);
    cudaGetLastError();
    c->span_pending = false;
  }
  if (c->timing_pending) {
    CU(cudaSetDevice(c->device));
    c->timing.knn_ms = 0.f;
    c->timing.reduce_ms = 0.f;
    for (int k = 0; k < c->n_prof_chunks; ++k) {
      float a = 0.f, b = 0.f;
      CU(cudaEventSynchronize(c->chunk_ev[3 * k + 2]));
      CU(cudaEventElapsedTime(&a, c->chunk_ev[3 * k], c->chunk_ev[3 * k + 1]));
      CU(cudaEventElapsedTime(&b, c->chunk_ev[3 * k + 1], c->chunk_ev[3 * k + 2]));
      c->timing.knn_ms += a;
      c->timing.reduce_ms += b;
    }
    c->timing_pending = false;
  }
  *out = c->timing;
  return IAM_OK;
}

// Size an image slot (allocation, device record) without touching its contents.
static int prepare_image(iam_ctx* c, int id, int n) {
  if ((int)c->images.size() <= id) c->images.resize(id + 1);
  Image& im = c->images[id];
  const int n_pad = std::max(iam::kSuperRows, iam::round_up(iam::round_up(n, iam::kBRows), iam::kSuperRows));  // whole B tiles and whole units
  const BlockLayout bl = block_layout(c, n_pad);
  const size_t total = bl.total;
  if (im.block_bytes < total) {
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->up_stream));
    CU(cudaStreamSynchronize(c->up_stream2));
    if (im.block) CU(cudaFree(im.block));
    im.block = nullptr;
    im.block_bytes = 0;
    CU(cudaMalloc(reinterpret_cast<void**>(&im.block), total));
    im.block_bytes = total;
  }
  if (!im.ready) CU(cudaEventCreateWithFlags(&im.ready, cudaEventDisableTiming));
  if (im.n != n || im.n_pad != n_pad) c->shape_epoch++;
  im.n = n;
  im.n_pad = n_pad;
  im.dev.raw = im.block + bl.raw;
  im.dev.a_form = im.block + bl.a_form;
  im.dev.b_form = im.block + bl.b_form;
  im.dev.i8_form = im.block + bl.i8_form;
  im.dev.perm = reinterpret_cast<const int*>(im.block + bl.perm);
  im.dev.rowc = reinterpret_cast<const int*>(im.block + bl.rowc);
  im.dev.meta = reinterpret_cast<const int*>(im.block);
  im.dev.kp_xy = (im.kp && im.kp_n == n) ? im.kp : nullptr;  // coordinates of another descriptor count are stale
  im.dev.n = n;
  im.dev.n_pad = n_pad;
  c->imgs_dirty = true;
  return IAM_OK;
}

// H2D copy (when the source is on the host) + layout conversion on the upload stream; records im.ready.
static int enqueue_upload(iam_ctx* c, int id, const void* src, bool src_on_host, int dtype, const int32_t* host_keys) {
  Image& im = c->images[id];
  const int n = im.n, n_pad = im.n_pad;
  const BlockLayout bl = block_layout(c, n_pad);
  // Which operand forms this upload builds.  Hamming: the wide (e4m3) forms.  L2: both the fp16 forms and the byte
  // layout, except inside iam_match_images, whose uploads are on the critical path and build only the one its
  // kernels will read.
  const bool l2 = c->norm == IAM_NORM_L2;
  const bool want_i8 = l2 && (c->feed_mode ? !c->feed_wide : true);
  const bool want_wide = !l2 || (c->feed_mode ? c->feed_wide : true);
  if (c->compute_pending) {  // WAR: kernels of the previous call may still read the operands we are about to replace
    CU(cudaStreamWaitEvent(c->up_stream, c->compute_done, 0));
    CU(cudaStreamWaitEvent(c->up_stream2, c->compute_done, 0));
    c->compute_pending = false;
  }
  const bool lane2 = c->feed_mode && ((c->up_rr++ & 1u) != 0u);
  cudaStream_t us = lane2 ? c->up_stream2 : c->up_stream;
  Buffer& stg = lane2 ? c->stage2 : c->stage;
  if (im.keys) {  // keys belong to the previous descriptor set
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFree(im.keys));
    im.keys = nullptr;
    im.dev.kp_key = nullptr;
    c->imgs_dirty = true;
  }
  if (host_keys && n > 0) {
    CU(cudaMalloc(reinterpret_cast<void**>(&im.keys), size_t(n) * sizeof(int)));
    CU(cudaMemcpyAsync(im.keys, host_keys, size_t(n) * sizeof(int), cudaMemcpyHostToDevice, us));
    im.dev.kp_key = im.keys;
    c->imgs_dirty = true;
  }
  const void* dsrc = src;
  if (src_on_host) {
    const size_t bytes = size_t(n) * c->desc_bytes * (dtype == IAM_DTYPE_F32 ? 4 : 1);
    if (stg.cap < bytes) {
      CU(cudaStreamSynchronize(us));  // a conversion may still be reading the old staging buffer
      CU(stg.ensure(std::max<size_t>(bytes, 256)));
    }
    if (bytes) CU(cudaMemcpyAsync(stg.p, src, bytes, cudaMemcpyHostToDevice, us));
    c->h2d_bytes += bytes;
    dsrc = stg.p;
  }
  // flag words: non-zero = exact / byte layout usable.  Hamming sources are exact by construction: no flag traffic.
  if (l2) CU(cudaMemsetAsync(im.block, 1, 2 * sizeof(int), us));
  if (c->profiling && !c->feed_mode) CU(cudaEventRecord(c->ev[4], us));
  iam::I8Out i8{};
  if (want_i8) {
    i8.form = im.block + bl.i8_form;
    i8.perm = reinterpret_cast<int*>(im.block + bl.perm);
    i8.rowc = reinterpret_cast<int*>(im.block + bl.rowc);
    i8.nrm = reinterpret_cast<int*>(im.block + bl.nrm);
    i8.even_mask = reinterpret_cast<uint32_t*>(im.block + bl.mask);
    i8.ctx_flag = c->feed_mode ? c->d_ctx_flag : nullptr;
  }
  cudaError_t e = iam::launch_convert(c->norm, c->desc_bytes, dsrc, dtype, n, n_pad, im.block + bl.raw,
                                      want_wide ? im.block + bl.a_form : nullptr, want_wide ? im.block + bl.b_form : nullptr,
                                      reinterpret_cast<int*>(im.block), want_i8 ? &i8 : nullptr, us);
  if (e != cudaSuccess) return fail(IAM_E_CUDA, "convert launch: %s", cudaGetErrorString(e));
  c->timing.total_launches += want_i8 ? 2 : 1;
  im.has_wide = want_wide;
  im.has_i8 = want_i8;
  im.i8ok = want_i8 ? -1 : 0;
  if (c->profiling && !c->feed_mode) CU(cudaEventRecord(c->ev[5], us));
  if (!c->feed_mode) CU(cudaEventRecord(im.ready, us));  // feed mode: one event per wave instead
  if (c->feed_mode) CU(cudaEventRecord(c->probe_ev[lane2 ? 1 : 0], us));  // "has this lane run dry?" (host narrowing)
  im.seq = c->feed_mode ? 0 : ++c->up_seq;
  im.exact = (c->norm == IAM_NORM_L2 && dtype == IAM_DTYPE_F32) ? -1 : 1;  // resolved lazily (no sync per upload)
  return IAM_OK;
}

// iam_match_images, uint8 source, byte layout: the rows are copied straight to their final place on an upload lane;
// the conversion (norms, parity partition, byte form) is left to ONE batched launch per wave on the compute stream
// (match_core), so that the matching kernel needs to leave no SMs free for per-image conversion kernels.
static int enqueue_copy_u8(iam_ctx* c, int id, const void* src, iam::ConvJob* job) {
  Image& im = c->images[id];
  const int n = im.n, n_pad = im.n_pad;
  const BlockLayout bl = block_layout(c, n_pad);
  if (c->compute_pending) {  // WAR: kernels of the previous call may still read the operands we are about to replace
    CU(cudaStreamWaitEvent(c->up_stream, c->compute_done, 0));
    CU(cudaStreamWaitEvent(c->up_stream2, c->compute_done, 0));
    c->compute_pending = false;
  }
  const bool lane2 = (c->up_rr++ & 1u) != 0u;
  cudaStream_t us = lane2 ? c->up_stream2 : c->up_stream;
  const size_t bytes = size_t(n) * c->desc_bytes;
  CU(cudaMemsetAsync(im.block, 1, 2 * sizeof(int), us));   // flag words: non-zero = exact / byte layout usable
  if (bytes) CU(cudaMemcpyAsync(im.block + bl.raw, src, bytes, cudaMemcpyHostToDevice, us));
  c->h2d_bytes += bytes;
  job->src = im.block + bl.raw;
  job->raw = im.block + bl.raw;
  job->meta = reinterpret_cast<int*>(im.block);
  job->nrm = reinterpret_cast<int*>(im.block + bl.nrm);
  job->even_mask = reinterpret_cast<uint32_t*>(im.block + bl.mask);
  job->form = im.block + bl.i8_form;
  job->perm = reinterpret_cast<int*>(im.block + bl.perm);
  job->rowc = reinterpret_cast<int*>(im.block + bl.rowc);
  job->n = n;
  job->n_pad = n_pad;
  im.has_wide = false;
  im.has_i8 = true;
  im.i8ok = -1;
  CU(cudaEventRecord(c->probe_ev[lane2 ? 1 : 0], us));
  im.seq = 0;
  im.exact = 1;
  return IAM_OK;
}

static int upload_common(iam_ctx* c, int id, const void* src, bool src_on_host, int n, int dtype) {
  int rc = prepare_image(c, id, n);
  if (rc) return rc;
  return enqueue_upload(c, id, src, src_on_host, dtype, nullptr);
}

int iam_upload_descriptors(iam_ctx* c, int id, const void* ptr, int n, int dtype, int pinned) {
  int rc = bind(c);
  if (rc) return rc;
  if (id < 0 || id > (1 << 24)) return fail(IAM_E_ARG, "bad image id %d", id);
  if (n < 0 || (n > 0 && !ptr)) return fail(IAM_E_ARG, "bad descriptor buffer");
  if (dtype != IAM_DTYPE_U8 && dtype != IAM_DTYPE_F32) return fail(IAM_E_ARG, "unknown dtype %d", dtype);
  if (c->norm == IAM_NORM_HAMMING && dtype != IAM_DTYPE_U8) return fail(IAM_E_ARG, "Hamming descriptors must be uint8");
  (void)pinned;
  return upload_common(c, id, ptr, true, n, dtype);
}

int iam_upload_descriptors_device(iam_ctx* c, int id, const void* dptr, int n, int dtype) {
  int rc = bind(c);
  if (rc) return rc;
  if (id < 0 || id > (1 << 24)) return fail(IAM_E_ARG, "bad image id %d", id);
  if (n < 0 || (n > 0 && !dptr)) return fail(IAM_E_ARG, "bad descriptor buffer");
  if (dtype != IAM_DTYPE_U8 && dtype != IAM_DTYPE_F32) return fail(IAM_E_ARG, "unknown dtype %d", dtype);
  if (c->norm == IAM_NORM_HAMMING && dtype != IAM_DTYPE_U8) return fail(IAM_E_ARG, "Hamming descriptors must be uint8");
  return upload_common(c, id, dptr, false, n, dtype);
}

int iam_release_descriptors(iam_ctx* c, int id) {
  int rc = bind(c);
  if (rc) return rc;
  if (id < 0 || id >= (int)c->images.size()) return IAM_OK;
  Image& im = c->images[id];
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaStreamSynchronize(c->up_stream));
  if (im.block) CU(cudaFree(im.block));
  if (im.keys) CU(cudaFree(im.keys));
  if (im.kp) CU(cudaFree(im.kp));
  if (im.ready) CU(cudaEventDestroy(im.ready));
  im = Image{};
  c->shape_epoch++;
  c->imgs_dirty = true;
  return IAM_OK;
}

int iam_upload_keypoint_keys(iam_ctx* c, int id, const int32_t* keys, int n) {
  int rc = bind(c);
  if (rc) return rc;
  if ((rc = check_image(c, id)) != IAM_OK) return rc;
  Image& im = c->images[id];
  if (n != im.n || (n > 0 && !keys)) return fail(IAM_E_ARG, "keys must have one entry per descriptor (%d), got %d", im.n, n);
  for (int i = 0; i < n; ++i)
    if (keys[i] < 0 || keys[i] >= n) return fail(IAM_E_ARG, "key %d out of range [0,%d) at %d", keys[i], n, i);
  if (im.keys) {
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFree(im.keys));
    im.keys = nullptr;
  }
  if (n > 0) {
    CU(cudaMalloc(reinterpret_cast<void**>(&im.keys), size_t(n) * sizeof(int)));
    CU(cudaMemcpyAsync(im.keys, keys, size_t(n) * sizeof(int), cudaMemcpyHostToDevice, c->up_stream));
    CU(cudaStreamSynchronize(c->up_stream));
  }
  im.dev.kp_key = im.keys;
  c->imgs_dirty = true;
  return IAM_OK;
}

static int upload_keypoints_one(iam_ctx* c, int id, const float* xy, int n, bool wait) {
  if (id < 0 || id > (1 << 24)) return fail(IAM_E_ARG, "bad image id %d", id);
  if (n < 0 || (n > 0 && !xy)) return fail(IAM_E_ARG, "bad keypoint buffer");
  if ((int)c->images.size() <= id) c->images.resize(id + 1);
  Image& im = c->images[id];
  // The descriptors may arrive later (iam_match_images uploads them itself): the coordinates are matched
  // against the descriptor count when a GMS run needs them.
  if (im.kp_n < n || !im.kp) {
    CU(cudaStreamSynchronize(c->stream));
    if (im.kp) CU(cudaFree(im.kp));
    im.kp = nullptr;
    im.kp_n = 0;
    if (n > 0) CU(cudaMalloc(reinterpret_cast<void**>(&im.kp), size_t(n) * sizeof(float2)));
  } else if (c->compute_pending) {  // a kernel of the previous call may still read the old coordinates
    CU(cudaStreamSynchronize(c->stream));
  }
  im.kp_n = n;
  if (n > 0) {
    CU(cudaMemcpyAsync(im.kp, xy, size_t(n) * sizeof(float2), cudaMemcpyHostToDevice, c->up_stream));
    if (wait) CU(cudaStreamSynchronize(c->up_stream));
  }
  im.dev.kp_xy = (im.kp && im.kp_n == im.n) ? im.kp : nullptr;
  c->imgs_dirty = true;
  return IAM_OK;
}

int iam_upload_keypoints(iam_ctx* c, int id, const float* xy, int n) {
  int rc = bind(c);
  if (rc) return rc;
  return upload_keypoints_one(c, id, xy, n, true);
}

int iam_upload_keypoints_batch(iam_ctx* c, int n_images, const int32_t* ids, const float* const* xy, const int32_t* counts) {
  int rc = bind(c);
  if (rc) return rc;
  if (n_images < 0 || (n_images > 0 && (!ids || !xy || !counts))) return fail(IAM_E_ARG, "bad arguments");
  for (int i = 0; i < n_images; ++i)
    if ((rc = upload_keypoints_one(c, ids[i], xy[i], counts[i], false)) != IAM_OK) {
      cudaStreamSynchronize(c->up_stream);
      return rc;
    }
  CU(cudaStreamSynchronize(c->up_stream));   // one wait for the whole batch: the caller's arrays are free again
  return IAM_OK;
}

int iam_gms_filter(iam_ctx* c, const float* xy1, int n1, const float* xy2, int n2, const int32_t* matches, int n_matches,
                   int width_px, int height_px, int with_rotation, int with_scale, double threshold_factor,
                   uint8_t* out_mask) {
  int rc = bind(c);
  if (rc) return rc;
  if (n_matches < 0 || n1 < 0 || n2 < 0 || (n_matches > 0 && (!xy1 || !xy2 || !matches || !out_mask))) return fail(IAM_E_ARG, "bad arguments");
  if (n_matches > iam::kGmsMaxMatches) return fail(IAM_E_UNSUPPORTED, "the GMS filter handles at most %d matches (got %d)", iam::kGmsMaxMatches, n_matches);
  if (width_px <= 0 || height_px <= 0) return fail(IAM_E_ARG, "the GMS filter needs the image size");
  if (n_matches == 0) return IAM_OK;
  for (int m = 0; m < n_matches; ++m)
    if (matches[2 * m] < 0 || matches[2 * m] >= n1 || matches[2 * m + 1] < 0 || matches[2 * m + 1] >= n2)
      return fail(IAM_E_ARG, "match %d indexes outside the keypoint tables", m);
  // one self-contained job: [coordinates 1][coordinates 2][table][count][job][two image records]
  const size_t o_xy2 = size_t(n1) * 8, o_tab = o_xy2 + size_t(n2) * 8, o_cnt = o_tab + size_t(n_matches) * 8;
  const size_t o_job = (o_cnt + 4 + 15) / 16 * 16, o_img = o_job + sizeof(iam::RedJob);
  const size_t total = o_img + 2 * sizeof(iam::ImgDev);
  CU(cudaStreamSynchronize(c->stream));
  CU(c->packed_d.ensure(total + 16));
  uint8_t* d = c->packed_d.as<uint8_t>();
  std::vector<uint8_t> h(total, 0);
  if (n1) memcpy(h.data(), xy1, size_t(n1) * 8);
  if (n2) memcpy(h.data() + o_xy2, xy2, size_t(n2) * 8);
  memcpy(h.data() + o_tab, matches, size_t(n_matches) * 8);
  memcpy(h.data() + o_cnt, &n_matches, 4);
  iam::RedJob jb{0, n1, n2, 0, 1, {0, 0, 0}};
  memcpy(h.data() + o_job, &jb, sizeof jb);
  iam::ImgDev im[2] = {};
  im[0].kp_xy = reinterpret_cast<const float2*>(d);
  im[0].n = n1;
  im[1].kp_xy = reinterpret_cast<const float2*>(d + o_xy2);
  im[1].n = n2;
  memcpy(h.data() + o_img, im, sizeof im);
  CU(cudaMemcpyAsync(d, h.data(), total, cudaMemcpyHostToDevice, c->stream));
  iam::GmsParams gp{};
  gp.threshold_factor = threshold_factor;
  gp.width = width_px;
  gp.height = height_px;
  gp.with_rotation = with_rotation;
  gp.with_scale = with_scale & 1;
  gp.archive_wrap = (with_scale & 2) != 0;
  gp.gate_min_pairs = 0;
  cudaError_t e = iam::launch_gms(reinterpret_cast<const iam::RedJob*>(d + o_job), 1, reinterpret_cast<const iam::ImgDev*>(d + o_img), gp,
                                  n_matches, reinterpret_cast<int*>(d + o_tab), reinterpret_cast<int*>(d + o_cnt), c->stream);
  if (e != cudaSuccess) return fail(IAM_E_CUDA, "GMS launch: %s", cudaGetErrorString(e));
  c->timing.total_launches += 1;
  std::vector<int32_t> kept(size_t(n_matches) * 2 + 1);
  CU(cudaMemcpyAsync(kept.data(), d + o_tab, size_t(n_matches) * 8 + 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  // the kernel compacts in order: walk both lists to recover the mask
  const int n_kept = kept[size_t(n_matches) * 2];
  int o = 0;
  for (int m = 0; m < n_matches; ++m) {
    const bool k = o < n_kept && kept[2 * o] == matches[2 * m] && kept[2 * o + 1] == matches[2 * m + 1];
    out_mask[m] = k ? 1 : 0;
    if (k) ++o;
  }
  if (o != n_kept) return fail(IAM_E_CUDA, "GMS result is not an ordered subset of its input (%d of %d placed)", o, n_kept);
  return IAM_OK;
}

int iam_num_descriptors(iam_ctx* c, int id) {
  if (!c || id < 0 || id >= (int)c->images.size()) return -1;
  return c->images[id].n;
}

int iam_descriptors_exact(iam_ctx* c, int id) {
  if (!c || id < 0 || id >= (int)c->images.size() || c->images[id].n < 0) return -1;
  if (cudaSetDevice(c->device) != cudaSuccess) return -1;
  return resolve_exact(c, id);
}

int iam_knn_pairs(iam_ctx* c, const int32_t* pairs, int n_pairs, int k, int n_stride, int32_t* out_idx_fwd,
                  float* out_dist_fwd, int32_t* out_idx_rev, float* out_dist_rev) {
  int rc = bind(c);
  if (rc) return rc;
  if (n_pairs < 0 || (n_pairs && !pairs)) return fail(IAM_E_ARG, "bad pair list");
  if (k < 1 || k > 3) return fail(IAM_E_ARG, "k=%d unsupported (1..3; the reference uses 2 and 3)", k);
  if (!out_idx_fwd || !out_dist_fwd) return fail(IAM_E_ARG, "forward outputs are required");
  if ((out_idx_rev == nullptr) != (out_dist_rev == nullptr)) return fail(IAM_E_ARG, "reverse outputs must both be given or both be null");
  if (n_stride <= 0) return fail(IAM_E_ARG, "n_stride must be positive");
  const bool both = out_idx_rev != nullptr;
  if ((rc = prepare_plan(c, pairs, n_pairs, k, both)) != IAM_OK) return rc;
  const Plan& pl = c->plan;
  int engine;
  if ((rc = pick_engine(c, pairs, n_pairs, &engine)) != IAM_OK) return rc;
  int kind = 0;
  if (engine == IAM_ENGINE_UMMA && (rc = pick_kind(c, pairs, n_pairs, &kind)) != IAM_OK) return rc;
  if ((rc = sync_imgs(c)) != IAM_OK) return rc;
  CU(c->knn_idx.ensure(std::max<size_t>(1, pl.max_chunk_rows) * k * sizeof(int)));
  CU(c->knn_dist.ensure(std::max<size_t>(1, pl.max_chunk_rows) * k * sizeof(float)));
  c->timing.knn_launches = 0;
  c->timing.knn_ms = 0.f;
  c->timing.reduce_ms = 0.f;
  const int dirs = both ? 2 : 1;
  const int n_chunks = (int)pl.chunk_rows.size();
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int p0 = pl.chunk_pair_begin[ch], p1 = pl.chunk_pair_begin[ch + 1];
    const int u0 = pl.chunk_unit_begin[ch], u1 = pl.chunk_unit_begin[ch + 1];
    if (p1 == p0) continue;
    if ((rc = wait_uploads(c, pairs, p0, p1)) != IAM_OK) return rc;
    if (c->profiling) CU(cudaEventRecord(c->ev[0], c->stream));
    if ((rc = launch_knn(c, engine, kind, k, u0, u1 - u0)) != IAM_OK) return rc;
    cudaError_t e = iam::launch_finish_dist(c->norm, c->knn_dist.as<float>(), c->knn_idx.as<int>(), pl.chunk_rows[ch] * k, c->stream);
    if (e != cudaSuccess) return fail(IAM_E_CUDA, "finish launch: %s", cudaGetErrorString(e));
    c->timing.total_launches += 1;
    if (c->profiling) CU(cudaEventRecord(c->ev[1], c->stream));
    const size_t per = size_t(p1 - p0) * n_stride * k;
    CU(c->packed_i.ensure(per * sizeof(int)));
    CU(c->packed_d.ensure(per * sizeof(float)));
    for (int dir = 0; dir < dirs; ++dir) {
      pack_knn_kernel<<<p1 - p0, 256, 0, c->stream>>>(c->jobs.as<iam::RedJob>(), p0 * dirs, p1 - p0, dirs, dir, k, n_stride,
                                                      c->knn_idx.as<int>(), c->knn_dist.as<float>(), c->packed_i.as<int>(), c->packed_d.as<float>());
      CU(cudaGetLastError());
      c->timing.total_launches += 1;
      int32_t* oi = (dir ? out_idx_rev : out_idx_fwd) + size_t(p0) * n_stride * k;
      float* od = (dir ? out_dist_rev : out_dist_fwd) + size_t(p0) * n_stride * k;
      CU(cudaMemcpyAsync(oi, c->packed_i.p, per * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CU(cudaMemcpyAsync(od, c->packed_d.p, per * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
    }
    if (c->profiling) {
      float ms = 0.f;
      CU(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
      c->timing.knn_ms += ms;
    }
  }
  return mark_compute(c);
}

}  // extern "C"

namespace {
struct UploadFeed {  // host-side sources for iam_match_images: enqueue an image's upload right before its first use
  const void* const* ptrs = nullptr;
  const int32_t* const* keys = nullptr;
  const int32_t* ids = nullptr;
  int n_images = 0;
  int dtype = 0;
  std::vector<int> slot_of_id;   // image id -> index into ptrs
  std::vector<char> done;
  std::vector<std::pair<int, int>> waves_done;  // pair ranges whose done_ev was recorded, in order
  // float32 -> byte narrowing on the host (host_narrow.h): job index per slot (-1: none), in order of first use
  bool narrow = false;
  bool narrow_backward = false;  // the workers walk the images from the last one (few workers: NarrowPool::run)
  bool narrow_always = false;  // IAM_HOST_NARROW=2 (tests): never fall back to a float32 upload of a narrowable image
  iam::NarrowJob* jobs = nullptr;
  std::vector<int> job_of_slot;
  std::vector<cudaEvent_t> upload_waves;  // wave events of this call, in order (pacing of the enqueueing thread)
  // uint8 rows go straight to their final place and are converted wave by wave on the compute stream
  bool batch = false;
  int conv_used = 0;      // jobs of earlier waves in conv_jobs_h / conv_jobs_d
};

int match_core(iam_ctx* c, const int32_t* pairs, int n_pairs, const iam_match_params* prm, int waves, UploadFeed* feed,
               void** d_table, void** d_count);
}  // namespace

extern "C" {

int iam_match_pairs_device(iam_ctx* c, const int32_t* pairs, int n_pairs, const iam_match_params* prm, void** d_table,
                           void** d_count) {
  int rc = bind(c);
  if (rc) return rc;
  return match_core(c, pairs, n_pairs, prm, 0, nullptr, d_table, d_count);
}

int iam_match_images(iam_ctx* c, int n_images, const int32_t* image_ids, const void* const* host_ptrs,
                     const int32_t* counts, int dtype, const int32_t* const* key_ptrs, const int32_t* pairs, int n_pairs,
                     const iam_match_params* prm, int32_t* out_table, int32_t* out_count) {
  int rc = bind(c);
  if (rc) return rc;
  if (n_images < 0 || (n_images && (!image_ids || !host_ptrs || !counts)) || !out_table || !out_count)
    return fail(IAM_E_ARG, "bad arguments");
  if (dtype != IAM_DTYPE_U8 && dtype != IAM_DTYPE_F32) return fail(IAM_E_ARG, "unknown dtype %d", dtype);
  if (c->norm == IAM_NORM_HAMMING && dtype != IAM_DTYPE_U8) return fail(IAM_E_ARG, "Hamming descriptors must be uint8");
  if (n_pairs < 0 || (n_pairs && !pairs) || !prm) return fail(IAM_E_ARG, "bad arguments");
  if (key_ptrs)  // the keys index bitsets sized by the descriptor count (dedupe_kernel): same rule as iam_upload_keypoint_keys
    for (int i = 0; i < n_images; ++i)
      if (key_ptrs[i])
        for (int k = 0; k < counts[i]; ++k)
          if (key_ptrs[i][k] < 0 || key_ptrs[i][k] >= counts[i])
            return fail(IAM_E_ARG, "keypoint key %d of image %d out of range [0, %d)", key_ptrs[i][k], image_ids[i], counts[i]);
  UploadFeed feed;
  feed.ptrs = host_ptrs;
  feed.keys = key_ptrs;
  feed.ids = image_ids;
  feed.n_images = n_images;
  feed.dtype = dtype;
  feed.done.assign(n_images, 0);
  int max_id = -1;
  for (int i = 0; i < n_images; ++i) {
    if (image_ids[i] < 0 || image_ids[i] > (1 << 24)) return fail(IAM_E_ARG, "bad image id %d", image_ids[i]);
    if (counts[i] < 0 || (counts[i] > 0 && !host_ptrs[i])) return fail(IAM_E_ARG, "bad descriptor buffer for image %d", image_ids[i]);
    max_id = std::max(max_id, image_ids[i]);
  }
  feed.slot_of_id.assign(max_id + 1, -1);
  for (int i = 0; i < n_images; ++i) {
    feed.slot_of_id[image_ids[i]] = i;
    if ((rc = prepare_image(c, image_ids[i], counts[i])) != IAM_OK) return rc;
    c->images[image_ids[i]].seq = 0;  // contents are stale until this call uploads them
  }
  // L2: which operand forms this call's uploads build and its kernels read.  The byte layout (kind::i8) unless
  // the context already met descriptors it cannot hold, the caller forces fp16 operands, or a resident image
  // (not part of this call's uploads) has no usable byte layout.
  bool wide = c->norm != IAM_NORM_L2 || c->l2_wide_sticky || c->engine == IAM_ENGINE_UMMA_F16;
  if (c->norm == IAM_NORM_L2) {
    auto resident = [&](int id) { return id < 0 || id >= (int)feed.slot_of_id.size() || feed.slot_of_id[id] < 0; };
    for (int i = 0; i < 2 * n_pairs && !wide; ++i) {
      const int id = pairs[i];
      if (!resident(id)) continue;
      if ((rc = check_image(c, id)) != IAM_OK) return rc;
      Image& im = c->images[id];
      if (!im.has_i8 || resolve_flags(c, id) != 0 || im.i8ok != 1) wide = true;
    }
    if (wide)
      for (int i = 0; i < 2 * n_pairs; ++i)
        if (resident(pairs[i]) && check_image(c, pairs[i]) == IAM_OK && !c->images[pairs[i]].has_wide)
          return fail(IAM_E_STATE, "image %d holds only the byte layout but this call needs fp16 operands: upload it again", pairs[i]);
  }
  c->feed_wide = wide;
  // enough waves that the first kernels start after a few per cent of the bytes have crossed PCIe
  // (measured: 16 / 32 / 41 waves give 29.9 / 30.9 / 30.8 ms for 1990 pairs -- past 16 the matching itself is the
  // critical path, more waves only add launches)
  int max_waves = 16;
  if (const char* env = getenv("IAM_MAX_WAVES")) max_waves = std::max(1, atoi(env));  // A/B aid
  const int waves = std::max(1, std::min(max_waves, n_pairs / 96));
  const auto h0 = std::chrono::steady_clock::now();
  if (c->compute_pending) {
    CU(cudaStreamWaitEvent(c->up_stream, c->compute_done, 0));
    CU(cudaStreamWaitEvent(c->up_stream2, c->compute_done, 0));
    c->compute_pending = false;
  }
  CU(cudaEventRecord(c->span[0], c->up_stream));
  CU(cudaEventRecord(c->span[2], c->stream));
  // The matching kernel fills every SM it is given (registers, shared memory); leave a few SMs to the layout
  // conversion kernels of later waves so that uploads really overlap the matching of earlier waves.
  // float32 descriptors bound for the byte layout are integers in 0..255 (or the call is repeated on fp16 operands
  // anyway): worker threads narrow them to bytes while earlier waves upload and match (host_narrow.h)
  int narrow_mode = 1;  // 0: off, 1: adaptive (default), 2: always
  if (const char* env = getenv("IAM_HOST_NARROW")) narrow_mode = atoi(env);
  feed.narrow = narrow_mode != 0 && c->norm == IAM_NORM_L2 && !wide && dtype == IAM_DTYPE_F32;
  feed.narrow_always = narrow_mode == 2;
  if (feed.narrow && !c->narrow_pool_tried) {
    c->narrow_pool_tried = true;
    const int threads = iam::narrow_default_threads();
    if (threads > 0) {
      try {
        c->narrow_pool.reset(new iam::NarrowPool(threads));
      } catch (...) {  // no threads to be had: the float32 rows cross the bus as they are
        c->narrow_pool.reset();
      }
    }
  }
  if (feed.narrow && c->narrow_pool) {  // the arena of the previous call must have left the host
    CU(cudaStreamSynchronize(c->up_stream));
    CU(cudaStreamSynchronize(c->up_stream2));
  }
  c->h2d_bytes = 0;
  c->narrowed_images = 0;
  feed.narrow_backward = c->narrow_pool && c->narrow_pool->backward();
  c->feed_mode = true;
  // SMs kept free of the matching kernel for the per-image conversion kernels of later waves (the path float32 rows
  // and few-worker ranks take): 8, measured on the upload-bound strip (4 / 6 / 12 reserved: 24.2 / 22.9 / 21.9 ms against
  // 20.6) and the setting of the multi-GPU runs.  (A compute-bound single-GPU survey job ran 2.5 % faster with 4, but
  // the upload span grew from 104 to 164 ms -- not something an upload-bound 8-rank job can afford.)
  c->reserve_sms = waves > 1 ? 8 : 0;
  // Batched conversion: when every image is expected to arrive as uint8 rows (native uint8 input, or float32 with
  // enough narrowing workers to walk the list in order), the rows go straight to their final place and each wave is
  // converted by two launches on the compute stream -- no per-image conversion kernels next to the matching kernel,
  // so nothing is reserved for them.  IAM_BATCH_CONVERT=0 restores the per-image path (A/B aid).
  feed.batch = c->norm == IAM_NORM_L2 && !wide && (dtype == IAM_DTYPE_U8 || (feed.narrow && c->narrow_pool && !feed.narrow_backward));
  if (const char* env = getenv("IAM_BATCH_CONVERT")) feed.batch = feed.batch && atoi(env) != 0;
  if (feed.batch) {
    if (c->conv_jobs_cap < n_images) {
      CU(cudaStreamSynchronize(c->up_stream));
      if (c->conv_jobs_h) cudaFreeHost(c->conv_jobs_h);
      c->conv_jobs_h = nullptr;
      c->conv_jobs_cap = 0;
      if (cudaHostAlloc(reinterpret_cast<void**>(&c->conv_jobs_h), size_t(n_images) * sizeof(iam::ConvJob), cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        feed.batch = false;
      } else {
        c->conv_jobs_cap = n_images;
      }
    }
    if (feed.batch) CU(c->conv_jobs_d.ensure(size_t(std::max(1, n_images)) * sizeof(iam::ConvJob)));
    if (feed.batch) c->reserve_sms = 0;
  }
  if (const char* env = getenv("IAM_RESERVE_SMS")) c->reserve_sms = waves > 1 ? std::max(0, atoi(env)) : 0;  // A/B aid
  rc = match_core(c, pairs, n_pairs, prm, waves, &feed, nullptr, nullptr);
  c->feed_mode = false;
  c->reserve_sms = 0;
  if (c->narrow_pool && feed.jobs) {  // no worker may touch the job list (or this call's buffers) past this point
    for (int k = 0; k < feed.n_images && k < c->narrow_jobs_cap; ++k) {
      int expect = iam::NarrowJob::kFree;  // images nobody needed any more: not worth narrowing
      c->narrow_jobs[k].state.compare_exchange_strong(expect, iam::NarrowJob::kTaken, std::memory_order_acq_rel);
    }
    c->narrow_pool->finish();
  }
  if (rc != IAM_OK) return rc;
  for (int i = 0; i < n_images; ++i)  // images no pair referenced are still part of the resident set
    if (!feed.done[i] && (rc = enqueue_upload(c, image_ids[i], host_ptrs[i], true, dtype, key_ptrs ? key_ptrs[i] : nullptr)) != IAM_OK)
      return rc;
  CU(cudaEventRecord(c->lane2_ev, c->up_stream2));
  CU(cudaStreamWaitEvent(c->up_stream, c->lane2_ev, 0));
  CU(cudaEventRecord(c->span[1], c->up_stream));
  CU(cudaEventRecord(c->span[3], c->stream));
  c->timing.host_enqueue_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - h0).count();
  c->timing.waves = waves;
  c->timing.h2d_bytes = c->h2d_bytes;
  c->timing.narrowed_images = c->narrowed_images;
  c->span_pending = true;
  // Results travel back wave by wave on their own stream: only the last wave's rows are copied after the last
  // kernel.  (With a pageable destination each copy blocks this thread, not the GPU: everything is enqueued.)
  const size_t row = size_t(prm->cap) * 2;
  for (size_t w = 0; w < feed.waves_done.size(); ++w) {
    const int p0 = feed.waves_done[w].first, p1 = feed.waves_done[w].second;
    CU(cudaStreamWaitEvent(c->dl_stream, c->done_ev[w], 0));
    CU(cudaMemcpyAsync(out_count + p0, c->out_count.as<int>() + p0, size_t(p1 - p0) * sizeof(int), cudaMemcpyDeviceToHost, c->dl_stream));
    CU(cudaMemcpyAsync(out_table + size_t(p0) * row, c->out_table.as<int>() + size_t(p0) * row, size_t(p1 - p0) * row * sizeof(int),
                       cudaMemcpyDeviceToHost, c->dl_stream));
  }
  CU(cudaStreamSynchronize(c->dl_stream));
  CU(cudaStreamSynchronize(c->stream));
  if (c->norm == IAM_NORM_L2 && !wide) {
    // The uploads built the byte layout optimistically.  A conversion that met a non-integer component or a norm
    // beyond the layout's capacity cleared the context word: those results are void, the call is repeated on
    // fp16 operands and the context stays there (SURF / RootSIFT projects pay this once).
    int ok = 1;
    CU(cudaMemcpy(&ok, c->d_ctx_flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (!ok) {
      c->l2_wide_sticky = true;
      return iam_match_images(c, n_images, image_ids, host_ptrs, counts, dtype, key_ptrs, pairs, n_pairs, prm, out_table, out_count);
    }
  }
  return IAM_OK;
}

}  // extern "C"

namespace {

// Hand every float32 image of this call to the narrowing workers, in the order the waves will need them.
int start_narrowing(iam_ctx* c, const Plan& pl, const int32_t* pairs, UploadFeed* feed) {
  feed->job_of_slot.assign(feed->n_images, -1);
  if (!c->narrow_pool) {
    feed->narrow = false;
    return IAM_OK;
  }
  std::vector<int> order;
  order.reserve(feed->n_images);
  size_t bytes = 0;
  for (size_t ch = 0; ch + 1 < pl.chunk_pair_begin.size(); ++ch)
    for (int i = 2 * pl.chunk_pair_begin[ch]; i < 2 * pl.chunk_pair_begin[ch + 1]; ++i) {
      const int id = pairs[i];
      const int slot = id < (int)feed->slot_of_id.size() ? feed->slot_of_id[id] : -1;
      if (slot < 0 || feed->job_of_slot[slot] >= 0 || c->images[id].n <= 0) continue;
      feed->job_of_slot[slot] = (int)order.size();
      order.push_back(slot);
      bytes += size_t(c->images[id].n) * c->desc_bytes;
    }
  if (order.empty()) {
    feed->narrow = false;
    return IAM_OK;
  }
  if (c->narrow_cap < bytes) {  // page-locked: the H2D copies out of it are asynchronous
    CU(cudaStreamSynchronize(c->up_stream));
    CU(cudaStreamSynchronize(c->up_stream2));
    if (c->narrow_arena) CU(cudaFreeHost(c->narrow_arena));
    c->narrow_arena = nullptr;
    c->narrow_cap = 0;
    if (cudaHostAlloc(reinterpret_cast<void**>(&c->narrow_arena), bytes + bytes / 8, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();  // no page-locked memory for the arena: plain float32 uploads, not an error
      c->narrow_arena = nullptr;
      feed->narrow = false;
      feed->job_of_slot.assign(feed->n_images, -1);
      return IAM_OK;
    }
    c->narrow_cap = bytes + bytes / 8;
  }
  if (c->narrow_jobs_cap < (int)order.size()) {
    c->narrow_jobs.reset(new iam::NarrowJob[order.size()]);
    c->narrow_jobs_cap = (int)order.size();
  }
  // Which images the workers take: all of them, unless IAM_NARROW_FRACTION=x (A/B aid) sends a share 1 - x as float32,
  // spread evenly over the order of first use.  Measured on B200 (16 host threads): the 500-frame strip (4 pairs per
  // frame, PCIe-bound without narrowing) does 99.9 k pairs/s end to end with every image narrowed and 95.0 / 89.1 /
  // 77.5 k with 85 / 70 / 50 % -- float32 rows also cost a staging copy and a 2.5 x longer conversion kernel next to
  // the matching kernel; the 2812-frame survey (15 pairs per frame, compute-bound) does 116.8 k either way.  When the
  // workers cannot keep up (few host threads per rank) the enqueueing thread claims images for float32 upload as the
  // bus runs dry (match_core), which splits the list between bus and workers where they meet.
  double frac = 1.0;
  if (const char* env = getenv("IAM_NARROW_FRACTION")) frac = std::min(1.0, std::max(0.0, atof(env)));
  if (feed->narrow_always) frac = 1.0;
  size_t off = 0;
  double owed = 0.0;   // float32 uploads owed so far (error diffusion over the order of first use)
  for (size_t k = 0; k < order.size(); ++k) {
    const int slot = order[k];
    const Image& im = c->images[feed->ids[slot]];
    iam::NarrowJob& job = c->narrow_jobs[k];
    job.src = static_cast<const float*>(feed->ptrs[slot]);
    job.dst = c->narrow_arena + off;
    job.n = size_t(im.n) * c->desc_bytes;
    owed += 1.0 - frac;
    const bool as_f32 = owed >= 1.0 - 1e-9;
    if (as_f32) owed -= 1.0;
    job.state.store(as_f32 ? iam::NarrowJob::kTaken : iam::NarrowJob::kFree, std::memory_order_relaxed);
    off += job.n;
  }
  feed->jobs = c->narrow_jobs.get();
  c->narrow_pool->start(feed->jobs, (int)order.size());
  return IAM_OK;
}

int match_core(iam_ctx* c, const int32_t* pairs, int n_pairs, const iam_match_params* prm, int waves, UploadFeed* feed,
               void** d_table, void** d_count) {
  int rc;
  if (n_pairs < 0 || (n_pairs && !pairs) || !prm) return fail(IAM_E_ARG, "bad arguments");
  if (prm->cap <= 0 || prm->cap > 65536) return fail(IAM_E_ARG, "cap=%d out of range", prm->cap);
  if (prm->reduce_mode != IAM_REDUCE_LOWE && prm->reduce_mode != IAM_REDUCE_REF_METRIC) return fail(IAM_E_ARG, "unknown reduce mode %d", prm->reduce_mode);
  const int k = 2;
  if (prm->gms) {
    if (prm->cap > iam::kGmsMaxMatches) return fail(IAM_E_UNSUPPORTED, "the GMS filter handles at most %d matches per direction (cap=%d)", iam::kGmsMaxMatches, prm->cap);
    if (prm->width_px <= 0 || prm->height_px <= 0) return fail(IAM_E_ARG, "the GMS filter needs the image size (width_px, height_px)");
    for (int i = 0; i < 2 * n_pairs; ++i) {
      const int id = pairs[i];
      if (id < 0 || id >= (int)c->images.size() || c->images[id].n < 0) return fail(IAM_E_STATE, "image %d has no descriptors uploaded", id);
      if (c->images[id].n > 0 && (!c->images[id].kp || c->images[id].kp_n != c->images[id].n))
        return fail(IAM_E_STATE, "the GMS filter needs iam_upload_keypoints for image %d", id);
    }
  }
  if ((rc = prepare_plan(c, pairs, n_pairs, k, true, waves)) != IAM_OK) return rc;
  const Plan& pl = c->plan;
  int engine;
  if ((rc = pick_engine(c, pairs, n_pairs, &engine)) != IAM_OK) return rc;
  int kind = 0;
  if (engine == IAM_ENGINE_UMMA && (rc = pick_kind(c, pairs, n_pairs, &kind)) != IAM_OK) return rc;
  if (feed && feed->keys) {  // device key pointers must be in the image table before it is made resident
    for (int i = 0; i < feed->n_images; ++i) {
      Image& im = c->images[feed->ids[i]];
      if (im.keys) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaFree(im.keys));
        im.keys = nullptr;
      }
      if (feed->keys[i] && im.n > 0) CU(cudaMalloc(reinterpret_cast<void**>(&im.keys), size_t(im.n) * sizeof(int)));
      im.dev.kp_key = im.keys;
      c->imgs_dirty = true;
    }
  }
  if ((rc = sync_imgs(c)) != IAM_OK) return rc;
  if (feed && feed->narrow && (rc = start_narrowing(c, pl, pairs, feed)) != IAM_OK) return rc;

  const size_t cap = prm->cap;
  int max_chunk_pairs = 1;
  for (size_t ch = 0; ch + 1 < pl.chunk_pair_begin.size(); ++ch)
    max_chunk_pairs = std::max(max_chunk_pairs, pl.chunk_pair_begin[ch + 1] - pl.chunk_pair_begin[ch]);
  const int cand_stride = std::max(1, pl.max_n);
  CU(c->knn_idx.ensure(std::max<size_t>(1, pl.max_chunk_rows) * k * sizeof(int)));
  CU(c->knn_dist.ensure(std::max<size_t>(1, pl.max_chunk_rows) * k * sizeof(float)));
  CU(c->cand_metric.ensure(size_t(max_chunk_pairs) * 2 * cand_stride * sizeof(double)));
  CU(c->cand_qt.ensure(size_t(max_chunk_pairs) * 2 * cand_stride * sizeof(int2)));
  CU(c->job_table.ensure(size_t(max_chunk_pairs) * 2 * cap * 2 * sizeof(int)));
  CU(c->job_count.ensure(size_t(max_chunk_pairs) * 2 * sizeof(int)));
  CU(c->out_table.ensure(std::max<size_t>(1, size_t(n_pairs)) * cap * 2 * sizeof(int)));
  CU(c->out_count.ensure(std::max<size_t>(1, size_t(n_pairs)) * sizeof(int)));

  iam::ReduceParams rp{};
  rp.ratio = prm->match_ratio;
  rp.thresh = prm->max_distance * prm->match_ratio;  // double product, as Python evaluates matcher.py:261
  rp.mode = prm->reduce_mode;
  rp.cap = prm->cap;
  rp.min_pairs = prm->min_pairs;

  c->timing.knn_launches = 0;
  c->timing.knn_ms = 0.f;
  c->timing.reduce_ms = 0.f;
  const int n_chunks = (int)pl.chunk_rows.size();
  int n_prof = 0;
  std::vector<std::pair<int, int>> ran;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int p0 = pl.chunk_pair_begin[ch], p1 = pl.chunk_pair_begin[ch + 1];
    const int u0 = pl.chunk_unit_begin[ch], u1 = pl.chunk_unit_begin[ch + 1];
    if (p1 == p0) continue;
    if (feed) {  // enqueue the uploads this chunk is the first to need (upload stream; overlaps earlier chunks' kernels)
      bool any_upload = false;
      int n_conv = 0, max_conv_pad = 0;   // images of this wave left to the batched conversion
      // With narrowing the enqueueing thread paces itself on the upload stream (at most two waves ahead): an image
      // is claimed for a float32 upload only when PCIe is about to run dry, which gives the workers time to get ahead.
      if (feed->narrow && feed->narrow_backward && !feed->narrow_always && feed->upload_waves.size() >= 2) CU(cudaEventSynchronize(feed->upload_waves[feed->upload_waves.size() - 2]));
      for (int i = 2 * p0; i < 2 * p1; ++i) {
        const int id = pairs[i];
        const int slot = id < (int)feed->slot_of_id.size() ? feed->slot_of_id[id] : -1;
        if (slot >= 0 && !feed->done[slot]) {
          feed->done[slot] = 1;
          Image& im = c->images[id];
          const int32_t* hk = feed->keys ? feed->keys[slot] : nullptr;
          int* keep = im.keys;  // allocated above; enqueue_upload must reuse it, not free it
          im.keys = nullptr;
          const void* usrc = feed->ptrs[slot];
          int udtype = feed->dtype;
          if (feed->narrow && feed->job_of_slot[slot] >= 0) {
            // Bytes when a worker has narrowed (or is narrowing) this image; the float32 rows when no worker has
            // reached it yet -- PCIe then carries them while the workers go on with later images.
            iam::NarrowJob& job = feed->jobs[feed->job_of_slot[slot]];
            int st = job.state.load(std::memory_order_acquire);
            if (st == iam::NarrowJob::kFree && (feed->narrow_always || feed->narrow_backward)) {
              int expect = iam::NarrowJob::kFree;
              const int claim = feed->narrow_always ? iam::NarrowJob::kBusy : iam::NarrowJob::kTaken;
              st = job.state.compare_exchange_strong(expect, claim, std::memory_order_acq_rel) ? claim : expect;
              if (st == iam::NarrowJob::kBusy && feed->narrow_always) {  // IAM_HOST_NARROW=2: this thread does it itself
                st = iam::narrow_f32_to_u8(job.src, job.dst, job.n) ? iam::NarrowJob::kBad : iam::NarrowJob::kDone;
                job.state.store(st, std::memory_order_release);
              }
            }
            while (st == iam::NarrowJob::kBusy || st == iam::NarrowJob::kFree) {
              // kFree is only still seen in forward order (the workers are about to reach this image): wait for
              // them while the bus has work queued, send the float32 rows as soon as it would run dry
              if (st == iam::NarrowJob::kFree && !feed->batch && cudaEventQuery(c->probe_ev[0]) == cudaSuccess &&
                  cudaEventQuery(c->probe_ev[1]) == cudaSuccess) {
                int expect = iam::NarrowJob::kFree;
                if (job.state.compare_exchange_strong(expect, iam::NarrowJob::kTaken, std::memory_order_acq_rel)) break;
              }
              std::this_thread::yield();
              st = job.state.load(std::memory_order_acquire);
            }
            if (st == iam::NarrowJob::kDone) {
              usrc = job.dst;
              udtype = IAM_DTYPE_U8;
              c->narrowed_images++;
            }
          }
          if (feed->batch && udtype == IAM_DTYPE_U8 && feed->conv_used + n_conv < c->conv_jobs_cap) {
            if ((rc = enqueue_copy_u8(c, id, usrc, &c->conv_jobs_h[feed->conv_used + n_conv])) != IAM_OK) return rc;
            max_conv_pad = std::max(max_conv_pad, im.n_pad);
            ++n_conv;
          } else if ((rc = enqueue_upload(c, id, usrc, true, udtype, nullptr)) != IAM_OK) {
            return rc;
          }
          im.keys = keep;
          im.dev.kp_key = keep;
          if (hk && keep) CU(cudaMemcpyAsync(keep, hk, size_t(im.n) * sizeof(int), cudaMemcpyHostToDevice, c->up_stream));
          any_upload = true;
        }
      }
      if (any_upload) {  // the compute stream is in order: waiting for this wave's uploads covers all earlier ones
        if ((int)c->wave_ev.size() <= ch) {
          const size_t old = c->wave_ev.size();
          c->wave_ev.resize(ch + 1, nullptr);
          for (size_t w = old; w < c->wave_ev.size(); ++w) CU(cudaEventCreateWithFlags(&c->wave_ev[w], cudaEventDisableTiming));
        }
        CU(cudaEventRecord(c->lane2_ev, c->up_stream2));          // fold lane 2 into lane 1, then one event for both
        CU(cudaStreamWaitEvent(c->up_stream, c->lane2_ev, 0));
        iam::ConvJob* d_jobs = c->conv_jobs_d.as<iam::ConvJob>() + feed->conv_used;
        if (n_conv > 0)   // the wave's conversion jobs ride behind its rows
          CU(cudaMemcpyAsync(d_jobs, c->conv_jobs_h + feed->conv_used, size_t(n_conv) * sizeof(iam::ConvJob), cudaMemcpyHostToDevice, c->up_stream));
        CU(cudaEventRecord(c->wave_ev[ch], c->up_stream));
        CU(cudaStreamWaitEvent(c->stream, c->wave_ev[ch], 0));
        feed->upload_waves.push_back(c->wave_ev[ch]);
        if (n_conv > 0) {  // two launches convert the whole wave, on the stream (and all the SMs) of its matching kernel
          cudaError_t ce = iam::launch_convert_u8_batch(d_jobs, n_conv, max_conv_pad, c->desc_bytes, c->d_ctx_flag, c->stream);
          if (ce != cudaSuccess) return fail(IAM_E_CUDA, "batched convert launch: %s", cudaGetErrorString(ce));
          c->timing.total_launches += 2;
          feed->conv_used += n_conv;
        }
      }
    } else if ((rc = wait_uploads(c, pairs, p0, p1)) != IAM_OK) {
      return rc;
    }
    const bool prof = c->profiling && !feed;  // per-chunk event triples, summed in iam_get_timing
    cudaEvent_t* pev = nullptr;
    if (prof) {
      while (c->chunk_ev.size() < size_t(3 * (n_prof + 1))) {
        cudaEvent_t e = nullptr;
        CU(cudaEventCreate(&e));
        c->chunk_ev.push_back(e);
      }
      pev = &c->chunk_ev[3 * n_prof++];
    }
    if (prof) CU(cudaEventRecord(pev[0], c->stream));
    if ((rc = launch_knn(c, engine, kind, k, u0, u1 - u0)) != IAM_OK) return rc;
    if (prof) CU(cudaEventRecord(pev[1], c->stream));
    // the finishing pass over the lists (sqrt, padding rows: launch_finish_dist in iam_knn_pairs) is folded into the reduction
    const iam::RedJob* jobs = c->jobs.as<iam::RedJob>() + size_t(p0) * 2;
    cudaError_t e = iam::launch_reduce(jobs, (p1 - p0) * 2, c->knn_idx.as<int>(), c->knn_dist.as<float>(), k, rp, c->cand_metric.as<double>(),
                           c->cand_qt.as<int2>(), cand_stride, c->job_table.as<int>(), c->job_count.as<int>(),
                           c->norm == IAM_NORM_L2 ? 1 : 2, c->stream);
    if (e != cudaSuccess) return fail(IAM_E_CUDA, "reduce launch: %s", cudaGetErrorString(e));
    if (prm->gms) {  // matcher.py:285, between the metric reduction and filter_duplicates
      iam::GmsParams gp{};
      gp.threshold_factor = prm->gms_threshold;
      gp.width = prm->width_px;
      gp.height = prm->height_px;
      gp.with_rotation = prm->gms_rotation;
      gp.with_scale = prm->gms_scale;
      gp.archive_wrap = prm->gms == 2;
      gp.gate_min_pairs = prm->dedupe ? 0 : prm->min_pairs;  // the gate of matcher.py:296-298 follows filter_duplicates
      e = iam::launch_gms(jobs, (p1 - p0) * 2, c->d_imgs.as<iam::ImgDev>(), gp, prm->cap, c->job_table.as<int>(), c->job_count.as<int>(), c->stream);
      if (e != cudaSuccess) return fail(IAM_E_CUDA, "GMS launch: %s", cudaGetErrorString(e));
      c->timing.total_launches += 1;
    }
    if (prm->dedupe) {
      e = iam::launch_dedupe(jobs, (p1 - p0) * 2, c->d_imgs.as<iam::ImgDev>(), prm->cap, prm->min_pairs, pl.max_n,
                             c->job_table.as<int>(), c->job_count.as<int>(), c->stream);
      if (e != cudaSuccess) return fail(IAM_E_CUDA, "dedupe launch: %s", cudaGetErrorString(e));
      c->timing.total_launches += 1;
    }
    e = iam::launch_crosscheck(jobs, p1 - p0, c->job_table.as<int>(), c->job_count.as<int>(), prm->cap, prm->cross_check, pl.max_n,
                               c->out_table.as<int>() + size_t(p0) * cap * 2, c->out_count.as<int>() + p0, c->stream);
    if (e != cudaSuccess) return fail(IAM_E_CUDA, "cross-check launch: %s", cudaGetErrorString(e));
    c->timing.total_launches += 2;
    if (prof) CU(cudaEventRecord(pev[2], c->stream));
    if (feed) {
      const size_t w = feed->waves_done.size();
      if (c->done_ev.size() <= w) {
        c->done_ev.resize(w + 1, nullptr);
        CU(cudaEventCreateWithFlags(&c->done_ev[w], cudaEventDisableTiming));
      }
      CU(cudaEventRecord(c->done_ev[w], c->stream));
      feed->waves_done.emplace_back(p0, p1);
    }
  }
  c->timing_pending = n_prof > 0;  // resolved lazily in iam_get_timing
  c->n_prof_chunks = n_prof;
  if ((rc = mark_compute(c)) != IAM_OK) return rc;
  c->last_pairs = n_pairs;
  c->last_cap = prm->cap;
  if (d_table) *d_table = c->out_table.p;
  if (d_count) *d_count = c->out_count.p;
  return IAM_OK;
}

}  // namespace

extern "C" {

int iam_fetch_tables(iam_ctx* c, int32_t* out_table, int32_t* out_count) {
  int rc = bind(c);
  if (rc) return rc;
  if (!out_table || !out_count) return fail(IAM_E_ARG, "null output");
  if (c->last_pairs > 0) {
    CU(cudaMemcpyAsync(out_count, c->out_count.p, size_t(c->last_pairs) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(out_table, c->out_table.p, size_t(c->last_pairs) * c->last_cap * 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  return IAM_OK;
}

int iam_pack_tables_device(iam_ctx* c, void** d_rows, void** d_offsets, long long* total) {
  int rc = bind(c);
  if (rc) return rc;
  if (!d_rows || !d_offsets || !total) return fail(IAM_E_ARG, "null output");
  const int n = c->last_pairs;
  CU(c->csr_off.ensure(size_t(n + 1) * sizeof(int)));
  cudaError_t e = iam::launch_scan_counts(c->out_count.as<int>(), n, c->csr_off.as<int>(), c->stream);
  if (e != cudaSuccess) return fail(IAM_E_CUDA, "scan launch: %s", cudaGetErrorString(e));
  int tot = 0;
  CU(cudaMemcpyAsync(&tot, c->csr_off.as<int>() + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));  // the payload's size decides the allocation (and the caller's collective)
  CU(c->csr_rows.ensure(std::max<size_t>(1, size_t(tot)) * 2 * sizeof(int)));
  e = iam::launch_pack_tables(c->out_table.as<int>(), c->out_count.as<int>(), c->csr_off.as<int>(), n, c->last_cap,
                              c->csr_rows.as<int>(), c->stream);
  if (e != cudaSuccess) return fail(IAM_E_CUDA, "pack launch: %s", cudaGetErrorString(e));
  c->timing.total_launches += n > 0 ? 2 : 1;
  *d_rows = c->csr_rows.p;
  *d_offsets = c->csr_off.p;
  *total = tot;
  return IAM_OK;
}

int iam_fetch_packed_tables(iam_ctx* c, int32_t* out_rows, long long cap_rows, int32_t* out_offsets, long long* total) {
  if (!out_offsets || !total || cap_rows < 0 || (cap_rows > 0 && !out_rows)) return fail(IAM_E_ARG, "bad arguments");
  void *d_rows = nullptr, *d_off = nullptr;
  int rc = iam_pack_tables_device(c, &d_rows, &d_off, total);
  if (rc) return rc;
  if (*total > cap_rows) return fail(IAM_E_UNSUPPORTED, "%lld match rows, the caller's array holds %lld", *total, cap_rows);
  if (*total > 0) CU(cudaMemcpyAsync(out_rows, d_rows, size_t(*total) * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(out_offsets, d_off, size_t(c->last_pairs + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return IAM_OK;
}

int iam_match_pairs(iam_ctx* c, const int32_t* pairs, int n_pairs, const iam_match_params* prm, int32_t* out_table,
                    int32_t* out_count, int32_t* out_table_rev, int32_t* out_count_rev) {
  if (!out_table || !out_count) return fail(IAM_E_ARG, "null output");
  if ((out_table_rev == nullptr) != (out_count_rev == nullptr)) return fail(IAM_E_ARG, "reverse outputs must both be given or both be null");
  int rc = iam_match_pairs_device(c, pairs, n_pairs, prm, nullptr, nullptr);
  if (rc) return rc;
  rc = iam_fetch_tables(c, out_table, out_count);
  if (rc) return rc;
  if (out_table_rev) {
    // independent reverse direction = the same pipeline on the swapped pair list
    std::vector<int32_t> sw(size_t(n_pairs) * 2);
    for (int p = 0; p < n_pairs; ++p) {
      sw[2 * p] = pairs[2 * p + 1];
      sw[2 * p + 1] = pairs[2 * p];
    }
    rc = iam_match_pairs_device(c, sw.data(), n_pairs, prm, nullptr, nullptr);
    if (rc) return rc;
    rc = iam_fetch_tables(c, out_table_rev, out_count_rev);
  }
  return rc;
}

int iam_debug_tile(iam_ctx* c, int q_id, int t_id, int q_tile, int t_tile, uint32_t lbo, uint32_t sbo,
                   uint32_t kstep_bytes, int ksteps, float* out_host) {
  int rc = bind(c);
  if (rc) return rc;
  if ((rc = check_image(c, q_id)) != IAM_OK || (rc = check_image(c, t_id)) != IAM_OK) return rc;
  if (!out_host) return fail(IAM_E_ARG, "null output");
  const Image& q = c->images[q_id];
  const Image& t = c->images[t_id];
  if (q_tile < 0 || t_tile < 0 || (q_tile + 1) * iam::kTileRows > q.n_pad || (t_tile + 1) * iam::kTileRows > t.n_pad)
    return fail(IAM_E_ARG, "tile index out of range");
  CU(c->packed_d.ensure(128 * 128 * sizeof(float)));
  cudaError_t e;
  if (lbo == 0) {  // byte layout (kind::i8): the layout's own strides; only the sign of `ksteps` is used
    if (!q.has_i8 || !t.has_i8) return fail(IAM_E_STATE, "no byte layout resident");
    const size_t tb = iam::LayD<iam::Kind::I8>::kTileBytes;
    e = iam::launch_umma_tile_debug(iam::kKindI8, q.dev.i8_form + size_t(q_tile) * tb, t.dev.i8_form + size_t(t_tile) * tb, 0, 0,
                                    0, ksteps, c->packed_d.as<float>(), c->stream);
  } else {
    if (!q.has_wide || !t.has_wide) return fail(IAM_E_STATE, "no wide operand forms resident");
    e = iam::launch_umma_tile_debug(c->norm == IAM_NORM_L2 ? iam::kKindF16 : iam::kKindF8,
                                    q.dev.a_form + size_t(q_tile) * iam::kTileBytes,
                                    t.dev.b_form + size_t(t_tile) * iam::kTileBytes, lbo, sbo, kstep_bytes, ksteps,
                                    c->packed_d.as<float>(), c->stream);
  }
  if (e != cudaSuccess) return fail(IAM_E_CUDA, "debug tile launch: %s", cudaGetErrorString(e));
  CU(cudaMemcpyAsync(out_host, c->packed_d.p, 128 * 128 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return IAM_OK;
}

int iam_debug_narrow(const float* src, uint8_t* dst, size_t n) {
  if (n && (!src || !dst)) return fail(IAM_E_ARG, "null buffer");
  return iam::narrow_f32_to_u8(src, dst, n);
}

int iam_debug_minimal_solver(int model, const float* x1, const float* y1, const float* x2, const float* y2,
                              float* out_models) {
  if (!x1 || !y1 || !x2 || !y2 || !out_models) return fail(IAM_E_ARG, "null argument");
  return iam::debug_minimal_solver(model, x1, y1, x2, y2, out_models);
}

int iam_ransac_pairs(iam_ctx* c, int model, const float* pts1, const float* pts2, const int32_t* off, int n_pairs,
                     const double* K, double threshold_px, double prob, int max_iters, uint32_t seed, uint8_t* out_mask,
                     double* out_model, int32_t* out_inliers) {
  int rc = bind(c);
  if (rc) return rc;
  if (n_pairs < 0 || !off || !out_mask || !out_model || !out_inliers) return fail(IAM_E_ARG, "bad arguments");
  if (model != IAM_MODEL_ESSENTIAL && model != IAM_MODEL_HOMOGRAPHY && model != IAM_MODEL_FUNDAMENTAL && model != IAM_MODEL_AFFINE_PARTIAL)
    return fail(IAM_E_ARG, "unknown model %d", model);
  if (model == IAM_MODEL_ESSENTIAL && !K) return fail(IAM_E_ARG, "K is required for the essential-matrix model");
  std::string err;
  rc = iam::ransac_pairs(model, pts1, pts2, off, n_pairs, K, threshold_px, prob, max_iters, seed, out_mask, out_model,
                         out_inliers, &c->ransac_scratch, c->stream, &err);
  if (rc != 0) return fail(rc, "%s", err.c_str());
  c->timing.total_launches += 1;
  return IAM_OK;
}

int iam_ransac_tables(iam_ctx* c, int model, const double* K, double threshold_px, double prob, int max_iters,
                      uint32_t seed, int min_pairs, int compact, uint8_t* out_mask, double* out_model,
                      int32_t* out_inliers) {
  int rc = bind(c);
  if (rc) return rc;
  if (model != IAM_MODEL_ESSENTIAL && model != IAM_MODEL_HOMOGRAPHY && model != IAM_MODEL_FUNDAMENTAL && model != IAM_MODEL_AFFINE_PARTIAL)
    return fail(IAM_E_ARG, "unknown model %d", model);
  if (model == IAM_MODEL_ESSENTIAL && !K) return fail(IAM_E_ARG, "K is required for the essential-matrix model");
  const int n = c->last_pairs, cap = c->last_cap;
  if (n <= 0) return IAM_OK;
  if (!c->plan_valid || (int)c->plan.jobs.size() != 2 * n) return fail(IAM_E_STATE, "no match tables resident: call iam_match_pairs_device / iam_match_images first");
  for (size_t j = 0; j < c->plan.jobs.size(); j += 2) {
    const Image& a = c->images[c->plan.jobs[j].q_slot];
    const Image& b = c->images[c->plan.jobs[j].t_slot];
    if ((a.n > 0 && (!a.kp || a.kp_n != a.n)) || (b.n > 0 && (!b.kp || b.kp_n != b.n)))
      return fail(IAM_E_STATE, "iam_ransac_tables needs iam_upload_keypoints for images %d and %d", c->plan.jobs[j].q_slot, c->plan.jobs[j].t_slot);
  }
  if ((rc = sync_imgs(c)) != IAM_OK) return rc;
  if (out_mask) CU(c->ransac_mask.ensure(size_t(n) * cap));
  CU(c->ransac_model.ensure(size_t(n) * 9 * sizeof(float)));
  CU(c->ransac_inl.ensure(size_t(n) * sizeof(int)));
  std::string err;
  rc = iam::ransac_tables(model, c->out_table.as<int>(), c->out_count.as<int>(), cap, n, c->jobs.as<iam::RedJob>(),
                          c->d_imgs.as<iam::ImgDev>(), K, threshold_px, prob, max_iters, seed, min_pairs, compact != 0,
                          out_mask ? c->ransac_mask.as<uint8_t>() : nullptr, c->ransac_model.as<float>(),
                          c->ransac_inl.as<int>(), c->stream, &err);
  if (rc != 0) return fail(rc, "%s", err.c_str());
  c->timing.total_launches += 1;
  if (out_mask) CU(cudaMemcpyAsync(out_mask, c->ransac_mask.p, size_t(n) * cap, cudaMemcpyDeviceToHost, c->stream));
  std::vector<float> hm;
  if (out_model) {
    hm.resize(size_t(n) * 9);
    CU(cudaMemcpyAsync(hm.data(), c->ransac_model.p, hm.size() * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  }
  if (out_inliers) CU(cudaMemcpyAsync(out_inliers, c->ransac_inl.p, size_t(n) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (out_mask || out_model || out_inliers) CU(cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i < hm.size(); ++i) out_model[i] = hm[i];
  return IAM_OK;
}


// ---- two-view triangulation (smart.py:26-63, :116-131) ---------------------------------------------------------

int iam_triangulate_pairs(iam_ctx* c, int n_pairs, const double* proj1, const double* proj2, const int32_t* off,
                          const double* x1, const double* x2, double* out_points, double* out_stats) {
  int rc = bind(c);
  if (rc) return rc;
  if (n_pairs < 0 || (n_pairs > 0 && (!proj1 || !proj2 || !off || !out_stats))) return fail(IAM_E_ARG, "bad arguments");
  if (n_pairs == 0) return IAM_OK;
  if (off[0] != 0) return fail(IAM_E_ARG, "off[0] must be 0");
  for (int p = 0; p < n_pairs; ++p)
    if (off[p + 1] < off[p]) return fail(IAM_E_ARG, "offsets must not decrease (pair %d)", p);
  const size_t total = size_t(off[n_pairs]);
  if (total > 0 && (!x1 || !x2)) return fail(IAM_E_ARG, "bad arguments");
  // one staging block: proj1 | proj2 | x1 | x2 | off
  const size_t b_proj = size_t(n_pairs) * 12 * sizeof(double), b_x = total * 2 * sizeof(double);
  const size_t b_off = (size_t(n_pairs) + 1) * sizeof(int32_t);
  CU(c->tri_in.ensure(2 * b_proj + 2 * b_x + b_off + 64));
  CU(c->tri_out.ensure((total * 3 + size_t(n_pairs) * 2) * sizeof(double) + 64));
  uint8_t* in = c->tri_in.as<uint8_t>();
  double* d_p1 = reinterpret_cast<double*>(in);
  double* d_p2 = reinterpret_cast<double*>(in + b_proj);
  double* d_x1 = reinterpret_cast<double*>(in + 2 * b_proj);
  double* d_x2 = reinterpret_cast<double*>(in + 2 * b_proj + b_x);
  int32_t* d_off = reinterpret_cast<int32_t*>(in + 2 * b_proj + 2 * b_x);
  double* d_pts = c->tri_out.as<double>();
  double* d_stats = d_pts + total * 3;
  CU(cudaMemcpyAsync(d_p1, proj1, b_proj, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(d_p2, proj2, b_proj, cudaMemcpyHostToDevice, c->stream));
  if (total) {
    CU(cudaMemcpyAsync(d_x1, x1, b_x, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_x2, x2, b_x, cudaMemcpyHostToDevice, c->stream));
  }
  CU(cudaMemcpyAsync(d_off, off, b_off, cudaMemcpyHostToDevice, c->stream));
  CU(iam::triangulate_pairs(n_pairs, d_p1, d_p2, d_off, d_x1, d_x2, d_pts, d_stats, c->stream));
  c->timing.total_launches += 1;
  if (out_points && total) CU(cudaMemcpyAsync(out_points, d_pts, total * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(out_stats, d_stats, size_t(n_pairs) * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return IAM_OK;
}

// ---- ORB detect + describe (image.py:243-245, :324) -------------------------------------------------------------

int iam_orb_detect(iam_ctx* c, const uint8_t* gray, int width, int height, int nfeatures, int max_out, float* out_kp,
                   uint8_t* out_des, int* out_n) {
  int rc = bind(c);
  if (rc) return rc;
  if (!gray || !out_kp || !out_des || !out_n || max_out < 0) return fail(IAM_E_ARG, "bad arguments");
  if (width <= 0 || height <= 0 || width > 65535 || height > 65535) return fail(IAM_E_ARG, "image size %d x %d out of range", width, height);
  std::string err;
  rc = iam::orb_detect(gray, width, height, nfeatures, iam::orb_pattern(), max_out, out_kp, out_des, out_n, &c->orb_scratch, c->stream, &err);
  if (rc < 0) return fail(rc == -5 ? IAM_E_UNSUPPORTED : rc == -1 ? IAM_E_ARG : IAM_E_CUDA, "%s", err.c_str());
  c->timing.total_launches += rc;
  return IAM_OK;
}

// ---- SIFT detect + describe (image.py:236-237, :324) ------------------------------------------------------------

int iam_sift_detect(iam_ctx* c, const uint8_t* gray, int width, int height, int max_out, float* out_kp, int32_t* out_octave,
                    uint8_t* out_des, int* out_n) {
  int rc = bind(c);
  if (rc) return rc;
  if (!gray || !out_kp || !out_octave || !out_des || !out_n || max_out < 0) return fail(IAM_E_ARG, "bad arguments");
  if (width < 2 || height < 2 || width > 16384 || height > 16384) return fail(IAM_E_ARG, "image size %d x %d out of range", width, height);
  std::string err;
  rc = iam::sift_detect(gray, width, height, max_out, out_kp, out_octave, out_des, out_n, &c->sift_scratch, c->stream, &err);
  if (rc < 0) return fail(rc == -5 ? IAM_E_UNSUPPORTED : rc == -1 ? IAM_E_ARG : rc == -3 ? IAM_E_NOMEM : IAM_E_CUDA, "%s", err.c_str());
  c->timing.total_launches += rc;
  return IAM_OK;
}

int iam_debug_orb_fast(iam_ctx* c, const uint8_t* gray, int width, int height, uint8_t* out_score) {
  int rc = bind(c);
  if (rc) return rc;
  if (!gray || !out_score || width <= 0 || height <= 0) return fail(IAM_E_ARG, "bad arguments");
  std::string err;
  rc = iam::orb_debug_fast_scores(gray, width, height, out_score, c->stream, &err);
  if (rc) return fail(IAM_E_CUDA, "%s", err.c_str());
  return IAM_OK;
}

// ---- bundle-adjustment residual / Jacobian (optimizer.py:174-279) ----------------------------------------------

int iam_ba_setup(iam_ctx* c, int n_cam, int n_pts, int n_obs, const int32_t* cam_idx, const int32_t* pt_idx,
                 const double* obs_uv) {
  int rc = bind(c);
  if (rc) return rc;
  if (n_cam < 0 || n_pts < 0 || n_obs < 0 || (n_obs > 0 && (!cam_idx || !pt_idx || !obs_uv))) return fail(IAM_E_ARG, "bad arguments");
  for (int i = 0; i < n_obs; ++i)
    if (cam_idx[i] < 0 || cam_idx[i] >= n_cam || pt_idx[i] < 0 || pt_idx[i] >= n_pts)
      return fail(IAM_E_ARG, "observation %d refers to camera %d / point %d outside the problem", i, cam_idx[i], pt_idx[i]);
  CU(cudaStreamSynchronize(c->stream));
  c->ba_n_obs = -1;
  const size_t no = std::max(1, n_obs);
  CU(c->ba_params.ensure((size_t(n_cam) * 7 + size_t(n_pts) * 3 + 1) * sizeof(double)));
  CU(c->ba_cam_idx.ensure(no * sizeof(int)));
  CU(c->ba_pt_idx.ensure(no * sizeof(int)));
  CU(c->ba_obs.ensure(no * 2 * sizeof(double)));
  CU(c->ba_res.ensure(no * 2 * sizeof(double)));
  if (n_obs > 0) {
    CU(cudaMemcpyAsync(c->ba_cam_idx.p, cam_idx, size_t(n_obs) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->ba_pt_idx.p, pt_idx, size_t(n_obs) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->ba_obs.p, obs_uv, size_t(n_obs) * 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  c->ba_n_cam = n_cam;
  c->ba_n_pts = n_pts;
  c->ba_n_obs = n_obs;
  return IAM_OK;
}

int iam_ba_upload_params(iam_ctx* c, const double* params) {
  int rc = bind(c);
  if (rc) return rc;
  if (c->ba_n_obs < 0) return fail(IAM_E_STATE, "iam_ba_setup has not been called");
  if (!params) return fail(IAM_E_ARG, "null parameter vector");
  const size_t n = size_t(c->ba_n_cam) * 7 + size_t(c->ba_n_pts) * 3;
  if (n) CU(cudaMemcpyAsync(c->ba_params.p, params, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return IAM_OK;
}

int iam_ba_eval_device(iam_ctx* c, const double* K4, const double* dist5, int want_jac, void** d_residual, void** d_jac) {
  int rc = bind(c);
  if (rc) return rc;
  if (c->ba_n_obs < 0) return fail(IAM_E_STATE, "iam_ba_setup has not been called");
  if (!K4 || !dist5) return fail(IAM_E_ARG, "K4 and dist5 are required");
  iam::BaCalib cal{K4[0], K4[1], K4[2], K4[3], dist5[0], dist5[1], dist5[2], dist5[3], dist5[4]};
  if (want_jac) CU(c->ba_jac.ensure(std::max<size_t>(1, c->ba_n_obs) * 2 * iam::kBaJacCols * sizeof(double)));
  const double* cams = c->ba_params.as<double>();
  cudaError_t e = iam::launch_ba(cams, cams + size_t(c->ba_n_cam) * 7, c->ba_cam_idx.as<int>(), c->ba_pt_idx.as<int>(),
                                 c->ba_obs.as<double>(), c->ba_n_obs, cal, c->ba_res.as<double>(),
                                 want_jac ? c->ba_jac.as<double>() : nullptr, c->stream);
  if (e != cudaSuccess) return fail(IAM_E_CUDA, "residual launch: %s", cudaGetErrorString(e));
  c->timing.total_launches += 1;
  if (d_residual) *d_residual = c->ba_res.p;
  if (d_jac) *d_jac = want_jac ? c->ba_jac.p : nullptr;
  return IAM_OK;
}

int iam_ba_eval(iam_ctx* c, const double* params, const double* K4, const double* dist5, double* out_residual,
                double* out_jac) {
  if (!out_residual) return fail(IAM_E_ARG, "null output");
  int rc = iam_ba_upload_params(c, params);
  if (rc) return rc;
  if ((rc = iam_ba_eval_device(c, K4, dist5, out_jac != nullptr, nullptr, nullptr)) != IAM_OK) return rc;
  if (c->ba_n_obs > 0) {
    CU(cudaMemcpyAsync(out_residual, c->ba_res.p, size_t(c->ba_n_obs) * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (out_jac)
      CU(cudaMemcpyAsync(out_jac, c->ba_jac.p, size_t(c->ba_n_obs) * 2 * iam::kBaJacCols * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  return IAM_OK;
}

int iam_ba_calib_jacobian(iam_ctx* c, const double* K4, const double* dist5, double* out_jac_calib) {
  int rc = bind(c);
  if (rc) return rc;
  if (c->ba_n_obs < 0) return fail(IAM_E_STATE, "iam_ba_setup has not been called");
  if (!K4 || !dist5 || !out_jac_calib) return fail(IAM_E_ARG, "K4, dist5 and the output are required");
  iam::BaCalib cal{K4[0], K4[1], K4[2], K4[3], dist5[0], dist5[1], dist5[2], dist5[3], dist5[4]};
  const size_t bytes = std::max<size_t>(1, c->ba_n_obs) * 16 * sizeof(double);
  CU(c->ba_jac_cal.ensure(bytes));
  const double* cams = c->ba_params.as<double>();
  cudaError_t e = iam::launch_ba_calib(cams, cams + size_t(c->ba_n_cam) * 7, c->ba_cam_idx.as<int>(), c->ba_pt_idx.as<int>(),
                                       c->ba_n_obs, cal, c->ba_jac_cal.as<double>(), c->stream);
  if (e != cudaSuccess) return fail(IAM_E_CUDA, "calibration Jacobian launch: %s", cudaGetErrorString(e));
  c->timing.total_launches += 1;
  if (c->ba_n_obs > 0)
    CU(cudaMemcpyAsync(out_jac_calib, c->ba_jac_cal.p, size_t(c->ba_n_obs) * 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return IAM_OK;
}

int iam_debug_ba_host(const double* cam7, const double* pt3, const double* uv, const double* K4, const double* dist5,
                      double* out_res2, double* out_jac20) {
  if (!cam7 || !pt3 || !uv || !K4 || !dist5 || !out_res2) return fail(IAM_E_ARG, "null argument");
  iam::BaCalib cal{K4[0], K4[1], K4[2], K4[3], dist5[0], dist5[1], dist5[2], dist5[3], dist5[4]};
  iam::ba_observation(cam7, pt3, uv[0], uv[1], cal, out_res2, out_jac20);
  return IAM_OK;
}

}  // extern "C"
