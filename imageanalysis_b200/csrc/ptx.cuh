// ptx.cuh — thin inline-PTX wrappers for the sm_100a features this library
// uses: mbarrier, bulk async copy (TMA 1-D, SASS UBLKCP), tcgen05 MMA /
// TMEM alloc / TMEM load (SASS UTC*MMA / LDTM), and the UMMA shared-memory
// matrix descriptor.  Bit layouts follow the PTX ISA "tcgen05 matrix
// descriptor" / "instruction descriptor" tables.
#pragma once
#include <cstdint>
#include <cstdio>

#include "layout.h"

namespace iam {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)  // suspend-time hint: sleep in hardware, do not spin on issue slots
      : "memory");
  return ok != 0;
}

#ifndef IAM_SPIN_LIMIT
#define IAM_SPIN_LIMIT (1u << 26)   // failed try_waits (each a hardware sleep) before a lost arrival traps instead of hanging the GPU
#endif

// The spin body is deliberately tiny (TRYWAIT, BRA, IADD, ISETP): waiting warps share the ALU pipe
// with the working epilogue warps, so every instruction spent polling is stolen from them.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins == IAM_SPIN_LIMIT) {
      printf("iamatch: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, (int)blockIdx.x,
             (int)threadIdx.x, parity);
      __trap();
    }
  }
}

// Hot-loop wait without the watchdog counter (two ALU-pipe instructions less per poll).  Only for waits
// whose producer side is itself guarded by mbar_wait: a lost arrival still traps there.
__device__ __forceinline__ void mbar_wait_bare(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// Address-based variants for hot loops that keep shared-memory addresses in registers (no per-use
// generic -> shared conversion, offsets fold into the instruction's immediate).
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// try_wait with the implementation-defined (short) time limit: the warp is parked and woken by the barrier
__device__ __forceinline__ bool mbar_try_wait_nohint_a(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking poll
__device__ __forceinline__ bool mbar_test_wait_a(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity)
      : "memory");
  return ok != 0;
}
#ifndef IAM_EPI_WAIT
#define IAM_EPI_WAIT 0
#endif
#ifndef IAM_MMA_WAIT
#define IAM_MMA_WAIT 0
#endif
__device__ __forceinline__ void mbar_wait_bare_a(uint32_t bar_addr, uint32_t parity) {
#if IAM_EPI_WAIT == 0
  while (!mbar_try_wait_a(bar_addr, parity)) {
  }
#elif IAM_EPI_WAIT == 1
  while (!mbar_try_wait_nohint_a(bar_addr, parity)) {
  }
#else
  while (!mbar_test_wait_a(bar_addr, parity)) {
  }
#endif
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity, int tag = 0) {
  uint32_t spins = 0;
#if IAM_MMA_WAIT == 0
  while (!mbar_try_wait_a(bar_addr, parity)) {
#else
  while (!mbar_try_wait_nohint_a(bar_addr, parity)) {
#endif
    if (++spins == IAM_SPIN_LIMIT) {
      printf("iamatch: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, (int)blockIdx.x,
             (int)threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ float lds_volatile_f32_a(uint32_t addr) {
  float v;
  asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_volatile_f32_a(uint32_t addr, float v) {
  asm volatile("st.volatile.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// Opaque to the optimiser: the value is kept in a register instead of being recomputed at every use.
__device__ __forceinline__ uint32_t pin_reg(uint32_t x) {
  asm volatile("" : "+r"(x));
  return x;
}

// ------------------------------------------------------- bulk copy (TMA 1-D)
// global -> shared, completion reported on an mbarrier in bytes.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// same, but each CTA of the cluster named in `cta_mask` receives the bytes at the same
// CTA-relative offset and its own mbarrier (same offset) gets the complete_tx
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                                   uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}

// ------------------------------------------------------------------ clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ tcgen05
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// arrive on an mbarrier once all MMAs previously issued by this thread retire
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void umma_commit_a(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr)
               : "memory");
}

// same, arriving on the barrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// K-major, no-swizzle ("interleaved") operand descriptor.
//   element (row r, 16-byte K-chunk c) lives at  start + (r/8)*SBO + c*LBO + (r%8)*16
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (sm_100)
  // base_offset = 0, lbo_mode = 0, layout_type (bits 61..63) = 0: SWIZZLE_NONE
  return d;
}

// The same descriptor in two 32-bit halves, for issue loops that step the start address: the low word of
// a tile at byte offset `o` from a base is  desc_lo(base) + (o >> 4)  as long as the sum stays inside the
// 14-bit address field (always true for shared-memory addresses), the high word never changes.
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
__host__ __device__ constexpr uint32_t smem_desc_hi(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14);
}
__device__ __forceinline__ uint64_t pack_desc(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

// Instruction descriptor, dense, fp32 accumulate, both operands K-major.
//   fmt: kind::f16 -> 0 = F16, 1 = BF16 ; kind::f8f6f4 -> 0 = E4M3, 1 = E5M2
__host__ __device__ constexpr uint32_t make_idesc(uint32_t m, uint32_t n, uint32_t a_fmt, uint32_t b_fmt) {
  return (1u << 4) /* D = F32 */ | (a_fmt << 7) | (b_fmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// kind::i8: unsigned 8-bit operands (format 0), s32 accumulate (D format 2)
__host__ __device__ constexpr uint32_t make_idesc_i8(uint32_t m, uint32_t n) {
  return (2u << 4) /* D = S32 */ | ((n >> 3) << 17) | ((m >> 4) << 24);
}
template <Kind kKind>
__host__ __device__ constexpr uint32_t make_idesc_kind(uint32_t m, uint32_t n) {
  return kKind == Kind::I8 ? make_idesc_i8(m, n) : make_idesc(m, n, 0, 0);
}

template <Kind kKind>
__device__ __forceinline__ void umma(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kKind == Kind::F16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(taddr),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else if constexpr (kKind == Kind::F8) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(taddr),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(taddr),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// A operand from tensor memory ("TS" form): D[tmem] (+)= A[tmem] . B[smem]
template <Kind kKind>
__device__ __forceinline__ void umma_ts(uint32_t taddr, uint32_t a_taddr, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  if constexpr (kKind == Kind::F16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(taddr),
        "r"(a_taddr), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else if constexpr (kKind == Kind::F8) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}" ::"r"(taddr),
        "r"(a_taddr), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(taddr),
        "r"(a_taddr), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// shared memory -> tensor memory: 128 rows x 256 bits (one K-step slab of an A tile,
// addressed by the same matrix descriptor the SS-form MMA would use) -> 128 lanes x 8 columns
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns.
template <typename T>
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, T (&v)[32]) {
  static_assert(sizeof(T) == 4, "32-bit accumulator cells");
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// Wait for outstanding tcgen05.ld; the loaded registers are threaded through
// as in/out operands so the compiler cannot hoist their uses above the wait.
template <typename T>
__device__ __forceinline__ void tmem_ld_wait(T (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                 "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]),
                 "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]),
                 "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive 32-bit columns.
template <typename T>
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, T* v) {
  static_assert(sizeof(T) == 4, "32-bit accumulator cells");
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// NC (a multiple of 16) consecutive columns as x32 loads plus at most one x16 load, all in flight together.
template <int NC, typename T>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, T (&v)[NC]) {
  static_assert(NC % 16 == 0, "whole x16 loads");
#pragma unroll
  for (int c = 0; c + 32 <= NC; c += 32) tmem_ld32(taddr + c, *reinterpret_cast<T(*)[32]>(&v[c]));
  if constexpr (NC % 32 == 16) tmem_ld16(taddr + NC - 16, &v[NC - 16]);
}

// Wait for outstanding tcgen05.ld of NC registers: the wait itself, then every loaded register passes through an
// empty volatile asm as an in/out operand -- volatile asms keep their order, so no use of v[] can be hoisted above
// the wait (the same guarantee tmem_ld_wait gives with one 32-operand statement).
template <int NC, typename T>
__device__ __forceinline__ void tmem_ld_wait_n(T (&v)[NC]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
#pragma unroll
  for (int j = 0; j < NC; ++j) asm volatile("" : "+r"(r[j]));
}

// 16-byte shared-memory accesses that the compiler may neither cache nor split
__device__ __forceinline__ float4 lds_volatile_v4(const void* p) {
  float4 v;
  asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(smem_u32(p))
               : "memory");
  return v;
}
__device__ __forceinline__ void sts_volatile_v4(void* p, float4 v) {
  asm volatile("st.volatile.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__device__ __forceinline__ float lds_volatile_f32(const void* p) {
  float v;
  asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void sts_volatile_f32(void* p, float v) {
  asm volatile("st.volatile.shared.f32 [%0], %1;" ::"r"(smem_u32(p)), "f"(v) : "memory");
}

__device__ __forceinline__ float fmin3(float a, float b, float c) { return fminf(fminf(a, b), c); }  // -> FMNMX3
__device__ __forceinline__ int imax3(int a, int b, int c) { return max(max(a, b), c); }                // -> VIMNMX3
__device__ __forceinline__ uint2 lds_volatile_v2b32_a(uint32_t addr) {
  uint2 v;
  asm volatile("ld.volatile.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_volatile_v2b32_a(uint32_t addr, uint32_t x, uint32_t y) {
  asm volatile("st.volatile.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint32_t lds_volatile_b32_a(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_volatile_b32_a(uint32_t addr, uint32_t v) {
  asm volatile("st.volatile.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

}  // namespace iam
