// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences), cluster helpers.
// Everything here is hand-written for Blackwell; there is no fallback path.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace jegal {
namespace ptx {

// ---------------------------------------------------------------- addresses
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// In a CTA pair the even CTA ("leader") owns the TMA-full / tmem-empty
// barriers.  Clearing bit 24 of a shared::cluster address maps a local smem
// offset onto the leader's copy of it.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Arrive on a barrier addressed in the shared::cluster window (own or peer CTA).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  // default semantics (.release at .cta scope): the only thing ordered before this arrive is the
  // tcgen05.ld of the accumulator (tcgen05.wait::ld + fence::before_thread_sync); a cluster-scope
  // release would drain every outstanding global store first (ERRBAR, 16 % of the epilogue)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

#ifndef JEGAL_WATCHDOG_SPINS
// A wait that does not complete within this many try_wait rounds traps instead
// of hanging the GPU (each try_wait sleeps in hardware for up to ~a few us).
#define JEGAL_WATCHDOG_SPINS (1u << 24)
#endif

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > JEGAL_WATCHDOG_SPINS) {
      printf("jegal: mbarrier watchdog: block %d thread %d bar 0x%x parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// Lean form for kernels whose hot code must stay inside the 32 KB L1.5 instruction cache: the poll loop
// and the watchdog's printf live out of line, a wait site is one try_wait and a predicated call.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > JEGAL_WATCHDOG_SPINS) {
      printf("jegal: mbarrier watchdog: block %d thread %d bar 0x%x parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

// 2D tiled load, completion bytes signalled on `bar` of this CTA.
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint32_t bar, uint32_t dst,
                                            int32_t c0, int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// Same, issued from either CTA of a pair; completion is signalled on the
// LEADER CTA's barrier (bar already masked with kPeerBitMask by the caller).
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint32_t bar_leader,
                                                 uint32_t dst, int32_t c0, int32_t c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      ".L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_leader), "r"(c0), "r"(c1),
        "l"(policy)
      : "memory");
}

__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ---------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

template <int kCtaGroup>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  if constexpr (kCtaGroup == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int kCtaGroup>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (kCtaGroup == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
  } else {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
  }
}

// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major, kind::f16 (bf16/fp16 in, fp32 out).
template <int kCtaGroup>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
  if constexpr (kCtaGroup == 1) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// Make `bar` (same smem offset in every CTA of `cta_mask`) receive one arrival
// when all tcgen05 ops issued so far by this thread have completed.
template <int kCtaGroup>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if constexpr (kCtaGroup == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     bar)
                 : "memory");
  } else {
    const uint16_t mask = 0x3;
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
        " [%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
  }
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane
// (lane_base + i), registers r[0..31] = columns col..col+31.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- descriptors
// K-major operand tile in smem written by TMA with CU_TENSOR_MAP_SWIZZLE_128B:
// rows of 128 B (64 x 16-bit), 8-row swizzle atoms 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // start address, 16 B units
  d |= static_cast<uint64_t>(1) << 16;                      // LBO (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // SBO: 8 rows x 128 B
  d |= static_cast<uint64_t>(1) << 46;                      // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor. ab_fmt: 0 = fp16, 1 = bf16. fp32 accumulate,
// both operands K-major, no negate / sparsity / saturate.
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t ab_fmt, uint32_t m, uint32_t n) {
  return (1u << 4) | (ab_fmt << 7) | (ab_fmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ---------------------------------------------------------------- math bits
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// atomic max on fp32 through integer atomics (dst initialised to -inf or any float).
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  if (v >= 0.0f) {
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  } else {
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
  }
}

// Explicit global-space forms (a generic-address atomicAdd inside a non-inlined function drags in
// the shared/local address-space dispatch).
__device__ __forceinline__ void st_global_cs_f32(float* addr, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
// 256-bit streaming store (sm_100: STG.E.EF.256): one full 32-byte sector per lane and instruction
__device__ __forceinline__ void st_global_cs_v8_f32(float* addr, const float (&v)[8]) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(addr), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void red_global_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_global_max_f32(float* addr, float v) {
  if (v >= 0.0f) {
    asm volatile("red.global.max.s32 [%0], %1;" ::"l"(addr), "r"(__float_as_int(v)) : "memory");
  } else {
    asm volatile("red.global.min.u32 [%0], %1;" ::"l"(addr), "r"(__float_as_uint(v)) : "memory");
  }
}

}  // namespace ptx
}  // namespace jegal
