// fa_ptx.cuh — thin inline-PTX layer for sm_100a: mbarrier, TMA, tcgen05 (MMA / TMEM), descriptors.
// Hand-written for this library; no CUTLASS/CuTe.  Field layouts of the descriptors follow the
// hardware ABI (PTX ISA "tcgen05 matrix descriptors"; the reference documents the same bit layout
// in flash_attn/cute/mma_sm100_desc.py:111-163,212-285).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "fa_tmem_ldst.cuh"

// FA_SPIN_GUARD=1 (build with FA_B200_SPIN_GUARD=1): every mbarrier wait gives up after ~2^24 polls, prints
// which barrier starved and traps -- bring-up aid; off in the shipped build.
#ifndef FA_SPIN_GUARD
#define FA_SPIN_GUARD 0
#endif

namespace fa {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t.reg .b32 R;\n\t"
      "elect.sync R|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
#if FA_SPIN_GUARD
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0) {  // wall-clock bound (2 s): a starved barrier is a protocol bug, say which one
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > 2000000000ull) {
        printf("[fa_b200] mbarrier timeout tag=%d block=(%d,%d,%d) thread=%d parity=%u\n", tag, blockIdx.x, blockIdx.y,
               blockIdx.z, threadIdx.x, parity);
        __trap();
      }
    }
  }
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}

// ---------------------------------------------------------------- TMA (cp.async.bulk.tensor)
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2,
                                                  int c3) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]
template <bool kFp8>
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  if constexpr (kFp8) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// D[tmem] (+)= A[tmem] * B[smem desc]
template <bool kFp8>
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  if constexpr (kFp8) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---------------------------------------------------------------- CTA pair (cluster of 2, tcgen05 cta_group::2)
// Measured with experiments/cta_pair_probe.cu on B200: a pair MMA with M = 256 gives each CTA 128 accumulator
// rows (lane = row, N columns); with M = 128 each CTA gets 64 rows laid out as lane = row + 64 * (n / (N/2)),
// column = n % (N/2).  B's N rows [0, N/2) come from the rank-0 CTA's shared memory, [N/2, N) from rank 1.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// Remote (or local) arrive through a shared::cluster address.  Default semantics (.release at CTA scope) on purpose:
// `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR and `.acquire.cluster` waits to CCTL.IVALL (L1 invalidate),
// which cost ~40 % of the pair kernel's issue slots when first tried (profiles/r01_pair_v0_ncu_stalls.txt).  What is
// ordered through these barriers is TMEM state (tcgen05 fences) and st.async payloads (tracked by tx-count), neither of
// which needs a generic-proxy cluster fence.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// 16-byte store into the shared memory of a CTA of the cluster that counts 16 tx-bytes on an mbarrier of THAT CTA
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, uint4 v, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cluster_bar)
               : "memory");
}
// 4-byte variant (work-item hand-over of the persistent pair kernel's scheduler)
__device__ __forceinline__ void st_async_b32(uint32_t cluster_addr, uint32_t v, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(cluster_addr), "r"(v), "r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t cluster_addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
// bulk copy from this CTA's shared memory into the shared memory of a CTA of the cluster; `bytes` (multiple of 16) are
// counted on an mbarrier of the DESTINATION CTA.  (Scattered 16-byte st.async moved the same data ~10x slower.)
__device__ __forceinline__ void bulk_copy_to_cluster(uint32_t cluster_dst, uint32_t src, uint32_t bytes, uint32_t cluster_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(cluster_dst), "r"(src), "r"(bytes), "r"(cluster_bar)
               : "memory");
}
// TMA load whose completion is signalled on an mbarrier of EITHER CTA of the pair (cluster address `cluster_bar`)
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* m, uint32_t cluster_bar, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier at this shared-memory offset in BOTH CTAs once all prior pair MMAs have completed
__device__ __forceinline__ void mma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(uint16_t(3))
               : "memory");
}

// ---------------------------------------------------------------- descriptors
// 64-bit shared-memory matrix descriptor, 128-byte swizzle, version 1 (Blackwell).
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1
//   bits [61,64) layout type (2 = SWIZZLE_128B)
__host__ __device__ constexpr uint64_t smem_desc_sw128(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16) | (uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__device__ __forceinline__ uint64_t smem_desc_at(uint64_t base_desc, uint32_t smem_addr) {
  return base_desc | uint64_t((smem_addr & 0x3FFFF) >> 4);
}
// 32-bit instruction descriptor (kind::f16 / kind::f8f6f4), fp32 accumulate.
//   a_fmt/b_fmt: kind::f16 -> 0 = fp16, 1 = bf16;  kind::f8f6f4 -> 0 = e4m3, 1 = e5m2
//   a_mn / b_mn: 0 = K-major operand, 1 = MN-major operand
__host__ __device__ constexpr uint32_t make_idesc(uint32_t a_fmt, uint32_t b_fmt, uint32_t a_mn, uint32_t b_mn,
                                                  uint32_t M, uint32_t N) {
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

// ---------------------------------------------------------------- math
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}
// ---- packed fp32x2 arithmetic (FFMA2 / FADD2 on sm_100: two fp32 lanes per instruction)
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// 2^x for a pair on the FMA pipe (no MUFU): round-to-nearest range reduction x = n + f, f in [-0.5, 0.5],
// 2^f by a degree-3 minimax polynomial (max rel. error 7.5e-5, far below the 16-bit rounding of P), 2^n by
// adding n to the exponent field.  Inputs must be <= ~120; inputs below -125 (incl. -inf of masked scores) give
// exactly 0 -- a fully masked row must keep a row sum of exactly 0.
__device__ __forceinline__ void ex2_poly2(float x0, float x1, float& r0, float& r1) {
  const float kMagic = 12582912.0f;  // 1.5 * 2^23: adding it rounds to an integer kept in the low mantissa bits
  const bool z0 = x0 < -125.0f, z1 = x1 < -125.0f;
  x0 = fmaxf(x0, -125.0f);
  x1 = fmaxf(x1, -125.0f);
  const uint64_t x = pack2(x0, x1);
  const uint64_t t = add2(x, pack2(kMagic, kMagic));
  const uint64_t n = add2(t, pack2(-kMagic, -kMagic));
  const uint64_t f = fma2(n, pack2(-1.0f, -1.0f), x);
  uint64_t pl = fma2(f, pack2(0.0551716648f, 0.0551716648f), pack2(0.2426111251f, 0.2426111251f));
  pl = fma2(pl, f, pack2(0.6932609677f, 0.6932609677f));
  pl = fma2(pl, f, pack2(0.9999280572f, 0.9999280572f));
  float t0, t1, p0, p1;
  unpack2(t, t0, t1);
  unpack2(pl, p0, p1);
  r0 = z0 ? 0.f : __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  r1 = z1 ? 0.f : __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// Same polynomial without the exact-zero select: inputs are clamped at -125, so a masked (-inf) score yields 2^-125 * p
// (~2e-38) instead of 0.  Harmless wherever the row is normalised by a finite row sum; callers that need "fully masked
// row" detection test the row max instead of `row_sum == 0` (saves one FSETP + select per element).
__device__ __forceinline__ void ex2_poly2_nz(uint64_t x, float& r0, float& r1) {
  const float kMagic = 12582912.0f;
  float x0, x1;
  unpack2(x, x0, x1);
  x = pack2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
  const uint64_t t = add2(x, pack2(kMagic, kMagic));
  const uint64_t n = add2(t, pack2(-kMagic, -kMagic));
  const uint64_t f = fma2(n, pack2(-1.0f, -1.0f), x);
  uint64_t pl = fma2(f, pack2(0.0551716648f, 0.0551716648f), pack2(0.2426111251f, 0.2426111251f));
  pl = fma2(pl, f, pack2(0.6932609677f, 0.6932609677f));
  pl = fma2(pl, f, pack2(0.9999280572f, 0.9999280572f));
  float t0, t1, p0, p1;
  unpack2(t, t0, t1);
  unpack2(pl, p0, p1);
  r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// degree-2 variant (rel. err 1.7e-3): enough when the result is quantised to e4m3 (3 mantissa bits); one FFMA2 less
__device__ __forceinline__ void ex2_poly2_deg2(float x0, float x1, float& r0, float& r1) {
  const float kMagic = 12582912.0f;
  const bool z0 = x0 < -125.0f, z1 = x1 < -125.0f;
  x0 = fmaxf(x0, -125.0f);
  x1 = fmaxf(x1, -125.0f);
  const uint64_t x = pack2(x0, x1);
  const uint64_t t = add2(x, pack2(kMagic, kMagic));
  const uint64_t n = add2(t, pack2(-kMagic, -kMagic));
  const uint64_t f = fma2(n, pack2(-1.0f, -1.0f), x);
  uint64_t pl = fma2(f, pack2(0.23842086f, 0.23842086f), pack2(0.70343314f, 0.70343314f));
  pl = fma2(pl, f, pack2(1.00044256f, 1.00044256f));
  float t0, t1, p0, p1;
  unpack2(t, t0, t1);
  unpack2(pl, p0, p1);
  r0 = z0 ? 0.f : __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  r1 = z1 ? 0.f : __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// {lo, hi} -> packed 16-bit pair (lo in bits [0,16))
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// four floats -> four e4m3 bytes (a in the lowest byte)
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
  uint16_t lo, hi;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(b), "f"(a));
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(d), "f"(c));
  return uint32_t(lo) | (uint32_t(hi) << 16);
}

}  // namespace fa
