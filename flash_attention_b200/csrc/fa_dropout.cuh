// fa_dropout.cuh — attention dropout: counter-based RNG shared by the forward, the backward and the S_dmask
// debug kernel (reference: csrc/flash_attn/src/dropout.h:15-93, philox.cuh:29-51 -- Philox4x32 with 7 rounds,
// one byte of randomness per attention element, keep iff byte <= floor(p_keep * 255)).
//
// Mapping = the REFERENCE's stream, bit for bit (dropout.h:20-47, philox.cuh:29-51), so that the same (seed, offset)
// drops the same elements as the reference's FA2 kernels (pinned on the GPU against the sign pattern of their S_dmask:
// tests/test_dropout_gpu.py).  FA2 draws one Philox4x32-7 value per thread and 16 x 32 (query x key) block of its
// mma.sync m16n8k16 accumulator; written out per ELEMENT (query R, key C) of head bh = batch * h + head:
//     counter = (lo, hi of [offset + bh * 32 + lane],  R / 16,  C / 32)        key = (seed_lo, seed_hi)
//     lane    = (R % 8) * 4 + (C % 8) / 2                                      (the accumulator's owner thread)
//     byte    = ((C % 32) / 16) * 8 + ((C % 16) / 8) * 4 + (C % 2) + 2 * ((R % 16) / 8)   of the 16 output bytes
// keep iff byte <= floor(p_keep * 255).  A forward thread here owns one query row: per 32 keys it makes 4 calls (the 4
// lanes of its row) and uses 8 bytes of each; a backward thread owns one key: per 16 queries it makes 8 calls and uses 2
// bytes of each.  The mask is a pure function of (seed, offset, batch, head, query, key): the backward regenerates it from
// rng_state, the CPU oracle restates it (oracle/attention_oracle.py:philox_keep_mask).
#pragma once
#include <cstdint>

namespace fa {

struct DropoutParams {
  uint32_t seed_lo, seed_hi, off_lo, off_hi;
  uint32_t thr;  // keep iff byte <= thr, thr = floor(p_keep * 255)   (dropout.h:17)
  float rp;      // 1 / p_keep
};

__device__ __forceinline__ uint4 philox4x32_7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// the reference's draw for accumulator-owner `lane` of the 16 x 32 block (row_blk, col_blk) of head `bh`
__device__ __forceinline__ uint4 dropout_draw(const DropoutParams& dp, uint32_t row_blk, uint32_t col_blk, uint32_t bh, uint32_t lane) {
  const uint64_t off = ((uint64_t(dp.off_hi) << 32) | dp.off_lo) + uint64_t(bh) * 32u + lane;
  return philox4x32_7(uint32_t(off), uint32_t(off >> 32), row_blk, col_blk, dp.seed_lo, dp.seed_hi);
}
__device__ __forceinline__ uint32_t draw_byte(const uint4& r, int idx) {   // byte `idx` (0..15) of the 16 output bytes
  const uint32_t w = (idx >> 2) == 0 ? r.x : ((idx >> 2) == 1 ? r.y : ((idx >> 2) == 2 ? r.z : r.w));
  return (w >> (8 * (idx & 3))) & 0xffu;
}
// one element (slow path: S_dmask checker kernel)
__device__ __forceinline__ bool dropout_keep(const DropoutParams& dp, uint32_t R, uint32_t C, uint32_t bh) {
  const uint4 r = dropout_draw(dp, R >> 4, C >> 5, bh, (R & 7u) * 4u + ((C & 7u) >> 1));
  const int idx = int(((C & 31u) >> 4) * 8u + ((C & 15u) >> 3) * 4u + (C & 1u) + 2u * ((R & 15u) >> 3));
  return draw_byte(r, idx) <= dp.thr;
}
// forward thread (one query row R): keep bits of the 32 keys [32 * col_blk, 32 * col_blk + 32), bit c = key offset c
__device__ __forceinline__ uint32_t dropout_keep_row32(const DropoutParams& dp, uint32_t R, uint32_t col_blk, uint32_t bh) {
  uint32_t keep = 0;
  const int vhi = int((R & 15u) >> 3) * 2;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const uint4 r = dropout_draw(dp, R >> 4, col_blk, bh, (R & 7u) * 4u + uint32_t(t));
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          keep |= (draw_byte(r, jj * 8 + a * 4 + e + vhi) <= dp.thr ? 1u : 0u) << (jj * 16 + a * 8 + 2 * t + e);
  }
  return keep;
}
// backward thread (one key C): keep bits of the 16 queries [16 * row_blk, 16 * row_blk + 16), bit r = query offset r
__device__ __forceinline__ uint32_t dropout_keep_col16(const DropoutParams& dp, uint32_t row_blk, uint32_t C, uint32_t bh) {
  uint32_t keep = 0;
  const int base = int(((C & 31u) >> 4) * 8u + ((C & 15u) >> 3) * 4u + (C & 1u));
  const uint32_t t = (C & 7u) >> 1;
#pragma unroll
  for (int gq = 0; gq < 8; ++gq) {
    const uint4 r = dropout_draw(dp, row_blk, C >> 5, bh, uint32_t(gq) * 4u + t);
    keep |= (draw_byte(r, base) <= dp.thr ? 1u : 0u) << gq;
    keep |= (draw_byte(r, base + 2) <= dp.thr ? 1u : 0u) << (gq + 8);
  }
  return keep;
}

// ---------------------------------------------------------------------------------------------------------
// S_dmask (return_softmax=True; "for testing only", flash_attn_interface.py:1196): the NORMALISED attention
// probabilities exp(s - lse), negated where the element is dropped (sign-bit encoding, dropout.h:34-36).
// Not fused and not fast: one warp per query row, recomputes q.k in fp32.
struct DmaskParams {
  const void* q; const void* k; const float* lse; void* s_out;
  int64_t q_batch_stride, q_row_stride, q_head_stride, k_batch_stride, k_row_stride, k_head_stride;
  const int* cu_seqlens_q; const int* cu_seqlens_k;
  const float* alibi_slopes; int64_t alibi_batch_stride;
  int b, h, h_k, sq, sk, d, sq_r, sk_r, total_q;
  int has_left, has_right, wl, wr;
  float scale;
  DropoutParams drop;
};

template <typename T>
__global__ void __launch_bounds__(256) fa_dmask_kernel(const DmaskParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp, head = blockIdx.y, bidb = blockIdx.z;
  if (i >= p.sq_r) return;
  int q_off = 0, sq_b = p.sq, k_off = 0, sk_b = p.sk;
  if (p.cu_seqlens_q) { q_off = p.cu_seqlens_q[bidb]; sq_b = p.cu_seqlens_q[bidb + 1] - q_off; }
  if (p.cu_seqlens_k) { k_off = p.cu_seqlens_k[bidb]; sk_b = p.cu_seqlens_k[bidb + 1] - k_off; }
  T* out = reinterpret_cast<T*>(p.s_out) + ((int64_t(bidb) * p.h + head) * p.sq_r + i) * p.sk_r;
  if (i >= sq_b) {
    for (int j = lane; j < p.sk_r; j += 32) out[j] = T(0.f);
    return;
  }
  const int shift = sk_b - sq_b;
  const T* qrow = reinterpret_cast<const T*>(p.q) + (p.cu_seqlens_q ? 0 : int64_t(bidb) * p.q_batch_stride) +
                  int64_t(q_off + i) * p.q_row_stride + int64_t(head) * p.q_head_stride;
  const T* kbase = reinterpret_cast<const T*>(p.k) + (p.cu_seqlens_k ? 0 : int64_t(bidb) * p.k_batch_stride) +
                   int64_t(head / (p.h / p.h_k)) * p.k_head_stride;
  const float lse = p.cu_seqlens_q ? p.lse[int64_t(head) * p.total_q + q_off + i]
                                   : p.lse[(int64_t(bidb) * p.h + head) * p.sq + i];
  const float slope = p.alibi_slopes ? p.alibi_slopes[bidb * p.alibi_batch_stride + head] : 0.f;
  for (int j = lane; j < p.sk_r; j += 32) {
    float val = 0.f;
    bool vis = j < sk_b;
    if (p.has_right) vis = vis && j <= i + shift + p.wr;
    if (p.has_left) vis = vis && j >= i + shift - p.wl;
    if (vis) {
      const T* krow = kbase + int64_t(k_off + j) * p.k_row_stride;
      float acc = 0.f;
      for (int c = 0; c < p.d; ++c) acc = fmaf(float(qrow[c]), float(krow[c]), acc);
      acc = acc * p.scale - slope * fabsf(float(i + shift - j));
      val = __expf(acc - lse);
      const bool keep = dropout_keep(p.drop, uint32_t(i), uint32_t(j), uint32_t(bidb * p.h + head));
      // dropped: negative sign.  A probability that underflows to 0 (in fp32, or when cast to fp16) would lose the sign
      // (-0.0 >= 0 is true), so a dropped entry is stored as at most -2^-24 (smallest fp16 subnormal, exact in bf16 too)
      if (!keep) val = -fmaxf(val, 5.9604645e-8f);
    }
    out[j] = T(val);
  }
}

}  // namespace fa
