// fa_dropout.cuh — attention dropout: counter-based RNG shared by the forward, the backward and the S_dmask
// debug kernel (reference: csrc/flash_attn/src/dropout.h:15-93, philox.cuh:29-51 -- Philox4x32 with 7 rounds,
// one byte of randomness per attention element, keep iff byte <= floor(p_keep * 255)).
//
// Mapping (ours, not FA2's thread-layout-dependent one): one Philox call covers a 4 x 4 patch of the
// (query, key) matrix of one (batch, head):
//     counter = (query / 4, key / 4, batch * h + head, offset_lo)     key = (seed_lo, seed_hi ^ offset_hi)
//     output word (query % 4), byte (key % 4)  ->  the element's random byte
// so a forward thread (one query, 128 keys) and a backward thread (one key, 64 queries) both use a quarter
// of every call they make, and the mask is a pure function of (seed, offset, batch, head, query, key):
// the backward regenerates it from rng_state, the CPU oracle restates it (oracle/attention_oracle.py:philox_keep_mask).
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

// the 4 x 4 patch (q4 = query / 4, k4 = key / 4) of head `bh`: word i = query 4*q4 + i, byte j = key 4*k4 + j
__device__ __forceinline__ uint4 dropout_patch(const DropoutParams& dp, uint32_t q4, uint32_t k4, uint32_t bh) {
  return philox4x32_7(q4, k4, bh, dp.off_lo, dp.seed_lo, dp.seed_hi ^ dp.off_hi);
}

__device__ __forceinline__ uint32_t select_word(const uint4& r, int i) {
  return i == 0 ? r.x : (i == 1 ? r.y : (i == 2 ? r.z : r.w));
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
      const uint4 r = dropout_patch(p.drop, uint32_t(i) >> 2, uint32_t(j) >> 2, uint32_t(bidb * p.h + head));
      const uint32_t byte = (select_word(r, i & 3) >> (8 * (j & 3))) & 0xffu;
      // dropped: negative sign.  A probability that underflows to 0 (in fp32, or when cast to fp16) would lose the sign
      // (-0.0 >= 0 is true), so a dropped entry is stored as at most -2^-24 (smallest fp16 subnormal, exact in bf16 too)
      if (byte > p.drop.thr) val = -fmaxf(val, 5.9604645e-8f);
    }
    out[j] = T(val);
  }
}

}  // namespace fa
