// fa_kvcache.cuh — decode prologue: append new K/V rows into the (paged) cache at cache_seqlens[b] + t and
// apply rotary embedding to the new keys and to the queries (plain CUDA, HBM-bound, tiny).
// Semantics: csrc/flash_attn/src/flash_fwd_kernel.h:691-828 + rotary.h; python contract
// flash_attn/flash_attn_interface.py:1516-1524: keys are rotated at positions cache_seqlens + t; queries at
// cache_seqlens + i when causal/local, else all at cache_seqlens.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdint>

namespace fa {

struct AppendParams {
  void* kcache; void* vcache;          // (b_c, sk, h_k, d) or paged (pages, page, h_k, d)
  const void* knew; const void* vnew;  // (b, s_new, h_k, d)
  const void* q; void* q_out;          // (b, sq, h, d): rotated copy of q (only with rotary)
  int64_t kc_bs, kc_rs, kc_hs, vc_bs, vc_rs, vc_hs;
  int64_t kn_bs, kn_rs, kn_hs, vn_bs, vn_rs, vn_hs;
  int64_t q_bs, q_rs, q_hs, qo_bs, qo_rs, qo_hs;
  const int32_t* cache_seqlens;        // (b)
  const int32_t* cache_batch_idx;      // (b) or null
  const int32_t* leftpad_k;            // (b) or null
  const int32_t* block_table; int64_t bt_bs; int page_size;
  const void* rot_cos; const void* rot_sin;  // (seqlen_ro, rotary_dim/2) same dtype as q, or null
  int rotary_dim; int interleaved; int q_pos_per_row;  // q_pos_per_row: 1 = causal/local (pos = len + i), 0 = all at len
  int b, s_new, sq, h, h_k, d;
};

template <typename T>
__device__ __forceinline__ void rot_pair(const T* cs, const T* sn, int pos, int half, int i, float x1, float x2,
                                         float& o1, float& o2) {
  const float c = float(cs[int64_t(pos) * half + i]), s = float(sn[int64_t(pos) * half + i]);
  o1 = x1 * c - x2 * s;
  o2 = x1 * s + x2 * c;
}

// grid: (max(s_new, sq), max(h, h_k), b), block: 128 threads; thread i handles element pairs
template <typename T>
__global__ void __launch_bounds__(128) fa_kv_append_kernel(const AppendParams p) {
  const int t = blockIdx.x, head = blockIdx.y, bidb = blockIdx.z;
  const int len = p.cache_seqlens[bidb];              // absolute rows already in the cache
  const int lp = p.leftpad_k ? p.leftpad_k[bidb] : 0;  // keys [lp, len) are valid: logical length len - lp
  const int half = p.rotary_dim / 2;
  const T* cs = reinterpret_cast<const T*>(p.rot_cos);
  const T* sn = reinterpret_cast<const T*>(p.rot_sin);
  // ---- new key / value row t of kv-head `head`
  if (p.knew && t < p.s_new && head < p.h_k) {
    const int pos = len - lp + t;  // logical position in the sequence (rotary index)
    int64_t kdst, vdst;
    if (p.block_table) {
      const int row = len + t;
      const int page = p.block_table[bidb * p.bt_bs + row / p.page_size];
      kdst = int64_t(page) * p.kc_bs + int64_t(row % p.page_size) * p.kc_rs + int64_t(head) * p.kc_hs;
      vdst = int64_t(page) * p.vc_bs + int64_t(row % p.page_size) * p.vc_rs + int64_t(head) * p.vc_hs;
    } else {
      const int cb = p.cache_batch_idx ? p.cache_batch_idx[bidb] : bidb;
      const int row = len + t;
      kdst = int64_t(cb) * p.kc_bs + int64_t(row) * p.kc_rs + int64_t(head) * p.kc_hs;
      vdst = int64_t(cb) * p.vc_bs + int64_t(row) * p.vc_rs + int64_t(head) * p.vc_hs;
    }
    const T* ks = reinterpret_cast<const T*>(p.knew) + int64_t(bidb) * p.kn_bs + int64_t(t) * p.kn_rs + int64_t(head) * p.kn_hs;
    const T* vs = reinterpret_cast<const T*>(p.vnew) + int64_t(bidb) * p.vn_bs + int64_t(t) * p.vn_rs + int64_t(head) * p.vn_hs;
    T* kd = reinterpret_cast<T*>(p.kcache) + kdst;
    T* vd = reinterpret_cast<T*>(p.vcache) + vdst;
    for (int i = threadIdx.x; i < p.d; i += 128) vd[i] = vs[i];
    if (cs) {
      for (int i = threadIdx.x; i < half; i += 128) {
        const int a = p.interleaved ? 2 * i : i, c = p.interleaved ? 2 * i + 1 : i + half;
        float o1, o2;
        rot_pair(cs, sn, pos, half, i, float(ks[a]), float(ks[c]), o1, o2);
        kd[a] = T(o1);
        kd[c] = T(o2);
      }
      for (int i = p.rotary_dim + threadIdx.x; i < p.d; i += 128) kd[i] = ks[i];
    } else {
      for (int i = threadIdx.x; i < p.d; i += 128) kd[i] = ks[i];
    }
  }
  // ---- query row t of head `head` (rotary only)
  if (cs && p.q_out && t < p.sq && head < p.h) {
    const int pos = p.q_pos_per_row ? len - lp + t : len - lp;
    const T* qs = reinterpret_cast<const T*>(p.q) + int64_t(bidb) * p.q_bs + int64_t(t) * p.q_rs + int64_t(head) * p.q_hs;
    T* qd = reinterpret_cast<T*>(p.q_out) + int64_t(bidb) * p.qo_bs + int64_t(t) * p.qo_rs + int64_t(head) * p.qo_hs;
    for (int i = threadIdx.x; i < half; i += 128) {
      const int a = p.interleaved ? 2 * i : i, c = p.interleaved ? 2 * i + 1 : i + half;
      float o1, o2;
      rot_pair(cs, sn, pos, half, i, float(qs[a]), float(qs[c]), o1, o2);
      qd[a] = T(o1);
      qd[c] = T(o2);
    }
    for (int i = p.rotary_dim + threadIdx.x; i < p.d; i += 128) qd[i] = qs[i];
  }
}

}  // namespace fa
