// fa_combine.cuh — split-KV combine (plain CUDA, HBM-bound).
// lse = logsumexp_s(lse_s); out = sum_s exp(lse_s - lse) * o_s; all partials -inf -> lse = +inf, out = 0
// (reference: csrc/flash_attn/src/flash_fwd_kernel.h:1117-1299, esp. :1196,1204).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdint>

namespace fa {

struct CombineParams {
  const float* o_partial;   // (splits, rows, d_part) fp32
  const float* lse_partial; // (splits, rows)
  void* out;
  float* lse;               // (rows) or nullptr
  int64_t rows;             // b * h * sq
  int splits;
  int d;                    // columns to write
  int d_part;               // row pitch of o_partial (kD of the forward kernel)
  // out addressing: row = (b*h + head)*sq + qpos  ->  out + b*obs + qpos*ors + head*ohs
  int h, sq;
  int64_t obs, ors, ohs;
};

template <typename T>
__device__ __forceinline__ T from_float(float x);
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ __half from_float<__half>(float x) { return __float2half_rn(x); }
template <> __device__ __forceinline__ float from_float<float>(float x) { return x; }

// one warp per output row; lane l owns columns [4l, 4l+4) (+128 per extra pass)
template <typename T>
__global__ void __launch_bounds__(256) fa_combine_kernel(const CombineParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * 8 + warp;
  if (row >= p.rows) return;
  float m = -INFINITY;
  for (int s = lane; s < p.splits; s += 32) m = fmaxf(m, p.lse_partial[int64_t(s) * p.rows + row]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float m_safe = (m == -INFINITY) ? 0.f : m;
  float tot = 0.f;
  for (int s = lane; s < p.splits; s += 32) tot += __expf(p.lse_partial[int64_t(s) * p.rows + row] - m_safe);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  const bool dead = !(tot > 0.f);
  const float lse = dead ? INFINITY : (__logf(tot) + m_safe);
  if (lane == 0 && p.lse) p.lse[row] = lse;

  const int64_t bh = row / p.sq;
  const int qpos = int(row - bh * p.sq);
  const int64_t bb = bh / p.h;
  const int head = int(bh - bb * p.h);
  T* out = reinterpret_cast<T*>(p.out) + bb * p.obs + int64_t(qpos) * p.ors + int64_t(head) * p.ohs;
  for (int c = lane * 4; c < p.d; c += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < p.splits; ++s) {
      const float w = dead ? 0.f : __expf(p.lse_partial[int64_t(s) * p.rows + row] - lse);
      const float4 v = *reinterpret_cast<const float4*>(p.o_partial + (int64_t(s) * p.rows + row) * p.d_part + c);
      acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
    }
    out[c] = from_float<T>(acc.x);
    out[c + 1] = from_float<T>(acc.y);
    out[c + 2] = from_float<T>(acc.z);
    out[c + 3] = from_float<T>(acc.w);
  }
}

}  // namespace fa
