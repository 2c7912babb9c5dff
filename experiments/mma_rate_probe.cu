// mma_rate_probe.cu — tcgen05.mma issue/execute rate on sm_100a: single-CTA vs CTA-pair, SS vs TS, M = 128 / 256.
// One thread issues `n` MMAs (N = 128, K = 16 each) back to back, commits, waits; prints clk per MMA.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

#include "../flash_attention_b200/csrc/fa_ptx.cuh"
#include "../flash_attention_b200/csrc/fa_tmem_ldst.cuh"

using namespace fa;

template <int kCtas>
__device__ __forceinline__ void body(long long* out, int mode, int n, int N) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bar = base + 163840, tptr = base + 163840 + 16;   // (N = 256 operands reach beyond the first 64 KB)
  const uint32_t rank = kCtas == 2 ? cluster_ctarank() : 0;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int idx = tid; idx < 65536 / 4; idx += blockDim.x) reinterpret_cast<uint32_t*>(smem)[idx] = 0;
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  __syncthreads();
  if (warp == 0) {
    if (kCtas == 2) { tmem_alloc_pair(tptr, 512); tmem_relinquish_pair(); }
    else { tmem_alloc(tptr, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 163840 + 16);
  // mode: 0 SS K-major, 1 TS (A from TMEM cols 256..), 2 SS MN-major A+B, 3 (pair only) M=128 SS MN-major
  if (rank == 0 && warp == 1 && elect_one()) {
    const uint32_t M = kCtas == 2 ? (mode == 3 ? 128 : 256) : 128;
    const uint32_t a_mn = (mode >= 2) ? 1 : 0, b_mn = (mode >= 1) ? 1 : 0;
    const uint32_t idesc = make_idesc(1, 1, mode == 1 ? 0 : a_mn, b_mn, M, uint32_t(N));
    const uint64_t dk = smem_desc_sw128(16, 1024), dmn = smem_desc_sw128(16384, 1024);
    uint32_t alo[8], blo[8];
    const uint32_t ahi = uint32_t((a_mn ? dmn : dk) >> 32), bhi = uint32_t((b_mn ? dmn : dk) >> 32);
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      alo[kk] = uint32_t(a_mn ? smem_desc_at(dmn, base + kk * 2048) : smem_desc_at(dk, base + (kk / 4) * 16384 + (kk % 4) * 32));
      blo[kk] = uint32_t(b_mn ? smem_desc_at(dmn, base + 32768 + kk * 2048)
                              : smem_desc_at(dk, base + 32768 + (kk / 4) * 16384 + (kk % 4) * 32));
    }
    const long long t0 = clock64();
    for (int i = 0; i < n; i += 8) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t ad = (uint64_t(ahi) << 32) | alo[kk], bd = (uint64_t(bhi) << 32) | blo[kk];
        if (kCtas == 2) {
          if (mode == 1) mma_ts_pair(tmem, tmem + 256 + kk * 8, bd, idesc, 1);
          else mma_ss_pair(tmem, ad, bd, idesc, 1);
        } else {
          if (mode == 1) mma_ts<false>(tmem, tmem + 256 + kk * 8, bd, idesc, 1);
          else mma_ss<false>(tmem, ad, bd, idesc, 1);
        }
      }
    }
    const long long t1 = clock64();
    if (kCtas == 2) mma_commit_pair(bar); else mma_commit(bar);
    mbar_wait(bar, 0, 1);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) { if (kCtas == 2) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512); }
}

__global__ void __launch_bounds__(128) rate1(long long* out, int mode, int n, int N) { body<1>(out, mode, n, N); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) rate2(long long* out, int mode, int n, int N) { body<2>(out, mode, n, N); }

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate1, cudaFuncAttributeMaxDynamicSharedMemorySize, 161 * 1024);
  cudaFuncSetAttribute(rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, 161 * 1024);
  const char* names[4] = {"SS K-major", "TS (A in TMEM), B MN-major", "SS MN-major A and B", "SS MN-major, M=128 pair (64 rows/CTA)"};
  for (int N : {64, 128, 256}) {
    const int n = 512;
    for (int ctas = 1; ctas <= 2; ++ctas) {
      for (int mode = 0; mode < (ctas == 2 ? 4 : 3); ++mode) {
        if (mode == 2) continue;
        long long h[2] = {0, 0};
        for (int rep = 0; rep < 2; ++rep) {
          if (ctas == 1) rate1<<<1, 128, 161 * 1024>>>(d, mode, n, N);
          else rate2<<<2, 128, 161 * 1024>>>(d, mode, n, N);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        const int M = ctas == 2 ? (mode == 3 ? 128 : 256) : 128;
        const double clk = double(h[1]) / n;
        printf("cta_group::%d %-42s M=%3d N=%3d K=16: %6.1f clk/MMA  -> %5.0f FLOP/clk/SM (peak 8192)\n", ctas, names[mode], M, N, clk,
               2.0 * (M / ctas) * N * 16 / clk);
      }
    }
  }
  return 0;
}
