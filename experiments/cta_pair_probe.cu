// cta_pair_probe.cu — bring-up probe for tcgen05 cta_group::2 (CTA-pair MMA) on sm_100a.
// Determines empirically (a) the TMEM accumulator layout of a pair MMA with M = 128 (64 rows per CTA) and
// M = 256, (b) which CTA supplies which half of the B operand (K-major and MN-major), (c) that the
// alloc / commit-multicast / cluster-barrier sequence used below is sufficient.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o experiments/cta_pair_probe.bin experiments/cta_pair_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../flash_attention_b200/csrc/fa_ptx.cuh"
#include "../flash_attention_b200/csrc/fa_tmem_ldst.cuh"

using namespace fa;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma2_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// K-major SW128 tile: element (row, k), 64 k per 128-byte row, 8-row atoms of 1024 B
__device__ __forceinline__ uint32_t off_sw128(int row, int col) {
  return (row / 8) * 1024 + (row % 8) * 128 + ((((col / 8) ^ (row % 8)) & 7) * 16) + (col % 8) * 2;
}

// test: 0 = M128 K-major B, 1 = M256 K-major B, 2 = M256 MN-major B, 3 = M256 TS (A from TMEM) K-major B,
//       4 = M128 MN-major A (64 M per CTA) + MN-major B
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe(float* out, int test, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  uint8_t* sA = smem;             // 16 KB
  uint8_t* sB = smem + 16384;     // 16 KB
  const uint32_t bar = base + 32768, tptr = base + 32768 + 16;
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rowsA = (test == 0 || test == 4) ? 64 : 128;
  for (int idx = tid; idx < 16384 / 2; idx += 128) {
    reinterpret_cast<uint16_t*>(sA)[idx] = 0;
    reinterpret_cast<uint16_t*>(sB)[idx] = 0;
  }
  __syncthreads();
  // A_r[i][k] = (i % 64 == k) * (rank + 1) * (i >= 64 ? 4 : 1)
  for (int idx = tid; idx < rowsA * 64; idx += 128) {
    const int i = idx / 64, k = idx % 64;
    const float val = (i % 64 == k) ? float((rank + 1) * (i >= 64 ? 4 : 1)) : 0.f;
    if (test == 4) *reinterpret_cast<__nv_bfloat16*>(sA + off_sw128(k, i)) = __float2bfloat16(val);       // MN-major: row = k, col = m
    else *reinterpret_cast<__nv_bfloat16*>(sA + off_sw128(i, k)) = __float2bfloat16(val);
  }
  // B (this CTA's half): local n index nn in [0, 64): global n = 64 * rank + nn.   value = mode ? k + 1 : n + 1
  for (int idx = tid; idx < 64 * 64; idx += 128) {
    const int nn = idx / 64, k = idx % 64;
    const float val = mode ? float(k + 1) : float(64 * rank + nn + 1);
    if (test == 2 || test == 4) *reinterpret_cast<__nv_bfloat16*>(sB + off_sw128(k, nn)) = __float2bfloat16(val);  // MN-major
    else *reinterpret_cast<__nv_bfloat16*>(sB + off_sw128(nn, k)) = __float2bfloat16(val);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  __syncthreads();
  if (warp == 0) {
    tmem_alloc2(tptr, 256);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 32768 + 16);
  // zero my TMEM lanes, columns [0, 256)
  {
    uint32_t z[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) z[c] = 0;
    for (int cc = 0; cc < 256; cc += 32) tmem_st_x32(tmem + (uint32_t(warp * 32) << 16) + cc, z);
    tmem_wait_st();
  }
  if (test == 3) {
    // stage A into TMEM columns [128, 160) as packed bf16 pairs: lane = row i, 32 columns = 64 k
    uint32_t pk[32];
    const int i = warp * 32 + lane;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const float v0 = (i % 64 == 2 * c) ? float((rank + 1) * (i >= 64 ? 4 : 1)) : 0.f;
      const float v1 = (i % 64 == 2 * c + 1) ? float((rank + 1) * (i >= 64 ? 4 : 1)) : 0.f;
      pk[c] = pack_bf16(v0, v1);
    }
    tmem_st_x32(tmem + (uint32_t(warp * 32) << 16) + 128, pk);
    tmem_wait_st();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  if (rank == 0 && warp == 1) {
    if (elect_one()) {
      const uint32_t M = (test == 0 || test == 4) ? 128 : 256;
      const uint32_t a_mn = test == 4 ? 1 : 0, b_mn = (test == 2 || test == 4) ? 1 : 0;
      const uint32_t idesc = make_idesc(1, 1, a_mn, b_mn, M, 128);
      const uint64_t dk = smem_desc_sw128(16, 1024), dmn = smem_desc_sw128(16384, 1024);
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t ad = a_mn ? smem_desc_at(dmn, base + kk * 2048) : smem_desc_at(dk, base + kk * 32);
        const uint64_t bd = b_mn ? smem_desc_at(dmn, base + 16384 + kk * 2048) : smem_desc_at(dk, base + 16384 + kk * 32);
        if (test == 3) mma2_ts(tmem, tmem + 128 + kk * 8, bd, idesc, kk > 0);
        else mma2_ss(tmem, ad, bd, idesc, kk > 0);
      }
      mma2_commit_mc(bar, 3);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0, 1);
  tc_fence_after();
  for (int cc = 0; cc < 128; cc += 32) {
    uint32_t v[32];
    tmem_ld_x32(tmem + (uint32_t(warp * 32) << 16) + cc, v);
    tmem_wait_ld();
    for (int c = 0; c < 32; ++c) out[(rank * 128 + warp * 32 + lane) * 128 + cc + c] = __uint_as_float(v[c]);
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 0) tmem_dealloc2(tmem, 256);
}

int main(int argc, char** argv) {
  float* d;
  const size_t n = 2 * 128 * 128;
  cudaMalloc(&d, n * sizeof(float));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024);
  std::vector<float> h0(n), h1(n);
  for (int test = 0; test < 5; ++test) {
    for (int mode = 0; mode < 2; ++mode) {
      cudaMemset(d, 0, n * sizeof(float));
      probe<<<2, 128, 34 * 1024>>>(d, test, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("test %d mode %d: CUDA error %s\n", test, mode, cudaGetErrorString(e)); return 1; }
      cudaMemcpy((mode ? h1 : h0).data(), d, n * sizeof(float), cudaMemcpyDeviceToHost);
    }
    {
      char name[128];
      snprintf(name, sizeof name, "gpurun_out/pair_probe_t%d.bin", test);
      if (FILE* f = fopen(name, "wb")) { fwrite(h0.data(), 4, n, f); fwrite(h1.data(), 4, n, f); fclose(f); }
    }
    // decode: h0 = w * (n + 1), h1 = w * (i + 1), w = (rank+1) * (row >= 64 ? 4 : 1)
    printf("=== test %d\n", test);
    for (int rank = 0; rank < 2; ++rank) {
      for (int lane = 0; lane < 128; lane += 1) {
        if (!(lane % 16 == 0 || lane % 16 == 15)) continue;
        printf("cta%d lane %3d:", rank, lane);
        for (int col : {0, 1, 31, 32, 63, 64, 65, 127}) {
          const float a = h0[(rank * 128 + lane) * 128 + col], b = h1[(rank * 128 + lane) * 128 + col];
          // w = gcd-ish: try w in {1,2,4,8}
          int w = 0, nn = -1, ii = -1;
          for (int cand : {8, 4, 2, 1}) {
            if (a > 0 && b > 0 && fmodf(a, cand) == 0 && fmodf(b, cand) == 0 && a / cand <= 128 && b / cand <= 64) { w = cand; break; }
          }
          if (w) { nn = int(a / w) - 1; ii = int(b / w) - 1; }
          printf(" c%d=(w%d,i%d,n%d)", col, w, ii, nn);
        }
        printf("\n");
      }
    }
  }
  printf("done\n");
  return 0;
}
