// softmax_rate_probe.cu — how fast can the forward's exponential phase run on one SM, as a function of how many
// softmax warps share a scheduler (sub-partition)?  Separates per-warp LATENCY from pipe THROUGHPUT:
//
//   each thread owns kElems scores in registers and repeats { scale (FFMA2) -> exp2 (MUFU.EX2, a share on the FMA-pipe
//   polynomial) -> row sum (FADD2) -> pack to bf16 (F2FP) -> 16-byte stores to shared memory (stand-in for tcgen05.st) }
//   exactly as fa_fwd_kernel's fast path does, optionally preceded by the row-max pass (FMNMX3).
//
//   blockDim = 128 * k  ->  k warps per scheduler.  Reported: clk per row-block per warp, and clk per row-block per
//   SCHEDULER (= the throughput floor when k warps overlap perfectly).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o experiments/softmax_rate_probe.bin experiments/softmax_rate_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>

#include "../flash_attention_b200/csrc/fa_ptx.cuh"

using namespace fa;

template <int kElems, int kPoly8, bool kWithMax, bool kPipe, int kW>
__global__ void __launch_bounds__(128 * kW, 1) probe(long long* out, float* sink, float scale_in, int reps) {
  extern __shared__ uint4 smem[];
  // scores live in shared memory (stand-in for the TMEM S tile; chunk j of thread t at [j * blockDim + t]: conflict-free),
  // pseudo-random in [-4, 4]; the packed results go to a second region laid out the same way
  float4* src = reinterpret_cast<float4*>(smem);
  uint4* dst = smem + blockDim.x * (kElems / 4);
  for (int j = 0; j < kElems / 4; ++j) {
    float4 v;
    v.x = float((threadIdx.x * 37 + (4 * j + 0) * 101) % 257) * (8.0f / 257.0f) - 4.0f;
    v.y = float((threadIdx.x * 37 + (4 * j + 1) * 101) % 257) * (8.0f / 257.0f) - 4.0f;
    v.z = float((threadIdx.x * 37 + (4 * j + 2) * 101) % 257) * (8.0f / 257.0f) - 4.0f;
    v.w = float((threadIdx.x * 37 + (4 * j + 3) * 101) % 257) * (8.0f / 257.0f) - 4.0f;
    src[j * blockDim.x + threadIdx.x] = v;
  }
  float row_sum = 0.f, m_ref = scale_in;  // scale_in = 0 at run time, unknown to the compiler
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    float s[kElems];
#pragma unroll
    for (int j = 0; j < kElems / 4; ++j) {
      const float4 v = src[j * blockDim.x + threadIdx.x];
      s[4 * j] = v.x; s[4 * j + 1] = v.y; s[4 * j + 2] = v.z; s[4 * j + 3] = v.w;
    }
    float mx = m_ref;
    if (kWithMax) {
      float a = fmax3(s[0], s[1], s[2]), b = fmax3(s[3], s[4], s[5]);
#pragma unroll
      for (int c = 6; c + 3 < kElems; c += 4) {
        a = fmax3(a, s[c], s[c + 1]);
        b = fmax3(b, s[c + 2], s[c + 3]);
      }
      mx = fmaxf(fmaxf(a, b), m_ref) * 1e-30f;  // ~0: keeps the dependency, not the value
    }
    const float scale_log2 = 0.1275f + mx + float(r) * scale_in;
    const float neg_ms = -0.51f - mx;
    const uint64_t sc2 = pack2(scale_log2, scale_log2), ng2 = pack2(neg_ms, neg_ms);
    uint64_t acc_a = pack2(0.f, 0.f), acc_b = pack2(0.f, 0.f);
    const uint64_t zero2 = pack2(0.f, 0.f);
    uint64_t tok_prev = zero2, tok_cur = zero2;
    uint32_t pk[16];
    auto stage_a = [&](int k, uint64_t sc) {
#pragma unroll
      for (int c = 0; c < 16; c += 2) unpack2(fma2(pack2(s[16 * k + c], s[16 * k + c + 1]), sc, ng2), s[16 * k + c], s[16 * k + c + 1]);
    };
    if (kPipe) stage_a(0, sc2);
#pragma unroll
    for (int k = 0; k < kElems / 16; ++k) {
      if (kPipe) { if (k + 1 < kElems / 16) stage_a(k + 1, k == 0 ? sc2 : fma2(tok_prev, zero2, sc2)); }
      else stage_a(k, sc2);
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        const int col = 16 * k + c;
        float e0, e1;
        if ((c >> 1) < 8 - kPoly8) { e0 = ex2(s[col]); e1 = ex2(s[col + 1]); }
        else ex2_poly2_nz(pack2(s[col], s[col + 1]), e0, e1);
        const uint64_t e2 = pack2(e0, e1);
        if (c & 2) acc_b = add2(acc_b, e2); else acc_a = add2(acc_a, e2);
        if (c == 10) { tok_prev = tok_cur; tok_cur = e2; }
        pk[(col & 31) / 2] = pack_bf16(e0, e1);
      }
      if (k & 1) {
#pragma unroll
        for (int x = 0; x < 4; ++x) dst[((k >> 1) * 4 + x) * blockDim.x + threadIdx.x] = make_uint4(pk[4 * x], pk[4 * x + 1], pk[4 * x + 2], pk[4 * x + 3]);
      }
    }
    float a0, a1, b0, b1;
    unpack2(acc_a, a0, a1);
    unpack2(acc_b, b0, b1);
    row_sum += (a0 + a1) + (b0 + b1);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = row_sum;
}

template <int kElems, int kPoly8, bool kWithMax, bool kPipe>
void run(const char* name) {
  long long* d;
  float* sink;
  cudaMalloc(&d, 8);
  cudaMalloc(&sink, 4096 * 4);
  const int reps = 2000;
  for (int k : {1, 2, 4}) {
    if (k == 4 && kElems > 64) continue;   // 512 threads x 128+ live scores do not fit the register file
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
      const size_t smem_bytes = size_t(128 * k) * (4 * kElems + 2 * kElems);
      if (k == 1) { auto kern = probe<kElems, kPoly8, kWithMax, kPipe, 1>; cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); kern<<<1, 128, smem_bytes>>>(d, sink, 0.0f, reps); }
      if (k == 2) { auto kern = probe<kElems, kPoly8, kWithMax, kPipe, 2>; cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); kern<<<1, 256, smem_bytes>>>(d, sink, 0.0f, reps); }
      if (k == 4) { auto kern = probe<kElems, kPoly8, kWithMax, kPipe, 4>; cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); kern<<<1, 512, smem_bytes>>>(d, sink, 0.0f, reps); }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    }
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double per_warp = double(h) / reps;
    printf("%-44s elems/thread=%3d poly8=%d  %d warp(s)/scheduler: %7.1f clk per row-block per warp, %7.1f clk per 128 scores per scheduler\n",
           name, kElems, kPoly8, k, per_warp, per_warp / k * (128.0 / kElems));
  }
}

int main() {
  run<128, 2, false, false>("exp phase (round-1 order)");
  run<128, 2, false, true>("exp phase (software-pipelined)");
  run<128, 0, false, true>("exp phase pipelined, MUFU only");
  run<128, 3, false, true>("exp phase pipelined");
  run<128, 4, false, true>("exp phase pipelined");
  run<128, 8, false, true>("exp phase pipelined, polynomial only");
  run<128, 2, true, true>("max + exp phase pipelined");
  run<64, 2, true, true>("max + exp phase pipelined (half row)");
  run<64, 3, true, true>("max + exp phase pipelined (half row)");
  return 0;
}
