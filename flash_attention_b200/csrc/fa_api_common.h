// fa_api_common.h — shared host-side helpers of libfa_b200.so (error slot, TMA descriptor builder).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>

namespace fa {
extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launch_count;
int set_error(const std::string& msg);
// rank-N tiled tensor map, 128-byte (default) or 64-byte swizzle, zero OOB fill; dims innermost first, strides in bytes (dims 1..rank-1)
int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes = 128);
int num_sms();   // of the CURRENT device (cached per device)
// One-time per-(kernel, device) setup: cudaFuncSetAttribute is per device / context, so a process that drives several GPUs
// must raise the dynamic shared-memory limit on each of them (a per-process flag left the second device at the 48 KB default).
struct PerDeviceOnce {
  std::atomic<bool> done[64] = {};
  template <typename F>
  cudaError_t run(F&& f) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return f();
    if (done[dev].load(std::memory_order_acquire)) return cudaSuccess;
    cudaError_t e = f();
    if (e == cudaSuccess) done[dev].store(true, std::memory_order_release);
    return e;
  }
};
// Block order.  Default is head-major (all blocks of one head are dispatched together: they share their streamed operands
// through L2, which measured more important than a balanced tail: pure work-major order cost 10 % on causal problems).
// The chunked work-major order keeps chunks of c adjacent blocks of a head together and dispatches heavy chunks of all
// heads first; it buys 0.5-2.5 % on the causal forward but DOUBLES its DRAM reads (406 -> 813 MB at b4 h16 s8192: more heads'
// K/V compete for L2, hit rate 90.8 -> 83.9 %), and nothing on the backward: off by default for both.
// forward: FA_B200_SCHED_CHUNK_FWD=c (m-blocks per chunk) for masked problems, negative: for all problems; 0 = head-major
inline int sched_chunk_fwd(bool masked) {
  static int c = [] { const char* e = getenv("FA_B200_SCHED_CHUNK_FWD"); return e ? atoi(e) : 0; }();
  if (c < 0) return -c;
  return masked ? c : 0;
}
inline int sched_chunk(bool masked) {
  static int c = [] { const char* e = getenv("FA_B200_SCHED_CHUNK"); return e ? atoi(e) : 0; }();
  if (c < 0) return -c;
  return masked ? c : 0;
}
}  // namespace fa

#include "fa_dropout.cuh"
namespace fa {
// p_dropout -> 8-bit keep threshold and rescale factor (reference: flash_api.cpp:113-121)
inline DropoutParams make_dropout_params(float p_dropout, uint64_t seed, uint64_t offset) {
  DropoutParams dp;
  const float p_keep = 1.f - p_dropout;
  dp.seed_lo = uint32_t(seed); dp.seed_hi = uint32_t(seed >> 32);
  dp.off_lo = uint32_t(offset); dp.off_hi = uint32_t(offset >> 32);
  dp.thr = uint32_t(floorf(p_keep * 255.0f));
  dp.rp = 1.f / p_keep;
  return dp;
}
}  // namespace fa
