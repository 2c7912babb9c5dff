// fa_api_common.h — shared host-side helpers of libfa_b200.so (error slot, TMA descriptor builder).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <string>

namespace fa {
extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launch_count;
int set_error(const std::string& msg);
// rank-N tiled tensor map, 128-byte (default) or 64-byte swizzle, zero OOB fill; dims innermost first, strides in bytes (dims 1..rank-1)
int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes = 128);
int num_sms();
}  // namespace fa
