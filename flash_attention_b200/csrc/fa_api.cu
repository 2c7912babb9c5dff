// fa_api.cu — C-ABI host layer of libfa_b200.so: argument checks, mask canonicalisation, split-KV
// heuristic, TMA tensor-map construction and kernel dispatch.  Restates the host logic of the
// reference's csrc/flash_attn/flash_api.cpp (set_params_fprop :44-177, set_params_splitkv :317-347,
// num_splits_heuristic :281-315, mha_fwd :368-536, mha_varlen_fwd :538-788, mha_fwd_kvcache :1243-1532)
// for the B200 kernels; no torch types cross this boundary (see include/fa_b200.h).
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/fa_b200.h"
#include "fa_api_common.h"
#include "fa_combine.cuh"
#include "fa_fwd_sm100.cuh"
#include "fa_fwd_wide_sm100.cuh"
#include "fa_kvcache.cuh"

namespace fa {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launch_count{0};

int set_error(const std::string& msg) {
  g_last_error = msg;
  return 1;
}

// ---------------------------------------------------------------- driver entry point for TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// rank-N tiled tensor map, 128-byte swizzle, zero OOB fill. dims/strides innermost first; strides in BYTES
// for dims 1..rank-1.
//
// Descriptors are cached per thread, keyed on everything that defines them (base pointer, dtype, rank, dims, strides, box,
// swizzle): a training loop calls with the same allocator-recycled buffers step after step, and a forward + backward
// pair encodes 3 + 10 maps -- the driver call costs more than the rest of the host path (tools/probe_overhead.py).
struct TmapKey {
  const void* base;
  uint64_t dims[5], strides[4];
  uint32_t box[5];
  int32_t dt, rank, swizzle, pad;
};
struct TmapSlot {
  TmapKey key;
  CUtensorMap map;
  bool valid;
};
static inline uint64_t tmap_hash(const TmapKey& k) {
  const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
  return h;
}
static bool tmap_cache_enabled() {
  static int on = [] { const char* e = getenv("FA_B200_TMAP_CACHE"); return e ? atoi(e) : 1; }();
  return on != 0;
}

int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  constexpr int kSlots = 256;
  thread_local TmapSlot cache[kSlots];
  TmapKey key;
  memset(&key, 0, sizeof key);
  key.base = base; key.dt = int32_t(dt); key.rank = rank; key.swizzle = swizzle_bytes;
  for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; if (i > 0) key.strides[i - 1] = strides_bytes[i - 1]; }
  TmapSlot& slot = cache[tmap_hash(key) % kSlots];
  if (tmap_cache_enabled() && slot.valid && memcmp(&slot.key, &key, sizeof key) == 0) {
    *out = slot.map;
    return 0;
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i] ? dims[i] : 1;
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  for (int i = 1; i < rank; ++i) {
    if (gstr[i - 1] % 16 != 0) return set_error("tensor strides must be multiples of 16 bytes (TMA)");
  }
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) return set_error("tensor base pointers must be 16-byte aligned (TMA)");
  CUresult r = fn(out, dt, rank, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf,
             "cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u,%u]", int(r),
             rank, (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2],
             (unsigned long long)(rank > 3 ? gdim[3] : 0), (unsigned long long)(rank > 4 ? gdim[4] : 0), bdim[0],
             bdim[1], bdim[2], rank > 3 ? bdim[3] : 0, rank > 4 ? bdim[4] : 0);
    return set_error(buf);
  }
  slot.key = key;
  slot.map = *out;
  slot.valid = true;
  return 0;
}

int num_sms() {
  static std::atomic<int> sms[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;   // planner queries on a CPU-only host
  int n = sms[dev].load(std::memory_order_relaxed);
  if (!n) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

static inline bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Fewest splits that reach >= 85 % of the best achievable wave efficiency; 1 if the grid already
// covers >= 80 % of the SMs (restates flash_api.cpp:281-315 for a 1-CTA-per-SM kernel).
int splits_heuristic(int ctas, int sms, int n_blocks, int max_splits) {
  if (ctas >= 0.8f * sms) return 1;
  max_splits = std::min(std::min(max_splits, sms), n_blocks);
  if (max_splits <= 1) return 1;
  float best = 0.f;
  auto eligible = [&](int s) { return s == 1 || cdiv(n_blocks, s) != cdiv(n_blocks, s - 1); };
  auto eff_of = [&](int s) {
    float waves = float(ctas) * s / sms;
    return waves / std::ceil(waves);
  };
  for (int s = 1; s <= max_splits; ++s)
    if (eligible(s)) best = std::max(best, eff_of(s));
  for (int s = 1; s <= max_splits; ++s)
    if (eligible(s) && eff_of(s) >= 0.85f * best) return s;
  return 1;
}

enum FwdMode { kDense = 0, kVarlen = 1, kKvcache = 2 };

#ifndef FA_B200_WIDE_DEFAULT
#define FA_B200_WIDE_DEFAULT 0
#endif

struct FwdPlan {
  int wide;      // 1: the wide-S kernel (one 128-row query tile per CTA, 256-key block pairs)
  int tiles;     // query tiles per CTA: 2 = throughput config (1 CTA/SM), 1 = latency config (2 CTAs/SM)
  int kD;        // padded head dim of the kernel (64 or 128)
  int gp;        // packed heads
  int num_m_blocks;
  int num_splits;
  int has_left, has_right, wl, wr;
  int dropout;
  size_t ws_bytes;
  size_t sched_off;   // byte offset of the dynamic scheduler's work counter inside the workspace ...
  size_t sched_bytes; // ... and its size (0: none requested); the last sched_bytes of ws_bytes, optional for the caller
};

static int plan_fwd(const fa_b200_fwd_args& a, FwdMode mode, FwdPlan* pl) {
  if (a.dtype != FA_B200_FP16 && a.dtype != FA_B200_BF16 && a.dtype != FA_B200_FP8_E4M3)
    return set_error("FlashAttention only support fp16, bf16 and fp8 e4m3 (forward) data type");
  if (a.b <= 0) return set_error("batch size must be positive");
  if (a.d <= 0 || a.d > 256) return set_error("FlashAttention forward only supports head dimension at most 256");
  if (a.d % 8 != 0) return set_error("head_size should be a multiple of 8");
  if (a.d > 128 && a.dtype == FA_B200_FP8_E4M3) return set_error("fp8 forward supports head dimension at most 128");
  if (a.h <= 0 || a.h_k <= 0 || a.h % a.h_k != 0)
    return set_error("Number of heads in key/value must divide number of heads in query");
  if (a.softcap < 0.f) return set_error("softcap must be non-negative");
  // packed queries; keys are packed too (cu_seqlens_k) or batched / paged with per-batch lengths (FA3: hopper/flash_api.cpp:742-760)
  if (mode == kVarlen && !a.cu_seqlens_q) return set_error("cu_seqlens_q is required");
  if (mode == kVarlen && !a.cu_seqlens_k && !a.block_table && !a.seqused_k)
    return set_error("cu_seqlens_k (packed keys) or seqused_k / block_table (batched or paged keys) is required");
  if (a.cache_batch_idx && mode != kKvcache) return set_error("cache_batch_idx is a kvcache-only argument");
  if (a.leftpad_k && mode == kDense) return set_error("leftpad_k is a varlen / kvcache argument");
  if (a.block_table && (a.cache_batch_idx || a.leftpad_k))
    return set_error("We don't support Paged KV and leftpad_k / cache_batch_idx running at the same time yet");
  if (a.block_table) {
    if (a.page_block_size <= 0 || a.page_block_size % 128 != 0)
      return set_error("Paged KV cache block size must be divisible by 128");
    if (a.num_pages <= 0) return set_error("num_pages must be positive for a paged KV cache");
  }
  pl->kD = a.d > 128 ? 256 : ((a.d <= 64 && a.dtype != FA_B200_FP8_E4M3) ? 64 : 128);  // fp8 rows: one 128-byte panel
  pl->dropout = a.p_dropout > 0.f;
  if (a.p_dropout < 0.f || a.p_dropout >= 1.f) return set_error("p_dropout must be in [0, 1)");
  if (pl->dropout) {
    if (a.softcap > 0.f) return set_error("Softcapping does not support dropout for now");  // flash_api.cpp:391
    if (mode == kKvcache || a.block_table) return set_error("dropout is not supported on the kvcache / paged path");
    if (a.dtype == FA_B200_FP8_E4M3) return set_error("dropout is not supported for fp8 inputs");
    if (a.num_splits > 1) return set_error("dropout is not supported with num_splits > 1");
    if (a.pack_gqa == 1) return set_error("dropout is not supported with pack_gqa");
  }
  if (a.s_dmask && (mode == kKvcache || a.dtype == FA_B200_FP8_E4M3))
    return set_error("return_softmax is only supported on the fp16 / bf16 dense and varlen forward");
  if (a.dtype == FA_B200_FP8_E4M3 && a.d % 16 != 0) return set_error("fp8 head_size should be a multiple of 16");
  // window canonicalisation (flash_api.cpp:414-419, set_params_fprop :150-164)
  int wl = a.window_size_left, wr = a.window_size_right;
  int causal = a.is_causal;
  if (wl >= a.seqlen_k) wl = -1;
  if (wr >= a.seqlen_k) wr = -1;
  if (a.seqlen_q == 1 && !a.alibi_slopes && mode != kVarlen) causal = 0;  // causal=true == causal=false for sq=1
  if (causal) wr = 0;
  pl->has_left = wl >= 0;
  pl->has_right = wr >= 0;
  pl->wl = wl;
  pl->wr = wr;
  // Tile configuration: short work items (decode, few K/V blocks per CTA) take the 1-tile kernel, of which two
  // CTAs fit on an SM and overlap each other's prologue / softmax / epilogue; long streams take the 2-tile kernel
  // (half the K/V shared-memory traffic per query row).  FA_B200_FWD_TILES=1|2 overrides (experiments).
  const int g = a.h / a.h_k;
  {
    int est_blocks = cdiv(a.seqlen_k, 128);
    if (pl->has_right && !pl->has_left && a.seqlen_q == a.seqlen_k) est_blocks = (est_blocks + 1) / 2;
    if (pl->has_left) est_blocks = std::min(est_blocks, cdiv(wl + std::max(wr, 0) + 256, 128));
    // measured (profiles/r01_v3_sweep_tiles.txt): causal forward is faster with 1-tile CTAs up to s = 16k
    // (est_blocks <= 64), non-causal only up to s = 1k (est_blocks <= 8)
    // round 2: with the persistent loop (prologue / epilogue overlapped, per-tile block ranges) the 2-tile kernel also
    // wins on long causal problems; FA_B200_FWD_TILE_LIMIT_CAUSAL overrides the crossover for experiments
    // round 2, final build (profiles/r02_z_short_knobs.log): the crossover sits lower than that -- at 8 blocks per item the 2-tile
    // kernel is 5 % faster both non-causal (s = 1024: 933 vs 889 TFLOP/s) and causal (s = 2048: 835 vs 793); at 4 they are equal
    static const int causal_limit = [] { const char* e = getenv("FA_B200_FWD_TILE_LIMIT_CAUSAL"); return e ? atoi(e) : 4; }();
    const int limit = pl->has_right ? causal_limit : 4;
    pl->tiles = (a.seqlen_q * (g > 1 && is_pow2(g) ? g : 1) <= 128 || est_blocks <= limit) ? 1 : 2;
    if (const char* e = getenv("FA_B200_FWD_TILES")) { if (e[0] == '1') pl->tiles = 1; if (e[0] == '2') pl->tiles = 2; }
    if (pl->kD == 256) pl->tiles = 1;  // S (128) + O (256) TMEM columns: one tile per CTA (forward only)
    if (pl->dropout) pl->tiles = 1;    // the dropout kernels are only instantiated for the 1-tile configuration
  }
  // Wide-S kernel (fa_fwd_wide_sm100.cuh: S = Q K^T 256 keys at a time, one query tile per CTA): dense fp16 / bf16, head
  // dim <= 128 in the 128 configuration, plain / causal / local masks only, long key streams.  FA_B200_FWD_WIDE=0/1 forces.
  pl->wide = 0;
  {
    const bool eligible = mode == kDense && a.dtype != FA_B200_FP8_E4M3 && pl->kD == 128 && !pl->dropout && a.softcap == 0.f &&
                          !a.alibi_slopes && !a.seqused_k && !a.seqused_q && !a.debug_s && a.num_splits <= 1 && a.pack_gqa != 1 &&
                          a.seqlen_q > 128;
    int wide_env = -1;   // (read per call: the tests switch it)
    if (const char* e = getenv("FA_B200_FWD_WIDE")) wide_env = atoi(e);
    int est_blocks = cdiv(a.seqlen_k, 128);
    if (pl->has_left) est_blocks = std::min(est_blocks, cdiv(wl + std::max(wr, 0) + 256, 128));
    if (eligible && (wide_env == 1 || (wide_env < 0 && FA_B200_WIDE_DEFAULT && est_blocks >= 16))) {
      pl->wide = 1;
      pl->tiles = 1;
    }
  }
  const int rows_cta = 128 * pl->tiles;
  // GQA head packing: fold the g query heads of one KV head into the M dimension when a CTA's worth of query
  // positions would otherwise be mostly empty (decode / short queries).
  int gp = 1;
  if (a.pack_gqa == 1 && g > 1) {
    if (!is_pow2(g) || g > 128) return set_error("pack_gqa needs a power-of-two group size <= 128");
    gp = g;
  } else if (a.pack_gqa == 0 && g > 1 && is_pow2(g) && g <= 128 && !pl->dropout && !pl->wide) {
    const int sq = a.seqlen_q;
    const double eff_unpacked = double(sq) / (cdiv(sq, rows_cta) * rows_cta);
    const double eff_packed = double(sq) * g / (cdiv(sq * g, rows_cta) * rows_cta);
    if (eff_packed > 1.2 * eff_unpacked) gp = g;
  }
  pl->gp = gp;
  pl->num_m_blocks = cdiv(a.seqlen_q * gp, rows_cta);
  // split-KV (flash_api.cpp:317-347): heuristic when num_splits == 0; never for packed-varlen q
  int splits = a.num_splits;
  if (splits < 0) return set_error("num_splits must be non-negative");
  if (pl->dropout || pl->wide) {
    splits = 1;  // flash_api.cpp:331: the reference only splits when p_dropout == 0
  } else if (mode == kVarlen) {
    if (a.block_table && splits > 1) return set_error("num_splits > 1 is not supported for varlen paged KV");
    splits = 1;
  } else if (splits == 0) {
    const int n_blocks = cdiv(a.seqlen_k, 128);
    const int ctas = pl->num_m_blocks * (a.h / gp) * a.b;
    const int slots = num_sms() * (pl->tiles == 1 ? 2 : 1);  // resident CTAs per wave
    splits = splits_heuristic(ctas, slots, n_blocks, 128);
    // B200 refinement for long-KV decode: the reference rule returns 1 as soon as the grid covers 80 % of the
    // SMs, but e.g. 256 CTAs on 148 SMs run 2 waves at 86 % occupancy.  When every split still streams >= 32
    // blocks (partials + combine are then noise) take the smallest split count that lifts wave efficiency by
    // > 5 % (C5: 1 -> 4 splits, 5662 -> 6132 GB/s measured).
    if (splits == 1 && mode == kKvcache && n_blocks >= 64) {
      auto eff = [&](int s) { float w = float(ctas) * s / slots; return w / std::ceil(w); };
      float best = eff(1);
      for (int s = 2; s <= 8 && n_blocks / s >= 32; ++s)
        if (eff(s) > best * 1.05f) { best = eff(s); splits = s; }
    }
  }
  if (splits > 128) return set_error("num_splits > 128 not supported");
  pl->num_splits = splits;
  pl->ws_bytes = 0;
  if (splits > 1) {
    const size_t rows = size_t(splits) * a.b * a.h * a.seqlen_q;
    pl->ws_bytes = rows * pl->kD * sizeof(float) + rows * sizeof(float);
    pl->ws_bytes = (pl->ws_bytes + 255) & ~size_t(255);
  }
  // packed varlen: the (b + 1)-entry m-block prefix sum the persistent kernel enumerates its work items with
  if (a.cu_seqlens_q) pl->ws_bytes = ((size_t(a.b) + 1) * sizeof(int) + 255) & ~size_t(255);
  // work counter of the dynamic scheduler: asked for when there can be more work items than resident CTAs (a B200 has
  // 148 SMs; the launch decides with the real SM count and falls back to the static deal if the caller gave no scratch)
  pl->sched_off = pl->sched_bytes = 0;
  {
    const int64_t items = int64_t(pl->num_m_blocks) * (a.h / pl->gp) * a.b * splits;
    if (items > 148 || a.cu_seqlens_q) {
      pl->sched_off = pl->ws_bytes;
      pl->sched_bytes = 256;
      pl->ws_bytes += pl->sched_bytes;
    }
  }
  return 0;
}

template <typename T, int kD, int kTiles, bool kDropout = false>
static int launch_fwd_kernel_t(const FwdKernelParams& kp, dim3 grid, cudaStream_t stream) {
  using Cfg = FwdCfg<kD, EltInfo<T>::kBytes, kTiles>;
  static PerDeviceOnce configured;
  auto kern = fa_fwd_kernel<T, kD, kTiles, kDropout>;
  if (cudaError_t e = configured.run([&] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes); });
      e != cudaSuccess)
    return set_error(std::string("cudaFuncSetAttribute(fwd): ") + cudaGetErrorString(e));
  kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(kp);
  g_launch_count++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(std::string("fa_fwd_kernel launch: ") + cudaGetErrorString(e));
  return 0;
}

template <typename T, int kD>
static int launch_fwd_kernel(const FwdKernelParams& kp, dim3 grid, cudaStream_t stream, int tiles, bool dropout = false) {
  if constexpr (!EltInfo<T>::kFp8) {
    if (dropout) return launch_fwd_kernel_t<T, kD, 1, true>(kp, grid, stream);
  }
  return tiles == 1 ? launch_fwd_kernel_t<T, kD, 1>(kp, grid, stream) : launch_fwd_kernel_t<T, kD, 2>(kp, grid, stream);
}

template <typename T>
static int launch_fwd_wide(const FwdKernelParams& kp, dim3 grid, cudaStream_t stream) {
  using Cfg = FwdWideCfg<FA_WIDE_GROUPS>;
  static PerDeviceOnce configured;
  auto kern = fa_fwd_wide_kernel<T, FA_WIDE_GROUPS>;
  if (cudaError_t e = configured.run([&] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes); });
      e != cudaSuccess)
    return set_error(std::string("cudaFuncSetAttribute(fwd wide): ") + cudaGetErrorString(e));
  kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(kp);
  g_launch_count++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(std::string("fa_fwd_wide_kernel launch: ") + cudaGetErrorString(e));
  return 0;
}

static int run_fwd(const fa_b200_fwd_args& a, FwdMode mode, cudaStream_t stream) {
  FwdPlan pl;
  if (!a.q || !a.k || !a.v || !a.out || !a.softmax_lse) return set_error("q, k, v, out, softmax_lse must be non-null");
  if (int rc = plan_fwd(a, mode, &pl)) return rc;
  const size_t ws_required = pl.ws_bytes - pl.sched_bytes;   // the scheduler's counter is optional
  if (ws_required > 0 && (!a.workspace || a.workspace_bytes < ws_required))
    return set_error(pl.num_splits > 1 ? "workspace too small for split-KV (query fa_b200_fwd_workspace_bytes)"
                                       : "workspace too small for the varlen work list (query fa_b200_fwd_workspace_bytes)");
  if (a.seqlen_q == 0 || (mode == kVarlen && a.total_q == 0)) return 0;

  const bool fp8 = a.dtype == FA_B200_FP8_E4M3;
  const int esz = fp8 ? 1 : 2;
  const uint32_t panel = 128 / esz;  // elements per 128-byte swizzle row
  const CUtensorMapDataType dt = fp8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                     : (a.dtype == FA_B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  const int g = a.h / a.h_k, gp = pl.gp;
  FwdKernelParams kp;
  memset(&kp, 0, sizeof kp);
  {
    // Q: (d, gp, rows, h/gp, b)
    const bool varlen = a.cu_seqlens_q != nullptr;
    uint64_t dims[5] = {uint64_t(a.d), uint64_t(gp), uint64_t(varlen ? a.total_q : a.seqlen_q), uint64_t(a.h / gp),
                        uint64_t(varlen ? 1 : a.b)};
    uint64_t str[4] = {uint64_t(a.q_head_stride) * esz, uint64_t(a.q_row_stride) * esz,
                       uint64_t(a.q_head_stride) * gp * esz,
                       uint64_t(varlen ? int64_t(a.total_q) * a.q_row_stride : a.q_batch_stride) * esz};
    if (dims[4] == 1) str[3] = dims[2] * str[1];  // any valid value
    uint32_t box[5] = {panel, uint32_t(gp), uint32_t(128 / gp), 1, 1};
    if (int rc = make_tmap(&kp.tm_q, dt, 5, a.q, dims, str, box)) return rc;
  }
  for (int which = 0; which < 2; ++which) {
    const void* base = which == 0 ? a.k : a.v;
    const int64_t bs = which == 0 ? a.k_batch_stride : a.v_batch_stride;
    const int64_t rs = which == 0 ? a.k_row_stride : a.v_row_stride;
    const int64_t hs = which == 0 ? a.k_head_stride : a.v_head_stride;
    uint64_t rows, nb;
    if (a.block_table) { rows = a.page_block_size; nb = a.num_pages; }
    else if (a.cu_seqlens_k) { rows = a.total_k; nb = 1; }
    else { rows = a.seqlen_k; nb = a.b_kvcache > 0 ? a.b_kvcache : a.b; }
    uint64_t dims[4] = {uint64_t(a.d), rows, uint64_t(a.h_k), nb};
    uint64_t str[3] = {uint64_t(rs) * esz, uint64_t(hs) * esz, uint64_t(bs) * esz};
    if (nb == 1) str[2] = rows * str[0];
    uint32_t box[4] = {panel, 128, 1, 1};
    if (int rc = make_tmap(which == 0 ? &kp.tm_k : &kp.tm_v, dt, 4, base, dims, str, box)) return rc;
    if (pl.wide) {
      kp.q_raw = a.q; kp.k_raw = a.k; kp.v_raw = a.v;
      kp.q_bs = a.q_batch_stride; kp.q_rs = a.q_row_stride; kp.q_hs = a.q_head_stride;
      kp.k_bs = a.k_batch_stride; kp.k_rs = a.k_row_stride; kp.k_hs = a.k_head_stride;
      kp.v_bs = a.v_batch_stride; kp.v_rs = a.v_row_stride; kp.v_hs = a.v_head_stride;
      uint32_t box2[4] = {panel, 256, 1, 1};
      if (int rc = make_tmap(which == 0 ? &kp.tm_k256 : &kp.tm_v256, dt, 4, base, dims, str, box2)) return rc;
    }
  }
  kp.o = a.out;
  kp.lse = a.softmax_lse;
  kp.o_batch_stride = a.o_batch_stride;
  kp.o_row_stride = a.o_row_stride;
  kp.o_head_stride = a.o_head_stride;
  kp.cu_seqlens_q = a.cu_seqlens_q;
  kp.cu_seqlens_k = a.cu_seqlens_k;
  kp.seqused_k = a.seqused_k;
  kp.seqused_q = a.seqused_q;
  kp.cache_batch_idx = a.cache_batch_idx;
  kp.leftpad_k = a.leftpad_k;
  kp.block_table = a.block_table;
  kp.block_table_batch_stride = a.block_table_batch_stride;
  kp.alibi_slopes = a.alibi_slopes;
  kp.alibi_batch_stride = a.alibi_slopes_batch_stride;
  kp.debug_s = a.debug_s;
  kp.q_descale = a.q_descale; kp.k_descale = a.k_descale; kp.v_descale = a.v_descale;
  kp.q_descale_bs = a.q_descale_batch_stride; kp.k_descale_bs = a.k_descale_batch_stride; kp.v_descale_bs = a.v_descale_batch_stride;
  kp.b = a.b; kp.h = a.h; kp.h_k = a.h_k; kp.sq = a.seqlen_q; kp.sk = a.seqlen_k; kp.d = a.d;
  kp.g = g; kp.gp = gp; kp.total_q = a.total_q;
  kp.page_size = a.block_table ? a.page_block_size : 0;
  kp.num_splits = pl.num_splits;
  kp.has_left = pl.has_left; kp.has_right = pl.has_right; kp.wl = pl.wl; kp.wr = pl.wr;
  kp.reverse_m = pl.has_right ? 1 : 0;   // causal: the last m-blocks see the most keys -> dispatched first
  if (a.softcap > 0.f) {  // s' = tanh(s * scale / softcap) * softcap  (flash_api.cpp:130-134)
    kp.softcap_pre = a.softmax_scale / a.softcap;
    kp.scale = a.softcap;
  } else {
    kp.softcap_pre = 0.f;
    kp.scale = a.softmax_scale;
  }
  kp.scale_log2 = kp.scale * float(M_LOG2E);
  if (pl.dropout) kp.drop = make_dropout_params(a.p_dropout, a.philox_seed, a.philox_offset);
  if (pl.num_splits > 1) {
    const size_t rows = size_t(pl.num_splits) * a.b * a.h * a.seqlen_q;
    kp.o_accum = reinterpret_cast<float*>(a.workspace);
    kp.lse_accum = kp.o_accum + rows * pl.kD;
  }
  // Persistent grid: one CTA per SM (two for the 1-tile configuration), fewer if there is less work, `sm_margin` SMs left
  // free (FA3).  Work items are enumerated inside the kernel (fwd_decode_work): heavy m-blocks first within groups of head
  // slots whose K/V fit L2 together; packed varlen batches through the m-block prefix sum written by the prepare kernel.
  kp.num_m_blocks = pl.num_m_blocks;
  const int nslots = a.h / gp;
  const int64_t hz_total = int64_t(nslots) * a.b * pl.num_splits;
  {
    const double kv_bytes = 2.0 * double(a.block_table ? a.seqlen_k : a.seqlen_k) * a.d * esz;   // K + V of one kv head
    int kv_heads = int(std::max(1.0, std::floor(32.0 * 1024 * 1024 / std::max(kv_bytes, 1.0))));
    int64_t grp = int64_t(kv_heads) * (g / gp);   // head slots that share those kv heads
    if (const char* e = getenv("FA_B200_SCHED_GROUP")) grp = atoi(e);
    grp = std::max<int64_t>(1, std::min<int64_t>(grp, hz_total));
    kp.sched_group = int(grp);
  }
  const int64_t total_items = int64_t(pl.num_m_blocks) * hz_total;
  if (total_items > INT32_MAX) return set_error("too many work items");
  kp.total_items = int(total_items);
  const int ctas_per_sm = pl.tiles == 1 && pl.kD != 256 && !pl.wide ? 2 : 1;
  int sms = num_sms() - std::max(0, a.sm_margin);
  if (sms < 1) sms = 1;
  int64_t ctas = int64_t(sms) * ctas_per_sm;
  // Dynamic draws pay for themselves when the items are long (>= 24 key blocks; causal 8k forward +6 %, C3 +6 %, same lease,
  // profiles/r02_sched_dynamic_abab.log); with 1-8 block items (causal 1k) the draw latency shows and the static deal,
  // which has many rounds to even out, is 10 % faster.  FA_B200_SCHED_DYNAMIC=0 / 2 forces static / dynamic.
  int sched_mode = 1;   // (read per call: the tests switch it)
  if (const char* e = getenv("FA_B200_SCHED_DYNAMIC")) sched_mode = atoi(e);
  const bool sched_dynamic = sched_mode == 2 || (sched_mode == 1 && a.seqlen_k >= 3072);
  int* sched_counter = (sched_dynamic && pl.sched_bytes && a.workspace && a.workspace_bytes >= pl.sched_off + sizeof(int))
                           ? reinterpret_cast<int*>(static_cast<char*>(a.workspace) + pl.sched_off) : nullptr;
  if (a.cu_seqlens_q) {
    int* prefix = reinterpret_cast<int*>(a.workspace);
    if (sched_counter) kp.sched_counter = sched_counter;   // zeroed by the prepare kernel
    fa_fwd_varlen_prepare<<<1, 256, 0, stream>>>(a.cu_seqlens_q, a.seqused_q, a.b, gp, 128 * pl.tiles, prefix, sched_counter);
    g_launch_count++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(std::string("fa_fwd_varlen_prepare launch: ") + cudaGetErrorString(e));
    kp.varlen_prefix = prefix;
    ctas = std::min<int64_t>(ctas, std::max<int64_t>(1, total_items));   // upper bound: max_seqlen m-blocks per batch
  } else {
    ctas = std::min<int64_t>(ctas, total_items);
    if (sched_counter && total_items > ctas) {
      if (cudaError_t e = cudaMemsetAsync(sched_counter, 0, sizeof(int), stream); e != cudaSuccess)
        return set_error(std::string("cudaMemsetAsync(work counter): ") + cudaGetErrorString(e));
      kp.sched_counter = sched_counter;
    }
  }
  dim3 grid(unsigned(ctas), 1, 1);
  int rc;
  if (pl.wide) {
    rc = a.dtype == FA_B200_BF16 ? launch_fwd_wide<__nv_bfloat16>(kp, grid, stream) : launch_fwd_wide<__half>(kp, grid, stream);
  } else if (fp8) {
    rc = launch_fwd_kernel<fp8_e4m3, 128>(kp, grid, stream, pl.tiles);
  } else if (a.dtype == FA_B200_BF16) {
    rc = pl.kD == 64 ? launch_fwd_kernel<__nv_bfloat16, 64>(kp, grid, stream, pl.tiles, pl.dropout)
       : pl.kD == 128 ? launch_fwd_kernel<__nv_bfloat16, 128>(kp, grid, stream, pl.tiles, pl.dropout)
       : pl.dropout ? launch_fwd_kernel_t<__nv_bfloat16, 256, 1, true>(kp, grid, stream)
                    : launch_fwd_kernel_t<__nv_bfloat16, 256, 1>(kp, grid, stream);
  } else {
    rc = pl.kD == 64 ? launch_fwd_kernel<__half, 64>(kp, grid, stream, pl.tiles, pl.dropout)
       : pl.kD == 128 ? launch_fwd_kernel<__half, 128>(kp, grid, stream, pl.tiles, pl.dropout)
       : pl.dropout ? launch_fwd_kernel_t<__half, 256, 1, true>(kp, grid, stream)
                    : launch_fwd_kernel_t<__half, 256, 1>(kp, grid, stream);
  }
  if (rc) return rc;
  if (pl.num_splits > 1) {
    CombineParams cp;
    cp.o_partial = kp.o_accum;
    cp.lse_partial = kp.lse_accum;
    cp.out = a.out;
    cp.lse = a.softmax_lse;
    cp.rows = int64_t(a.b) * a.h * a.seqlen_q;
    cp.splits = pl.num_splits;
    cp.d = a.d;
    cp.d_part = pl.kD;
    cp.h = a.h;
    cp.sq = a.seqlen_q;
    cp.obs = a.o_batch_stride; cp.ors = a.o_row_stride; cp.ohs = a.o_head_stride;
    const unsigned blocks = unsigned((cp.rows + 7) / 8);
    if (a.dtype == FA_B200_FP16) fa_combine_kernel<__half><<<blocks, 256, 0, stream>>>(cp);
    else fa_combine_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(cp);
    g_launch_count++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(std::string("fa_combine_kernel launch: ") + cudaGetErrorString(e));
  }
  if (a.s_dmask) {
    DmaskParams dm;
    memset(&dm, 0, sizeof dm);
    dm.q = a.q; dm.k = a.k; dm.lse = a.softmax_lse; dm.s_out = a.s_dmask;
    dm.q_batch_stride = a.q_batch_stride; dm.q_row_stride = a.q_row_stride; dm.q_head_stride = a.q_head_stride;
    dm.k_batch_stride = a.k_batch_stride; dm.k_row_stride = a.k_row_stride; dm.k_head_stride = a.k_head_stride;
    dm.cu_seqlens_q = a.cu_seqlens_q; dm.cu_seqlens_k = a.cu_seqlens_k;
    dm.alibi_slopes = a.alibi_slopes; dm.alibi_batch_stride = a.alibi_slopes_batch_stride;
    dm.b = a.b; dm.h = a.h; dm.h_k = a.h_k; dm.sq = a.seqlen_q; dm.sk = a.seqlen_k; dm.d = a.d;
    dm.sq_r = (a.seqlen_q + 127) / 128 * 128; dm.sk_r = (a.seqlen_k + 127) / 128 * 128; dm.total_q = a.total_q;
    dm.has_left = pl.has_left; dm.has_right = pl.has_right; dm.wl = pl.wl; dm.wr = pl.wr;
    dm.scale = a.softmax_scale;
    dm.drop = make_dropout_params(a.p_dropout, a.philox_seed, a.philox_offset);
    dim3 dgrid((dm.sq_r + 7) / 8, a.h, a.b);
    if (a.dtype == FA_B200_FP16) fa_dmask_kernel<__half><<<dgrid, 256, 0, stream>>>(dm);
    else fa_dmask_kernel<__nv_bfloat16><<<dgrid, 256, 0, stream>>>(dm);
    g_launch_count++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(std::string("fa_dmask_kernel launch: ") + cudaGetErrorString(e));
  }
  return 0;
}

}  // namespace fa

using namespace fa;

extern "C" {

int fa_b200_mha_fwd(const fa_b200_fwd_args* args, fa_b200_stream_t stream) {
  if (!args) return set_error("null args");
  if (args->cu_seqlens_q || args->cu_seqlens_k || args->block_table)
    return set_error("fa_b200_mha_fwd takes dense tensors; use the varlen / kvcache entry points");
  return run_fwd(*args, kDense, static_cast<cudaStream_t>(stream));
}

int fa_b200_mha_varlen_fwd(const fa_b200_fwd_args* args, fa_b200_stream_t stream) {
  if (!args) return set_error("null args");
  return run_fwd(*args, kVarlen, static_cast<cudaStream_t>(stream));
}

int fa_b200_mha_fwd_kvcache(const fa_b200_fwd_args* args, fa_b200_stream_t stream) {
  if (!args) return set_error("null args");
  if (args->cu_seqlens_q || args->cu_seqlens_k) return set_error("kvcache forward takes batched (non-packed) q");
  return run_fwd(*args, kKvcache, static_cast<cudaStream_t>(stream));
}

// which entry point the args are meant for: `entry` says so (the planner queries must agree with the launch: the kvcache
// entry plans differently from the dense one even without cache_seqlens / block_table); 0 = infer from the pointers
static FwdMode mode_of(const fa_b200_fwd_args* a) {
  if (a->entry == FA_B200_ENTRY_DENSE) return kDense;
  if (a->entry == FA_B200_ENTRY_VARLEN) return kVarlen;
  if (a->entry == FA_B200_ENTRY_KVCACHE) return kKvcache;
  return a->cu_seqlens_q ? kVarlen : (a->block_table || a->seqused_k ? kKvcache : kDense);
}

size_t fa_b200_fwd_workspace_bytes(const fa_b200_fwd_args* args) {
  FwdPlan pl;
  if (!args || plan_fwd(*args, mode_of(args), &pl)) return 0;
  return pl.ws_bytes;
}

int fa_b200_fwd_plan(const fa_b200_fwd_args* args, int32_t out[3]) {
  FwdPlan pl;
  if (!args || !out) return set_error("null args");
  if (int rc = plan_fwd(*args, mode_of(args), &pl)) return rc;
  out[0] = pl.num_splits; out[1] = pl.gp; out[2] = pl.num_m_blocks;
  return 0;
}

int fa_b200_fwd_kernel_kind(const fa_b200_fwd_args* args) {
  FwdPlan pl;
  if (!args || plan_fwd(*args, mode_of(args), &pl)) return -1;
  return pl.wide ? 3 : pl.tiles;
}

int fa_b200_fwd_num_splits(const fa_b200_fwd_args* args) {
  FwdPlan pl;
  if (!args || plan_fwd(*args, mode_of(args), &pl)) return -1;
  return pl.num_splits;
}

int fa_b200_fwd_combine(const float* o_partial, const float* lse_partial, void* out, float* lse, int32_t num_splits,
                        int64_t rows, int32_t d, int32_t dtype, fa_b200_stream_t stream) {
  if (!o_partial || !lse_partial || !out) return set_error("null pointer");
  if (d % 4 != 0) return set_error("combine: d must be a multiple of 4");
  if (num_splits < 1) return set_error("combine: num_splits must be >= 1");
  if (rows <= 0) return 0;
  CombineParams cp;
  cp.o_partial = o_partial; cp.lse_partial = lse_partial; cp.out = out; cp.lse = lse;
  cp.rows = rows; cp.splits = num_splits; cp.d = d; cp.d_part = d; cp.h = 1; cp.sq = 1;
  cp.obs = d; cp.ors = 0; cp.ohs = 0;
  const unsigned blocks = unsigned((rows + 7) / 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == FA_B200_BF16) fa_combine_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(cp);
  else if (dtype == FA_B200_FP16) fa_combine_kernel<__half><<<blocks, 256, 0, s>>>(cp);
  else return set_error("combine: output dtype must be fp16 or bf16");
  g_launch_count++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(std::string("fa_combine_kernel launch: ") + cudaGetErrorString(e));
  return 0;
}

int fa_b200_num_splits_heuristic(int batch_nheads_mblocks, int num_sms_, int num_n_blocks, int max_splits) {
  return splits_heuristic(batch_nheads_mblocks, num_sms_, num_n_blocks, max_splits);
}

int fa_b200_kvcache_append(const fa_b200_append_args* a, fa_b200_stream_t stream) {
  if (!a) return set_error("null args");
  if (a->dtype != FA_B200_FP16 && a->dtype != FA_B200_BF16) return set_error("kvcache append supports fp16 and bf16");
  if (!a->cache_seqlens) return set_error("If key is supplied, seqlens_k must also be passed in");
  if (a->knew && (!a->vnew || !a->kcache || !a->vcache)) return set_error("If key is supplied, value must also be passed in");
  if (a->rotary_cos) {
    if (!a->rotary_sin) return set_error("rotary_sin is required with rotary_cos");
    if (a->rotary_dim <= 0 || a->rotary_dim > a->d || a->rotary_dim % 16 != 0)
      return set_error("Only rotary dimensions divisible by 16 are currently supported");
  }
  if (a->block_table && (a->page_block_size <= 0)) return set_error("page_block_size must be positive");
  AppendParams p;
  p.kcache = a->kcache; p.vcache = a->vcache; p.knew = a->knew; p.vnew = a->vnew; p.q = a->q; p.q_out = a->q_out;
  p.kc_bs = a->kc_batch_stride; p.kc_rs = a->kc_row_stride; p.kc_hs = a->kc_head_stride;
  p.vc_bs = a->vc_batch_stride; p.vc_rs = a->vc_row_stride; p.vc_hs = a->vc_head_stride;
  p.kn_bs = a->kn_batch_stride; p.kn_rs = a->kn_row_stride; p.kn_hs = a->kn_head_stride;
  p.vn_bs = a->vn_batch_stride; p.vn_rs = a->vn_row_stride; p.vn_hs = a->vn_head_stride;
  p.q_bs = a->q_batch_stride; p.q_rs = a->q_row_stride; p.q_hs = a->q_head_stride;
  p.qo_bs = a->qo_batch_stride; p.qo_rs = a->qo_row_stride; p.qo_hs = a->qo_head_stride;
  p.cache_seqlens = a->cache_seqlens; p.cache_batch_idx = a->cache_batch_idx; p.leftpad_k = a->leftpad_k;
  p.block_table = a->block_table; p.bt_bs = a->block_table_batch_stride; p.page_size = a->page_block_size;
  p.rot_cos = a->rotary_cos; p.rot_sin = a->rotary_sin; p.rotary_dim = a->rotary_dim;
  p.interleaved = a->is_rotary_interleaved; p.q_pos_per_row = a->q_pos_per_row;
  p.b = a->b; p.s_new = a->knew ? a->seqlen_new : 0; p.sq = (a->rotary_cos && a->q_out) ? a->seqlen_q : 0;
  p.h = a->h; p.h_k = a->h_k; p.d = a->d;
  const int rows = p.s_new > p.sq ? p.s_new : p.sq, heads = a->h > a->h_k ? a->h : a->h_k;
  if (rows == 0 || a->b == 0) return 0;
  dim3 grid(rows, heads, a->b);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->dtype == FA_B200_BF16) fa_kv_append_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(p);
  else fa_kv_append_kernel<__half><<<grid, 128, 0, s>>>(p);
  g_launch_count++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(std::string("fa_kv_append_kernel launch: ") + cudaGetErrorString(e));
  return 0;
}
size_t fa_b200_sizeof_append_args(void) { return sizeof(fa_b200_append_args); }

#if FA_FWD_TS
// timeline-debugging builds only (not declared in include/fa_b200.h)
int fa_b200_debug_read_fwd_ts(unsigned long long* out, int n) {
  return cudaMemcpyFromSymbol(out, fa::g_fwd_ts, size_t(n) * sizeof(unsigned long long)) == cudaSuccess ? 0 : 1;
}
int fa_b200_debug_read_fwd_items(unsigned long long* ts, unsigned int* sm) {
  if (cudaMemcpyFromSymbol(ts, fa::g_fwd_item_ts, sizeof(unsigned long long) * 192 * 64 * 2) != cudaSuccess) return 1;
  return cudaMemcpyFromSymbol(sm, fa::g_fwd_item_sm, sizeof(unsigned int) * 192) == cudaSuccess ? 0 : 1;
}
#endif
int fa_b200_memcpy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t height,
                           int32_t kind, fa_b200_stream_t stream) {
  if (kind != 1 && kind != 2) return set_error("fa_b200_memcpy2d_async: kind must be 1 (host->device) or 2 (device->host)");
  cudaError_t e = cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, height,
                                    kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return set_error(std::string("cudaMemcpy2DAsync: ") + cudaGetErrorString(e));
  return 0;
}
__global__ void fa_write_rng_state_kernel(uint64_t* dst, uint64_t seed, uint64_t offset) { dst[0] = seed; dst[1] = offset; }
int fa_b200_write_rng_state(uint64_t* dst, uint64_t seed, uint64_t offset, fa_b200_stream_t stream) {
  if (!dst) return set_error("null pointer");
  fa_write_rng_state_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(dst, seed, offset);
  g_launch_count++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(std::string("fa_write_rng_state_kernel launch: ") + cudaGetErrorString(e));
  return 0;
}
const char* fa_b200_last_error(void) { return g_last_error.c_str(); }
size_t fa_b200_sizeof_fwd_args(void) { return sizeof(fa_b200_fwd_args); }
size_t fa_b200_sizeof_bwd_args(void) { return sizeof(fa_b200_bwd_args); }
const char* fa_b200_version(void) { return "fa_b200 0.1 (sm_100a)"; }
uint64_t fa_b200_launch_count(void) { return g_launch_count.load(); }

}  // extern "C"
