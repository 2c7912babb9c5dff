// fa_bwd_api.cu — C-ABI host layer for the backward pass (restates mha_bwd / mha_varlen_bwd,
// csrc/flash_attn/flash_api.cpp:800-1241, for the sm_100a kernels in fa_bwd_sm100.cuh).
//
// Differences from the reference host code, by design:
//   * GQA: dK/dV of a KV head are accumulated over its g query heads inside the kernel (TMEM), so there
//     are no (b, sk, h, d) dk_expanded/dv_expanded temporaries and no at::sum_out (flash_api.cpp:935-942,1002-1005).
//   * dq_accum is laid out (b, h, round128(max_sq), kD) so that one TMA box = one 128-row dQ tile.
#include <cstring>

#include "../../include/fa_b200.h"
#include "fa_api_common.h"
#include "fa_bwd256_sm100.cuh"
#include "fa_bwd2_sm100.cuh"
#include "fa_bwd_sm100.cuh"

namespace fa {

static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct BwdPlan {
  int kD, sq_r;
  int64_t ws_rows_head;   // packed varlen layout: rows per head (0 = dense layout)
  int has_left, has_right, wl, wr;
  size_t off_lse, off_dps, off_dq, off_sem, off_sched, total;
};

// head dims 129..256 run fa_bwd256_kernel (fa_bwd256_sm100.cuh); FA_B200_BWD_HD256=0 restores the round-1 behaviour
// of rejecting them
static bool hd256_enabled() {
  static int mode = [] { const char* e = getenv("FA_B200_BWD_HD256"); return e ? atoi(e) : 1; }();
  return mode != 0;
}

static int plan_bwd(const fa_b200_bwd_args& a, bool varlen, BwdPlan* pl) {
  if (a.dtype != FA_B200_FP16 && a.dtype != FA_B200_BF16) return set_error("FlashAttention only support fp16 and bf16 data type");
  if (!a.dout || !a.q || !a.k || !a.v || !a.out || !a.softmax_lse || !a.dq || !a.dk || !a.dv)
    return set_error("dout, q, k, v, out, softmax_lse, dq, dk, dv must be non-null");
  if (a.b <= 0) return set_error("batch size must be positive");
  if (a.d <= 0 || a.d > 256) return set_error("FlashAttention backward only supports head dimension at most 256");
  if (a.d % 8 != 0) return set_error("head_size should be a multiple of 8");
  if (a.d > 128 && !hd256_enabled()) return set_error("this build of fa_b200 supports head dimension at most 128");
  if (a.p_dropout < 0.f || a.p_dropout >= 1.f) return set_error("p_dropout must be in [0, 1)");
  if (a.p_dropout > 0.f && a.softcap > 0.f) return set_error("Softcapping does not support dropout for now");
  if (a.h <= 0 || a.h_k <= 0 || a.h % a.h_k != 0)
    return set_error("Number of heads in key/value must divide number of heads in query");
  if (varlen && (!a.cu_seqlens_q || !a.cu_seqlens_k)) return set_error("cu_seqlens_q and cu_seqlens_k are required");
  pl->kD = a.d <= 64 ? 64 : (a.d <= 128 ? 128 : 256);
  pl->sq_r = (a.seqlen_q + 127) / 128 * 128;
  int wl = a.window_size_left, wr = a.window_size_right;
  if (wl >= a.seqlen_k) wl = -1;
  if (wr >= a.seqlen_k) wr = -1;
  if (a.is_causal) wr = 0;
  pl->has_left = wl >= 0; pl->has_right = wr >= 0; pl->wl = wl; pl->wr = wr;
  // dense: (b, h, sq_r) rows; varlen: (h, R) rows with R = round4(total_q) + 132 * b + 128 (fa_bwd_sm100.cuh ws_rows):
  // grows with total_q, not b * max_seqlen_q (reference: total_q + 128 * b, flash_api.cpp:1142-1159)
  pl->ws_rows_head = varlen ? ((int64_t(a.total_q) + 3) / 4 * 4 + 132 * int64_t(a.b) + 128) : 0;
  const size_t stat = (varlen ? size_t(a.h) * size_t(pl->ws_rows_head) : size_t(a.b) * a.h * pl->sq_r) * sizeof(float);
  pl->off_lse = 0;
  pl->off_dps = align256(stat);
  pl->off_dq = pl->off_dps + align256(stat);
  pl->off_sem = pl->off_dq + align256(stat * pl->kD);   // deterministic mode: one counter per (b, h, query block),
  const size_t sem = size_t(a.b) * a.h * (pl->sq_r / 128) * sizeof(int);   // semaphores stay (b, h, query block)
  pl->off_sched = pl->off_sem + align256(sem * (pl->kD == 256 ? 4 : 2));   // work counter of the pair kernel's scheduler
  pl->total = pl->off_sched + 256;                                      // semaphores placed right behind dq_accum so one
                          // memset clears both (x2: the CTA-pair kernel orders each query half separately; x4: the
                          // head-dim 256 kernel has one chain per 64-query tile and output half)
  return 0;
}

template <typename T, int kD, bool kSoftcap, bool kDropout = false>
static int launch_bwd_main(const BwdKernelParams& kp, dim3 grid, cudaStream_t stream) {
  using Cfg = BwdCfg<kD>;
  static PerDeviceOnce configured;
  auto kern = fa_bwd_kernel<T, kD, kSoftcap, kDropout>;
  if (cudaError_t e = configured.run([&] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes); });
      e != cudaSuccess)
    return set_error(std::string("cudaFuncSetAttribute(bwd): ") + cudaGetErrorString(e));
  if (kp.sched_lpt) grid = dim3(grid.y, grid.z, grid.x);   // (h_k, b, key blocks): key-block index slowest
  kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(kp);
  g_launch_count++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(std::string("fa_bwd_kernel launch: ") + cudaGetErrorString(e));
  return 0;
}

// CTA-pair main loop (fa_bwd2_sm100.cuh): head dim 128 only
template <typename T, bool kSoftcap, bool kDropout>
static int launch_bwd_pair(const BwdKernelParams& kp, dim3 grid, cudaStream_t stream, int sm_margin) {
  static PerDeviceOnce configured;
  auto kern = fa_bwd2_kernel<T, kSoftcap, kDropout>;
  if (cudaError_t e = configured.run([&] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Bwd2Cfg::kSmemBytes); });
      e != cudaSuccess)
    return set_error(std::string("cudaFuncSetAttribute(bwd2): ") + cudaGetErrorString(e));
  // persistent: one CTA pair (cluster) per two SMs, fewer if there is less work; `sm_margin` SMs left free
  (void)grid;
  int sms = num_sms() - std::max(0, sm_margin);
  if (sms < 2) sms = 2;
  const int clusters = int(std::min<int64_t>(sms / 2, kp.total_items));
  grid = dim3(unsigned(2 * std::max(clusters, 1)), 1, 1);
  kern<<<grid, Bwd2Cfg::kThreads, Bwd2Cfg::kSmemBytes, stream>>>(kp);
  g_launch_count++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(std::string("fa_bwd2_kernel launch: ") + cudaGetErrorString(e));
  return 0;
}

// head dims 129..256 (fa_bwd256_sm100.cuh): two CTAs (output halves) per key block
template <typename T, bool kSoftcap, bool kDropout>
static int launch_bwd_hd256(const BwdKernelParams& kp, dim3 grid, cudaStream_t stream) {
  static PerDeviceOnce configured;
  auto kern = fa_bwd256_kernel<T, kSoftcap, kDropout>;
  if (cudaError_t e = configured.run([&] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Bwd256Cfg::kSmemBytes); });
      e != cudaSuccess)
    return set_error(std::string("cudaFuncSetAttribute(bwd256): ") + cudaGetErrorString(e));
  grid.x *= 2;
  kern<<<grid, Bwd256Cfg::kThreads, Bwd256Cfg::kSmemBytes, stream>>>(kp);
  g_launch_count++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(std::string("fa_bwd256_kernel launch: ") + cudaGetErrorString(e));
  return 0;
}

static bool use_pair_kernel(int kD) {
  // default on: measured 1125-1239 vs 1000-1084 TFLOP/s (bwd, b4 h16 s8192 d128); FA_B200_BWD_PAIR=0 selects fa_bwd_kernel
  static int mode = [] { const char* e = getenv("FA_B200_BWD_PAIR"); return e ? atoi(e) : 1; }();
  return kD == 128 && mode != 0;
}

template <typename T, int kD>
static int run_bwd_typed(const fa_b200_bwd_args& a, const BwdPlan& pl, bool varlen, cudaStream_t stream) {
  uint8_t* ws = reinterpret_cast<uint8_t*>(a.workspace);
  float* lse_log2 = reinterpret_cast<float*>(ws + pl.off_lse);
  float* dpsum = reinterpret_cast<float*>(ws + pl.off_dps);
  float* dq_accum = reinterpret_cast<float*>(ws + pl.off_dq);
  const int sq_r = pl.sq_r;
  // dq_accum is cleared by the preprocess kernel (same pass that reads O / dO); only the deterministic-mode
  // semaphores (and the 256-column accumulator of the head-dim 129..256 kernel) still need a memset
  cudaError_t e = cudaSuccess;
  float* clear_rows = dq_accum;
  if (kD == 256) {   // (the preprocess kernel clears 128 columns per row)
    clear_rows = nullptr;
    e = cudaMemsetAsync(dq_accum, 0, pl.total - pl.off_dq, stream);
  } else if (a.deterministic) {
    e = cudaMemsetAsync(ws + pl.off_sem, 0, pl.total - pl.off_sem, stream);
  }
  if (e != cudaSuccess) return set_error(std::string("cudaMemsetAsync(backward workspace): ") + cudaGetErrorString(e));
  {
    dim3 grid((sq_r + 31) / 32, a.h, a.b);
    fa_bwd_preprocess<T><<<grid, 256, 0, stream>>>(
        reinterpret_cast<const T*>(a.out), reinterpret_cast<const T*>(a.dout), a.softmax_lse, dpsum, lse_log2, a.softmax_d,
        a.o_batch_stride, a.o_row_stride, a.o_head_stride, a.do_batch_stride, a.do_row_stride, a.do_head_stride,
        a.cu_seqlens_q, a.b, a.h, a.seqlen_q, sq_r, a.d, a.total_q, clear_rows, kD, pl.ws_rows_head,
        reinterpret_cast<int*>(ws + pl.off_sched));
    g_launch_count++;
    if ((e = cudaGetLastError()) != cudaSuccess) return set_error(std::string("fa_bwd_preprocess launch: ") + cudaGetErrorString(e));
  }
  BwdKernelParams kp;
  memset(&kp, 0, sizeof kp);
  const CUtensorMapDataType dt = a.dtype == FA_B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  auto map16 = [&](CUtensorMap* tm, const void* base, int64_t bs, int64_t rs, int64_t hs, int rows_dense, int rows_varlen,
                   int heads) -> int {
    uint64_t rows = varlen ? rows_varlen : rows_dense, nb = varlen ? 1 : a.b;
    uint64_t dims[4] = {uint64_t(a.d), rows, uint64_t(heads), nb};
    uint64_t str[3] = {uint64_t(rs) * 2, uint64_t(hs) * 2, uint64_t(bs) * 2};
    if (nb == 1) str[2] = rows * str[0];
    uint32_t box[4] = {64, 128, 1, 1};
    return make_tmap(tm, dt, 4, base, dims, str, box);
  };
  if (int rc = map16(&kp.tm_q, a.q, a.q_batch_stride, a.q_row_stride, a.q_head_stride, a.seqlen_q, a.total_q, a.h)) return rc;
  if (int rc = map16(&kp.tm_do, a.dout, a.do_batch_stride, a.do_row_stride, a.do_head_stride, a.seqlen_q, a.total_q, a.h)) return rc;
  if (int rc = map16(&kp.tm_k, a.k, a.k_batch_stride, a.k_row_stride, a.k_head_stride, a.seqlen_k, a.total_k, a.h_k)) return rc;
  if (int rc = map16(&kp.tm_v, a.v, a.v_batch_stride, a.v_row_stride, a.v_head_stride, a.seqlen_k, a.total_k, a.h_k)) return rc;
  {
    constexpr uint64_t esz = 4;
    const uint64_t rows_head = varlen ? uint64_t(pl.ws_rows_head) : uint64_t(sq_r);
    uint64_t dims[4] = {uint64_t(kD), rows_head, uint64_t(a.h), uint64_t(varlen ? 1 : a.b)};
    uint64_t str[3] = {uint64_t(kD) * esz, rows_head * kD * esz, uint64_t(a.h) * rows_head * kD * esz};
    uint32_t box[4] = {uint32_t(128 / esz), 128, 1, 1};
    if (int rc = make_tmap(&kp.tm_dq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                           dq_accum, dims, str, box)) return rc;
  }
  const bool pair = use_pair_kernel(kD);
  if (pair || kD == 256) {
    auto map64 = [&](CUtensorMap* tm, const void* base, int64_t bs, int64_t rs, int64_t hs, int rows_dense, int rows_varlen,
                     int heads) -> int {
      uint64_t rows = varlen ? rows_varlen : rows_dense, nb = varlen ? 1 : a.b;
      uint64_t dims[4] = {uint64_t(a.d), rows, uint64_t(heads), nb};
      uint64_t str[3] = {uint64_t(rs) * 2, uint64_t(hs) * 2, uint64_t(bs) * 2};
      if (nb == 1) str[2] = rows * str[0];
      uint32_t box[4] = {64, 64, 1, 1};
      return make_tmap(tm, dt, 4, base, dims, str, box);
    };
    if (int rc = map16(&kp.tm_dk, a.dk, a.dk_batch_stride, a.dk_row_stride, a.dk_head_stride, a.seqlen_k, a.total_k, a.h_k)) return rc;
    if (int rc = map16(&kp.tm_dv, a.dv, a.dv_batch_stride, a.dv_row_stride, a.dv_head_stride, a.seqlen_k, a.total_k, a.h_k)) return rc;
    if (int rc = map64(&kp.tm_q64, a.q, a.q_batch_stride, a.q_row_stride, a.q_head_stride, a.seqlen_q, a.total_q, a.h)) return rc;
    if (int rc = map64(&kp.tm_do64, a.dout, a.do_batch_stride, a.do_row_stride, a.do_head_stride, a.seqlen_q, a.total_q, a.h)) return rc;
    const uint64_t rows_head = varlen ? uint64_t(pl.ws_rows_head) : uint64_t(sq_r);
    uint64_t dims[4] = {uint64_t(kD), rows_head, uint64_t(a.h), uint64_t(varlen ? 1 : a.b)};
    uint64_t str[3] = {uint64_t(kD) * 4, rows_head * kD * 4, uint64_t(a.h) * rows_head * kD * 4};
    uint32_t box[4] = {32, 64, 1, 1};
    if (int rc = make_tmap(&kp.tm_dq64, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dq_accum, dims, str, box)) return rc;
  }
  kp.lse_log2 = lse_log2;
  kp.dq_semaphore = a.deterministic ? reinterpret_cast<int*>(ws + pl.off_sem) : nullptr;
  kp.dpsum = dpsum;
  kp.dk = a.dk; kp.dv = a.dv;
  kp.dk_batch_stride = a.dk_batch_stride; kp.dk_row_stride = a.dk_row_stride; kp.dk_head_stride = a.dk_head_stride;
  kp.dv_batch_stride = a.dv_batch_stride; kp.dv_row_stride = a.dv_row_stride; kp.dv_head_stride = a.dv_head_stride;
  kp.cu_seqlens_q = a.cu_seqlens_q; kp.cu_seqlens_k = a.cu_seqlens_k;
  kp.alibi_slopes = a.alibi_slopes; kp.alibi_batch_stride = a.alibi_slopes_batch_stride;
  kp.b = a.b; kp.h = a.h; kp.h_k = a.h_k; kp.sq = a.seqlen_q; kp.sk = a.seqlen_k; kp.d = a.d; kp.g = a.h / a.h_k;
  kp.sq_r = sq_r;
  kp.ws_head_pitch = pl.ws_rows_head;
  kp.dq_accum = dq_accum;
  kp.has_left = pl.has_left; kp.has_right = pl.has_right; kp.wl = pl.wl; kp.wr = pl.wr;
  kp.scale = a.softmax_scale;
  if (a.softcap > 0.f) { kp.softcap_pre = a.softmax_scale / a.softcap; kp.scale_eff = a.softcap; }
  else { kp.softcap_pre = 0.f; kp.scale_eff = a.softmax_scale; }
  kp.scale_log2 = kp.scale_eff * float(M_LOG2E);
  dim3 grid((a.seqlen_k + 127) / 128, a.h_k, a.b);
  kp.sched_lpt = int64_t(a.h_k) * a.b <= 65535 ? sched_chunk(pl.has_left || pl.has_right) : 0;
  int rc = 0;
  bool launched = false;
  if (a.p_dropout > 0.f) {
    kp.drop = make_dropout_params(a.p_dropout, a.philox_seed, a.philox_offset);
    kp.rng_state = a.rng_state;
  }
  if constexpr (kD == 128) {
    if (pair) {
      // work items of the persistent pair kernel: (pair of key blocks, kv head, batch); groups of (kv head, batch)
      // combinations whose Q + dO + dq_accum stream (16 bytes per query x head-dim element and query head) shares L2
      kp.num_pairs = (a.seqlen_k + 255) / 256;
      const int64_t hz_total = int64_t(a.h_k) * a.b;
      const double q_bytes = double(a.seqlen_q) * a.d * (2 + 2 + 4) * (a.h / a.h_k);
      int64_t grp = int64_t(std::max(1.0, std::floor(48.0 * 1024 * 1024 / std::max(q_bytes, 1.0))));
      if (const char* e = getenv("FA_B200_SCHED_GROUP_BWD")) grp = atoi(e);
      kp.sched_group = int(std::max<int64_t>(1, std::min<int64_t>(grp, hz_total)));
      const int64_t total = int64_t(kp.num_pairs) * hz_total;
      if (total > INT32_MAX) return set_error("too many work items");
      kp.total_items = int(total);
      bool sched_dynamic = true;   // (items of the pair kernel are long: always worth it; FA_B200_SCHED_DYNAMIC=0: static deal)
      if (const char* e = getenv("FA_B200_SCHED_DYNAMIC")) sched_dynamic = atoi(e) != 0;
      {
        const int sms = std::max(2, num_sms() - std::max(0, a.sm_margin));
        if (sched_dynamic && total > sms / 2) kp.sched_counter = reinterpret_cast<int*>(ws + pl.off_sched);
      }
      rc = a.p_dropout > 0.f ? launch_bwd_pair<T, false, true>(kp, grid, stream, a.sm_margin)
         : a.softcap > 0.f ? launch_bwd_pair<T, true, false>(kp, grid, stream, a.sm_margin)
                           : launch_bwd_pair<T, false, false>(kp, grid, stream, a.sm_margin);
      launched = true;
    }
  }
  if constexpr (kD == 256) {
    kp.sched_lpt = 0;
    rc = a.p_dropout > 0.f ? launch_bwd_hd256<T, false, true>(kp, grid, stream)
       : a.softcap > 0.f ? launch_bwd_hd256<T, true, false>(kp, grid, stream)
                         : launch_bwd_hd256<T, false, false>(kp, grid, stream);
    launched = true;
  }
  if constexpr (kD != 256) {
    if (!launched) {
      if (a.p_dropout > 0.f) rc = launch_bwd_main<T, kD, false, true>(kp, grid, stream);
      else rc = a.softcap > 0.f ? launch_bwd_main<T, kD, true>(kp, grid, stream) : launch_bwd_main<T, kD, false>(kp, grid, stream);
    }
  }
  if (rc) return rc;
  {
    dim3 cgrid((a.seqlen_q + 7) / 8, a.h, a.b);
    fa_bwd_convert_dq<T, kD><<<cgrid, 256, 0, stream>>>(dq_accum, reinterpret_cast<T*>(a.dq), a.dq_batch_stride, a.dq_row_stride,
                                                      a.dq_head_stride, a.cu_seqlens_q, a.h, a.seqlen_q, sq_r, a.d,
                                                      a.softmax_scale, pl.ws_rows_head);
    g_launch_count++;
    if ((e = cudaGetLastError()) != cudaSuccess) return set_error(std::string("fa_bwd_convert_dq launch: ") + cudaGetErrorString(e));
  }
  return 0;
}

static int run_bwd(const fa_b200_bwd_args& a, bool varlen, cudaStream_t stream) {
  BwdPlan pl;
  if (int rc = plan_bwd(a, varlen, &pl)) return rc;
  if (!a.workspace || a.workspace_bytes < pl.total) return set_error("workspace too small for backward (query fa_b200_bwd_workspace_bytes)");
  if (reinterpret_cast<uintptr_t>(a.workspace) % 256 != 0) return set_error("backward workspace must be 256-byte aligned");
  if (a.seqlen_q == 0 || a.seqlen_k == 0) return 0;
  if (a.dtype == FA_B200_BF16)
    return pl.kD == 64    ? run_bwd_typed<__nv_bfloat16, 64>(a, pl, varlen, stream)
           : pl.kD == 128 ? run_bwd_typed<__nv_bfloat16, 128>(a, pl, varlen, stream)
                          : run_bwd_typed<__nv_bfloat16, 256>(a, pl, varlen, stream);
  return pl.kD == 64    ? run_bwd_typed<__half, 64>(a, pl, varlen, stream)
         : pl.kD == 128 ? run_bwd_typed<__half, 128>(a, pl, varlen, stream)
                        : run_bwd_typed<__half, 256>(a, pl, varlen, stream);
}

}  // namespace fa

using namespace fa;

extern "C" {
int fa_b200_mha_bwd(const fa_b200_bwd_args* args, fa_b200_stream_t stream) {
  if (!args) return set_error("null args");
  if (args->cu_seqlens_q || args->cu_seqlens_k) return set_error("fa_b200_mha_bwd takes dense tensors; use fa_b200_mha_varlen_bwd");
  return run_bwd(*args, false, static_cast<cudaStream_t>(stream));
}
int fa_b200_mha_varlen_bwd(const fa_b200_bwd_args* args, fa_b200_stream_t stream) {
  if (!args) return set_error("null args");
  return run_bwd(*args, true, static_cast<cudaStream_t>(stream));
}
#if FA_BWD2_TS
// timeline-debugging builds only (not declared in include/fa_b200.h): copies the clock64 stamps of the pair kernel
// per-CTA (smid, entry ns, exit ns) of the pair kernel launches since the last reset; returns the number of records
int fa_b200_debug_read_spans(unsigned long long* out, int max_records, int reset) {
  unsigned int n = 0;
  cudaMemcpyFromSymbol(&n, fa::g_bwd2_span_n, sizeof n);
  if (n > 8192u) n = 8192u;
  if (int(n) > max_records) n = unsigned(max_records);
  if (out) cudaMemcpyFromSymbol(out, fa::g_bwd2_span, size_t(n) * 3 * sizeof(unsigned long long));
  if (reset) { unsigned int z = 0; cudaMemcpyToSymbol(fa::g_bwd2_span_n, &z, sizeof z); }
  return int(n);
}
int fa_b200_debug_read_ts(unsigned long long* out, int n) {
  return cudaMemcpyFromSymbol(out, fa::g_bwd2_ts, size_t(n) * sizeof(unsigned long long)) == cudaSuccess ? 0 : 1;
}
#endif
size_t fa_b200_bwd_workspace_bytes(const fa_b200_bwd_args* args) {
  BwdPlan pl;
  if (!args || plan_bwd(*args, args->cu_seqlens_q != nullptr, &pl)) return 0;
  return pl.total;
}
}
