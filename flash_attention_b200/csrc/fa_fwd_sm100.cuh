// fa_fwd_sm100.cuh — forward attention mainloop for sm_100a (B200).
//
// PERSISTENT: the grid is one CTA per SM (two for the 1-tile configuration); every CTA loops over work items.  A work
// item is up to two 128-row query tiles ("tile 0/1", 256 packed rows) of one (batch, head-slot, split); its K/V blocks
// stream right-to-left (masked blocks first, like the reference: csrc/flash_attn/src/flash_fwd_kernel.h:138-140,274).
// Item order (fwd_decode_work): heavy (causal) m-blocks first inside groups of heads whose K/V fit L2 together (cf. the
// reference's LPT + L2 swizzle scheduler, flash_attn/cute/tile_scheduler.py:458-530).  With more items than CTAs they are
// drawn DYNAMICALLY from a global counter (struct FwdSched; cf. hopper/tile_scheduler.hpp DynamicPersistentTileScheduler);
// otherwise, or without scratch memory for the counter, dealt statically in snake order (round r: CTA c takes item
// r*G + c, or r*G + G-1-c on odd rounds).  Packed
// varlen batches are enumerated through a per-batch m-block prefix sum written by fa_fwd_varlen_prepare (cf.
// hopper/flash_prepare_scheduler.cu:43-215) instead of a max_seqlen grid with early exits.  Barrier phases, the K/V ring
// and TMEM stay live across items: the next item's Q / K loads and first S = QK^T run under the current item's epilogue.
//
//   warp 0-3  softmax warpgroup for tile 0 : tcgen05.ld S row -> mask -> online max/sum -> exp2 ->
//   warp 4-7  softmax warpgroup for tile 1   tcgen05.st P (aliases S in TMEM) ; lazy O rescale ; epilogue
//   warp 8    MMA issuer (one elected thread): S_i = Q_i K^T (SS), O_i += P_i V (TS, P read from TMEM)
//   warp 9    TMA producer: Q tiles once, K/V blocks through a 5-deep shared-memory ring
//
// TMEM (512 columns): S0 @0, S1 @128, O0 @256, O1 @384 (fp32); P_i overwrites S_i in place as packed
// 16-bit pairs.  S/P never touch HBM.  The two tiles ping-pong on the tensor pipe: while warpgroup i
// runs softmax on S_i the MMA thread executes the other tile's GEMMs.
//
// Semantics follow the reference kernels (file:line cited inline): bottom-right aligned causal /
// sliding-window masks (mask.h:172-173), ALiBi (mask.h:148-165, flash_fwd_kernel.h:294), tanh
// soft-cap (flash_api.cpp:130-134), LSE (softmax.h:169-186), GQA (flash_fwd_kernel.h:155), split-KV
// partials (flash_fwd_kernel.h:535-582,1026-1038), paged KV (flash_fwd_kernel.h:592-597).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "fa_dropout.cuh"
#include "fa_ptx.cuh"

#ifndef FA_FWD_REGS_SOFTMAX2
#define FA_FWD_REGS_SOFTMAX2 208   // setmaxnreg of the two softmax warpgroups (2-tile configuration); the rest gets 504 - 2x
#endif

namespace fa {

struct FwdKernelParams {
  CUtensorMap tm_q;  // 5-D: (d, gp, rows, h/gp, b)      box (64, gp, 128/gp, 1, 1)
  CUtensorMap tm_k;  // 4-D: (d, rows, h_k, b|pages|1)   box (64, 128, 1, 1)
  CUtensorMap tm_v;
  CUtensorMap tm_k256;  // wide-S kernel (fa_fwd_wide_sm100.cuh): the same tensors with a box of 256 rows
  CUtensorMap tm_v256;
  const void* q_raw;    // wide-S kernel: plain pointers + element strides of q / k / v (its exact per-row fallback reads them)
  const void* k_raw;
  const void* v_raw;
  int64_t q_bs, q_rs, q_hs, k_bs, k_rs, k_hs, v_bs, v_rs, v_hs;
  void* o;
  float* lse;
  float* o_accum;    // (splits, b, h, sq, kD) fp32 when num_splits > 1
  float* lse_accum;  // (splits, b, h, sq)
  int64_t o_batch_stride, o_row_stride, o_head_stride;
  const int* cu_seqlens_q;
  const int* cu_seqlens_k;
  const int* seqused_k;
  const int* seqused_q;        // valid queries per batch (rows beyond are neither computed nor written)
  const int* cache_batch_idx;  // kvcache: batch b reads cache row cache_batch_idx[b]
  const int* leftpad_k;        // kvcache: first valid key row of batch b
  const int* block_table;
  int64_t block_table_batch_stride;
  const float* alibi_slopes;
  int64_t alibi_batch_stride;
  const float* q_descale;
  const float* k_descale;
  const float* v_descale;
  int64_t q_descale_bs, k_descale_bs, v_descale_bs;
  float* debug_s;
  int b, h, h_k, sq, sk, d;
  int g;         // h / h_k
  int gp;        // query heads packed into the M dimension (1 or g)
  int total_q;   // varlen: rows of the packed q tensor (LSE row stride), else 0
  int page_size; // 0 = not paged
  int num_splits;
  int has_right, has_left;  // window limits active (causal => has_right, wr = 0)
  int wl, wr;
  int reverse_m;            // schedule heavy m-blocks first (causal)
  int sched_group;          // (head-slot x batch x split) combinations per scheduling group: their K/V share L2
  int num_m_blocks;         // dense: m-blocks per (batch, head slot); varlen: upper bound (max_seqlen_q)
  int total_items;          // dense: num_m_blocks * head slots * batch * splits; varlen: read from varlen_prefix[b]
  const int* varlen_prefix; // packed varlen: exclusive prefix sum of per-batch m-block counts, (b + 1) ints, else nullptr
  int* sched_counter;       // dynamic scheduling: zeroed before the launch; CTA c starts with item c and then takes item
                            // gridDim.x + atomicAdd(sched_counter, 1) until the list is exhausted.  nullptr: static rounds
  float scale;              // effective softmax scale (softcap if softcap > 0)
  float scale_log2;         // scale * log2(e)
  float softcap_pre;        // softmax_scale / softcap, 0 = off
  DropoutParams drop;       // kDropout kernels only
};

// kTiles = 2: the throughput configuration (256 query rows per CTA, 5-stage K/V ring, one CTA per SM).
// kTiles = 1: the latency configuration for short work items (decode, short / causal sequences): 128 rows per CTA,
//             2-stage ring, 256 TMEM columns, 256 threads -> TWO CTAs per SM, which overlap each other's softmax /
//             MMA phases and hide each other's prologue and epilogue.
template <int kD, int kBytes = 2, int kTiles = 2>
struct FwdCfg {
  static constexpr int kBlockM = 128;
  static constexpr int kBlockN = 128;
  static constexpr int kStages = kTiles == 2 ? 5 : 2;
  static constexpr int kThreads = 128 * (kTiles + 1);  // softmax warpgroup per tile + {MMA, TMA, tail cleaner, idle}
  static constexpr int kCtasPerSm = kTiles == 2 ? 1 : 2;
  static constexpr int kTmemCols = (kTiles * (128 + kD) <= 256) ? 256 : 512;  // S (128) + O (kD) per tile, power of two
  static constexpr int kWarpMma = 4 * kTiles, kWarpTma = 4 * kTiles + 1, kWarpClean = 4 * kTiles + 2;
  // setmaxnreg split: the CTA's register pool is kThreads * 168 (launch bounds 384 x 1 -> 168 regs/thread);
  // 2 softmax warpgroups + 1 MMA/TMA warpgroup must not ask for more than that or setmaxnreg.inc blocks forever.
  static constexpr int kRegsSoftmax = kTiles == 2 ? FA_FWD_REGS_SOFTMAX2 : 200;
  static constexpr int kRegsOther = kTiles == 2 ? (504 - 2 * FA_FWD_REGS_SOFTMAX2) : 56;
  static_assert(kTiles * kRegsSoftmax + kRegsOther <= (kTiles + 1) * (kTiles == 2 ? 168 : 128),
                "setmaxnreg split exceeds the CTA register pool");
  static constexpr int kPanels = kD * kBytes / 128;        // 128-byte swizzle panels per row (64 x 16-bit or 128 x 8-bit)
  static constexpr int kPanelElems = 128 / kBytes;
  static constexpr int kPanelBytes = 128 * 128;           // 128 rows x 128 B
  static constexpr int kTileBytes = kPanels * kPanelBytes;  // Q tile == K block == V block
  static constexpr int kSmemQ = kTiles * kTileBytes;
  static constexpr int kSmemKV = kStages * kTileBytes;
  static constexpr int kNumBars = 2 + 2 * kStages + 20;   // + q_empty[2], o_free[2], v_start, sched_full[2], sched_empty[2], o_half[2]
  static constexpr int kSmemBytes = kSmemQ + kSmemKV + 1024 /*barriers + tmem ptr*/ + 1024 /*align slack*/;
  static_assert(kSmemBytes <= 232448, "shared memory budget");  // hd256 (1 tile, 194 KB) runs one CTA per SM
};

// 16-bit element traits
template <typename T> struct Elt;
template <> struct Elt<__nv_bfloat16> {
  static constexpr uint32_t kFmt = 1;
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) { return pack_bf16(lo, hi); }
};
template <> struct Elt<__half> {
  static constexpr uint32_t kFmt = 0;
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) { return pack_f16(lo, hi); }
};
// FP8 e4m3 inputs (forward only): kind::f8f6f4 MMA, P quantised to e4m3 after a 2^8 up-scale, bf16 output
// (spec: hopper/flash_api.cpp mha_fwd descales; flash_attn/cute/flash_fwd_sm100.py:2003-2020,2179-2201)
struct fp8_e4m3 {};
template <> struct Elt<fp8_e4m3> {
  static constexpr uint32_t kFmt = 0;  // E4M3 in the f8f6f4 instruction descriptor
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) { return pack_bf16(lo, hi); }  // output = bf16
};
template <typename T> struct EltInfo { static constexpr int kBytes = 2; static constexpr bool kFp8 = false; using Out = T; };
template <> struct EltInfo<fp8_e4m3> { static constexpr int kBytes = 1; static constexpr bool kFp8 = true; using Out = __nv_bfloat16; };

__device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

#ifndef FA_FWD_TS
#define FA_FWD_TS 0   // 1: CTA (0,0,0) records clock64 stamps of its pipeline events (tools/probe_fwd_ts.py)
#endif
#if FA_FWD_TS
__device__ unsigned long long g_fwd_ts[16 * 16];   // [iteration - 8][event]
#define FTS(ev)                                                                                       \
  do {                                                                                                \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && it >= 8 && it < 24)                   \
      g_fwd_ts[(it - 8) * 16 + (ev)] = clock64();                                                      \
  } while (0)
// per item: globaltimer at the softmax warpgroup's entry into / exit from each work item, per CTA, + the SM id
__device__ unsigned long long g_fwd_item_ts[192 * 64 * 2];
__device__ unsigned int g_fwd_item_sm[192];
#define FITEM(slot_)                                                                                  \
  do {                                                                                                \
    if (tile == 0 && wq == 0 && lane == 0 && blockIdx.x < 192 && r < 64) {                             \
      unsigned long long t_;                                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                          \
      g_fwd_item_ts[(blockIdx.x * 64 + r) * 2 + (slot_)] = t_;                                         \
      if (r == 0) { unsigned int sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); g_fwd_item_sm[blockIdx.x] = sm_; } \
    }                                                                                                 \
  } while (0)
#else
#define FTS(ev) do {} while (0)
#define FITEM(slot_) do {} while (0)
#endif

#ifndef FA_FWD_POLY_PAIRS_FP8
#define FA_FWD_POLY_PAIRS_FP8 1
#endif
#ifndef FA_FWD_POLY_DEG_FP8
#define FA_FWD_POLY_DEG_FP8 3
#endif
#ifndef FA_FWD_ONEPASS
#define FA_FWD_ONEPASS 0    // 1: unmasked blocks take the exponentials against the STALE reference maximum in one pass (no row-max
                            // reduction in front of them); a row-sum guard catches the rows whose reference has to move.
                            // Measured (profiles/r02_j_fwd_latesum_abab.log, same lease): two-pass 1258, one-pass 1279, two-pass
                            // with the LATE row sum 1320 TFLOP/s at the C2 shape -- the late sum is the default, the one-pass
                            // variants (their guards put work back on the S -> P chain, the OR guard costs registers:
                            // profiles/r02_m_fwd_orguard_abab.log) stay as switches
#endif
#ifndef FA_FWD_SUM_LATE
#define FA_FWD_SUM_LATE 1   // 1 (two-pass blocks only): the row sum is taken AFTER the block's P has been handed to the MMA thread
                            // (the exponentials stay in the score registers), i.e. off the S -> P latency that paces the
                            // two-tile ping-pong (the reference does the same: flash_attn/cute/flash_fwd_sm100.py:2545-2546)
#endif
#ifndef FA_FWD_ORGUARD
#define FA_FWD_ORGUARD 0    // 1 (bf16 one-pass blocks): the guard of the stale-reference blocks is an OR over the packed 16-bit
                            // P words (P is written as 2^(x - 13): bit 14 of a bf16 exponent field <=> x >= 14) instead of the
                            // half-row sums, so the row sum can run late as well; fp16 kernels fall back to two-pass blocks
#endif
#ifndef FA_FWD_PSPLIT
#define FA_FWD_PSPLIT 64    // columns of P handed to the MMA thread with the first arrive (64 or 96; the rest with the second)
#endif
#ifndef FA_FWD_EXP_PIPE
#define FA_FWD_EXP_PIPE 1   // 1: software-pipelined exponential phase (see the softmax loop); 0: the round-1 loop
#endif
#ifndef FA_FWD_POLY8
#define FA_FWD_POLY8 (2 * FA_FWD_POLY_PAIRS)   // pipelined loop: polynomial pairs out of every 8 pairs (16 scores)
#endif
#ifndef FA_FWD_POLY8_FP8
#define FA_FWD_POLY8_FP8 (2 * FA_FWD_POLY_PAIRS_FP8)
#endif
#ifndef FA_FWD_POLY_PAIRS
#define FA_FWD_POLY_PAIRS 1  // measured on B200 (r01): 0 -> 1315, 1 -> 1339, 2 -> 1307, 3 -> 1147, 4 -> 1046 TFLOP/s fwd
#endif

template <int kV> struct HalfIdx { static constexpr int value = kV; };

// ---------------------------------------------------------------------------------------------- work items
struct FwdWork {
  int m_block, slot, bidb, split;
  int q_off, sq_b, k_off, sk_b, rows_total, row0, ntiles;
  int hk_idx, shift, nblk_total;
  int n_min, n_max, nblocks;   // the CTA's K/V block stream: j = n_max - 1 ... n_min  (iteration it <-> block n_max-1-it)
  int it_lo0, it_hi0, it_lo1, it_hi1;   // tile i takes part in stream iterations [it_lo_i, it_hi_i) (a causal / local
                               // mask can spare the lower-row tile's first block or the upper-row tile's last one)
  __device__ __forceinline__ int it_lo(int i) const { return i == 0 ? it_lo0 : it_lo1; }
  __device__ __forceinline__ int it_hi(int i) const { return i == 0 ? it_hi0 : it_hi1; }
  int v_tail;                  // the first streamed block is the ragged last block of the sequence
};

// Linear item index -> work item.  Returns false for an index that holds no work (past the end / empty m-block).
template <int kRowsPerCta, int kTiles>
__device__ __forceinline__ bool fwd_decode_work(const FwdKernelParams& p, int L, FwdWork& w) {
  const int nslots = p.h / p.gp;
  int m_idx, nm;
  if (p.varlen_prefix) {   // batch-major; inside a batch heavy m-blocks first, head slots fastest
    int lo = 0, hi = p.b;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (p.varlen_prefix[mid] * nslots <= L) lo = mid; else hi = mid;
    }
    w.bidb = lo;
    const int rem = L - p.varlen_prefix[lo] * nslots;
    nm = p.varlen_prefix[lo + 1] - p.varlen_prefix[lo];
    m_idx = rem / nslots;
    w.slot = rem - m_idx * nslots;
    w.split = 0;
  } else {
    const int hz_total = nslots * p.b * p.num_splits;
    const int per_group = p.num_m_blocks * p.sched_group;
    const int grp = L / per_group;
    const int within = L - grp * per_group;
    int gsz = hz_total - grp * p.sched_group;
    gsz = gsz < p.sched_group ? gsz : p.sched_group;
    m_idx = within / gsz;
    const int hz = grp * p.sched_group + (within - m_idx * gsz);
    nm = p.num_m_blocks;
    w.slot = hz % nslots;
    const int bz = hz / nslots;
    w.bidb = bz / p.num_splits;
    w.split = bz - w.bidb * p.num_splits;
  }
  w.m_block = p.reverse_m ? (nm - 1 - m_idx) : m_idx;
  const int bidb = w.bidb;
  w.q_off = 0; w.sq_b = p.sq; w.k_off = 0; w.sk_b = p.sk;
  if (p.cu_seqlens_q) { w.q_off = p.cu_seqlens_q[bidb]; w.sq_b = p.cu_seqlens_q[bidb + 1] - w.q_off; }
  if (p.cu_seqlens_k) { w.k_off = p.cu_seqlens_k[bidb]; w.sk_b = p.cu_seqlens_k[bidb + 1] - w.k_off; }
  if (p.seqused_k) w.sk_b = p.seqused_k[bidb];
  if (p.seqused_q) w.sq_b = p.seqused_q[bidb];
  if (p.leftpad_k) {  // keys start leftpad_k rows into the batch's rows; seqlen_k = seqused_k - leftpad_k (block_info.h:12-45)
    const int lp = p.leftpad_k[bidb];
    w.k_off += lp;
    w.sk_b -= lp;
  }
  w.rows_total = w.sq_b * p.gp;
  w.row0 = w.m_block * kRowsPerCta;
  if (w.row0 >= w.rows_total) return false;
  w.ntiles = (kTiles == 2 && w.row0 + 128 < w.rows_total) ? 2 : 1;
  w.hk_idx = (w.slot * p.gp) / p.g;
  w.shift = w.sk_b - w.sq_b;  // bottom-right alignment (mask.h:172-173)
  w.nblk_total = ceil_div(w.sk_b, 128);
  int s_lo = 0, s_hi = w.nblk_total;
  if (p.num_splits > 1) {  // equal block ranges per split (flash_fwd_kernel.h:535-543)
    const int per = ceil_div(w.nblk_total, p.num_splits);
    s_lo = w.split * per;
    s_hi = (w.split + 1) * per;
  }
  int t_min0 = 0, t_max0 = 0, t_min1 = 0, t_max1 = 0;
  w.n_min = 1 << 30; w.n_max = 0;
#pragma unroll
  for (int i = 0; i < kTiles; ++i) {
    if (i < w.ntiles) {
      const int r_lo = w.row0 + i * 128;
      const int r_hi = r_lo + 127 < w.rows_total - 1 ? r_lo + 127 : w.rows_total - 1;
      int nmax = w.nblk_total, nmin = 0;
      if (p.has_right) {
        const int lim = r_hi / p.gp + 1 + w.shift + p.wr;  // exclusive column limit of the tile's last row
        const int nb = lim <= 0 ? 0 : ceil_div(lim, 128);
        nmax = nb < nmax ? nb : nmax;
      }
      if (p.has_left) {
        const int lo = r_lo / p.gp + w.shift - p.wl;
        nmin = lo <= 0 ? 0 : lo / 128;
      }
      nmin = nmin > s_lo ? nmin : s_lo;
      nmax = nmax < s_hi ? nmax : s_hi;
      if (nmax > nmin) {
        if (i == 0) { t_min0 = nmin; t_max0 = nmax; } else { t_min1 = nmin; t_max1 = nmax; }
        w.n_min = nmin < w.n_min ? nmin : w.n_min;
        w.n_max = nmax > w.n_max ? nmax : w.n_max;
      }
    }
  }
  w.nblocks = w.n_max > w.n_min ? w.n_max - w.n_min : 0;
  w.it_lo0 = w.it_hi0 = w.it_lo1 = w.it_hi1 = 0;
  if (t_max0 > t_min0) { w.it_lo0 = w.n_max - t_max0; w.it_hi0 = w.n_max - t_min0; }
  if (kTiles == 2 && t_max1 > t_min1) { w.it_lo1 = w.n_max - t_max1; w.it_hi1 = w.n_max - t_min1; }
  w.v_tail = (w.nblocks > 0 && w.n_max == w.nblk_total && (w.sk_b & 127)) ? 1 : 0;
  return true;
}

// per-batch m-block counts of a packed varlen batch -> exclusive prefix sum (b + 1 ints): the persistent kernel
// enumerates exactly the m-blocks that exist (reference: hopper/flash_prepare_scheduler.cu:43-215)
static __global__ void __launch_bounds__(256) fa_fwd_varlen_prepare(const int* __restrict__ cu_seqlens_q, const int* __restrict__ seqused_q,
                                                            int b, int gp, int rows_per_cta, int* __restrict__ prefix,
                                                            int* __restrict__ sched_counter) {
  __shared__ int part[256];
  const int t = threadIdx.x;
  if (t == 0 && sched_counter) *sched_counter = 0;
  const int per = (b + 255) / 256;
  int sum = 0;
  for (int i = t * per; i < b && i < (t + 1) * per; ++i) {
    int sq = cu_seqlens_q[i + 1] - cu_seqlens_q[i];
    if (seqused_q) sq = seqused_q[i];
    sum += (sq * gp + rows_per_cta - 1) / rows_per_cta;
  }
  part[t] = sum;
  __syncthreads();
  if (t == 0) {
    int acc = 0;
    for (int i = 0; i < 256; ++i) { const int v = part[i]; part[i] = acc; acc += v; }
  }
  __syncthreads();
  int acc = part[t];
  for (int i = t * per; i < b && i < (t + 1) * per; ++i) {
    int sq = cu_seqlens_q[i + 1] - cu_seqlens_q[i];
    if (seqused_q) sq = seqused_q[i];
    prefix[i] = acc;
    acc += (sq * gp + rows_per_cta - 1) / rows_per_cta;
    if (i == b - 1) prefix[b] = acc;
  }
}


// Work-item sequence of one role of the persistent kernel.  Static mode (no counter): round r -> item r*G + c (snake
// order on odd rounds).  Dynamic mode: item c first, then whatever the CTA's TMA thread draws from the global counter;
// it publishes each draw through a two-slot shared-memory ring (sched_full / sched_empty mbarriers) that every other
// role reads in step.  Items are ordered heavy-first inside L2-sized head groups (fwd_decode_work), so a dynamic draw is
// longest-processing-time-first list scheduling: the static deal lost 8 % of a causal 8k forward to imbalance
// (profiles/r02_fwd_item_timeline.log).
struct FwdSched {
  int G, cta, total, rounds, k;
  uint32_t bar_full0, bar_empty0, item_addr;   // shared-memory addresses (barriers 8 bytes apart, items 4 bytes apart)
  bool dyn;
  __device__ __forceinline__ int static_item(int r) const { return r * G + ((r & 1) ? (G - 1 - cta) : cta); }
  __device__ __forceinline__ int first() {
    k = 0;
    return cta < total ? cta : -1;
  }
  // consumer roles; `lanes`: the whole warp calls (lane 0 signals after all lanes have read the slot)
  __device__ __forceinline__ int next(bool whole_warp, int lane) {
    ++k;
    if (!dyn) {
      if (k >= rounds) return -1;
      const int L = static_item(k);
      return L < total ? L : -1;
    }
    const int j = k - 1;
    mbar_wait(bar_full0 + 8u * (j & 1), (j >> 1) & 1, 600);
    int L;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(L) : "r"(item_addr + 4u * (j & 1)) : "memory");
    if (whole_warp) __syncwarp();
    if (!whole_warp || lane == 0) mbar_arrive(bar_empty0 + 8u * (j & 1));
    return L < total ? L : -1;
  }
  // TMA thread, dynamic mode: draw() may be called early (while the current item's last loads are being issued)
  __device__ __forceinline__ int draw(int* counter) const { return G + atomicAdd(counter, 1); }
  __device__ __forceinline__ int publish(int L) {   // returns the item to continue with (-1: none)
    const int j = k++;
    mbar_wait(bar_empty0 + 8u * (j & 1), ((j >> 1) & 1) ^ 1, 601);
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(item_addr + 4u * (j & 1)), "r"(L) : "memory");
    mbar_arrive(bar_full0 + 8u * (j & 1));
    return L < total ? L : -1;
  }
};

template <typename T, int kD, int kTiles, bool kDropout = false>
__global__ void __launch_bounds__((FwdCfg<kD, 2, kTiles>::kThreads), (FwdCfg<kD, 2, kTiles>::kCtasPerSm))
fa_fwd_kernel(const __grid_constant__ FwdKernelParams p) {
  constexpr int kBytes = EltInfo<T>::kBytes;
  constexpr bool kFp8 = EltInfo<T>::kFp8;
  using OutT = typename EltInfo<T>::Out;
  using Cfg = FwdCfg<kD, kBytes, kTiles>;
  constexpr int kRowsPerCta = 128 * kTiles;
  constexpr int S = Cfg::kStages;
  constexpr float kNegInf = -INFINITY;
  constexpr int kPolyDeg = kFp8 ? FA_FWD_POLY_DEG_FP8 : 3;
  constexpr int kPolyPairs = kFp8 ? FA_FWD_POLY_PAIRS_FP8 : FA_FWD_POLY_PAIRS;
  constexpr bool kBf16 = !kFp8 && Elt<T>::kFmt == 1;
  constexpr bool kOrGuard = FA_FWD_ORGUARD && FA_FWD_ONEPASS && FA_FWD_EXP_PIPE && kBf16 && !kDropout;
  constexpr bool kOnePassK = FA_FWD_ONEPASS && FA_FWD_EXP_PIPE && !kFp8 && !kDropout && (kOrGuard || !FA_FWD_ORGUARD);
  constexpr int kPSplit = kOnePassK ? 64 : FA_FWD_PSPLIT;   // (the one-pass guard works on halves)
  static_assert(kPSplit == 64 || kPSplit == 96, "FA_FWD_PSPLIT");
  constexpr bool kSumLate = FA_FWD_SUM_LATE && FA_FWD_EXP_PIPE && !kDropout && !kFp8 && (!kOnePassK || kOrGuard);

  // ------------------------------------------------------------------ item enumeration (identical in every role)
  const int G = int(gridDim.x), cta = int(blockIdx.x);
  const int total_items = p.varlen_prefix ? p.varlen_prefix[p.b] * (p.h / p.gp) : p.total_items;
  const int rounds = (total_items + G - 1) / G;

  // ------------------------------------------------------------------ shared memory carve-up
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sKV = sQ + Cfg::kSmemQ;
  const uint32_t sBar = sKV + Cfg::kSmemKV;
  auto bar_q_full = [&](int i) { return sBar + 8u * i; };
  auto bar_kv_full = [&](int s) { return sBar + 8u * (2 + s); };
  auto bar_kv_empty = [&](int s) { return sBar + 8u * (2 + S + s); };
  auto bar_s_full = [&](int i) { return sBar + 8u * (2 + 2 * S + i); };
  auto bar_p_full = [&](int i) { return sBar + 8u * (4 + 2 * S + i); };
  // bar_p_full(i): keys [0,64) of P_i are in TMEM; bar_p_full(i) + 16: keys [64,128)
  auto bar_o_full = [&](int i) { return sBar + 8u * (8 + 2 * S + i); };
  const uint32_t bar_v_clean = sBar + 8u * (10 + 2 * S);
  auto bar_q_empty = [&](int i) { return sBar + 8u * (11 + 2 * S + i); };   // every QK^T of the item has read Q_i
  auto bar_o_free = [&](int i) { return sBar + 8u * (13 + 2 * S + i); };    // the item's epilogue has read O_i out of TMEM
  // the MMA thread has reached an item whose first V block has a ragged tail: the cleaner warp may look at that item's
  // K/V slot.  (Without it the idle cleaner runs ahead of the pipeline by whole items, and a parity wait on a slot
  // barrier that is a full ring revolution behind passes at once: it then zeroed rows of a LIVE block of the item
  // still being computed -- a timing-dependent wrong result in split-KV decode with seqlen_k % 128 != 0.)
  const uint32_t bar_v_start = sBar + 8u * (15 + 2 * S);
  const uint32_t bar_sched_full0 = sBar + 8u * (16 + 2 * S), bar_sched_empty0 = sBar + 8u * (18 + 2 * S);
  auto bar_o_half = [&](int i) { return sBar + 8u * (20 + 2 * S + i); };   // first half of the block's P V has landed in O_i
  const uint32_t sTmemPtr = sBar + 8u * Cfg::kNumBars;
  FwdSched sched;
  sched.G = G; sched.cta = cta; sched.total = total_items; sched.rounds = rounds; sched.k = 0;
  sched.bar_full0 = bar_sched_full0; sched.bar_empty0 = bar_sched_empty0; sched.item_addr = sTmemPtr + 8u;
  sched.dyn = p.sched_counter != nullptr;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::kSmemQ + Cfg::kSmemKV + 8 * Cfg::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == Cfg::kWarpTma && lane == 0) {
    tma_prefetch_desc(&p.tm_q);
    tma_prefetch_desc(&p.tm_k);
    tma_prefetch_desc(&p.tm_v);
  }
  if (warp == Cfg::kWarpMma) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar_q_full(i), 1);
        mbar_init(bar_q_empty(i), 1);
        mbar_init(bar_s_full(i), 1);
        mbar_init(bar_p_full(i), 4);  // one arrive per softmax warp
        mbar_init(bar_p_full(i) + 16, 4);
        mbar_init(bar_o_full(i), 1);
        mbar_init(bar_o_free(i), 4);
        mbar_init(bar_o_half(i), 1);
      }
      for (int s = 0; s < S; ++s) {
        mbar_init(bar_kv_full(s), 1);
        mbar_init(bar_kv_empty(s), 1);
      }
      mbar_init(bar_v_clean, 1);
      mbar_init(bar_v_start, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar_sched_full0 + 8u * i, 1);
        mbar_init(bar_sched_empty0 + 8u * i, 2 + 4 * kTiles);   // MMA thread, cleaner warp, softmax warps
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(sTmemPtr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // register re-distribution (per warpgroup): softmax rows live in registers, the MMA/TMA warps need few
  if (warp >= Cfg::kWarpMma) {
  reg_dealloc<Cfg::kRegsOther>();
  // ================================================================== TMA producer
  if (warp == Cfg::kWarpTma) {
    if (elect_one()) {
      int ring = 0;              // running K/V slot counter across items: K_it -> ring, V_it -> ring + 1
      int q_items[2] = {0, 0};   // items so far that used Q tile i
      int nxt = -1;
      bool drawn = false;
      auto advance = [&]() -> int {
        if (!sched.dyn) return sched.next(false, 0);
        if (!drawn) nxt = sched.draw(p.sched_counter);
        drawn = false;
        return sched.publish(nxt);
      };
      for (int L = sched.first(); L >= 0; L = advance()) {
        FwdWork w;
        if (!fwd_decode_work<kRowsPerCta, kTiles>(p, L, w) || w.nblocks == 0) continue;
        const int bq = p.cu_seqlens_q ? 0 : w.bidb;
#pragma unroll
        for (int i = 0; i < kTiles; ++i) {
          if (w.it_hi(i) > w.it_lo(i)) {
            if (q_items[i] > 0) mbar_wait(bar_q_empty(i), (q_items[i] - 1) & 1, 110 + i);
            ++q_items[i];
            mbar_arrive_expect_tx(bar_q_full(i), Cfg::kTileBytes);
            const int qpos0 = w.q_off + (w.row0 + i * 128) / p.gp;
#pragma unroll
            for (int pn = 0; pn < Cfg::kPanels; ++pn)
              tma_load_5d(sQ + i * Cfg::kTileBytes + pn * Cfg::kPanelBytes, &p.tm_q, bar_q_full(i), pn * Cfg::kPanelElems, 0, qpos0,
                          w.slot, bq);
          }
        }
        for (int it = 0; it < w.nblocks; ++it) {
          if (sched.dyn && !drawn && it + 2 >= w.nblocks) {   // the next item's index is needed two blocks from now
            nxt = sched.draw(p.sched_counter);
            drawn = true;
          }
          const int j = w.n_max - 1 - it;
          int row = w.k_off + j * 128, c3 = p.cu_seqlens_k ? 0 : (p.cache_batch_idx ? p.cache_batch_idx[w.bidb] : w.bidb);
          if (p.page_size) {
            const int pos = j * 128;
            c3 = p.block_table[w.bidb * p.block_table_batch_stride + pos / p.page_size];
            row = pos % p.page_size;
          }
#pragma unroll
          for (int kv = 0; kv < 2; ++kv, ++ring) {
            const int s = ring % S;
            const uint32_t ph = (ring / S) & 1;
            mbar_wait(bar_kv_empty(s), ph ^ 1, 100 + s);
            mbar_arrive_expect_tx(bar_kv_full(s), Cfg::kTileBytes);
            const CUtensorMap* tm = kv == 0 ? &p.tm_k : &p.tm_v;
#pragma unroll
            for (int pn = 0; pn < Cfg::kPanels; ++pn)
              tma_load_4d(sKV + s * Cfg::kTileBytes + pn * Cfg::kPanelBytes, tm, bar_kv_full(s), pn * Cfg::kPanelElems, row, w.hk_idx,
                          c3);
          }
        }
      }
    }
    __syncwarp();
  }
  // ================================================================== MMA issuer (one elected thread)
  else if (warp == Cfg::kWarpMma) {
    if (elect_one()) {
      constexpr uint32_t fmt = Elt<T>::kFmt;
      constexpr uint32_t idesc_qk = make_idesc(fmt, fmt, 0, 0, 128, 128);
      constexpr uint32_t idesc_pv = make_idesc(fmt, fmt, 0, 1, 128, kD);
      constexpr uint64_t desc_k_major = smem_desc_sw128(16, 1024);
      constexpr uint64_t desc_mn_major = smem_desc_sw128(Cfg::kPanelBytes, 1024);
      constexpr uint32_t hi_k = uint32_t(desc_k_major >> 32), hi_mn = uint32_t(desc_mn_major >> 32);
      constexpr uint32_t lo_mn = uint32_t(desc_mn_major);  // carries the LBO field
      auto lo_of = [](uint32_t addr) { return (addr & 0x3FFFFu) >> 4; };
      auto mk = [](uint32_t hi, uint32_t lo) { return (uint64_t(hi) << 32) | lo; };
      const uint32_t tS[2] = {tmem_base + 0, tmem_base + 128};
      const uint32_t tO[2] = {tmem_base + kTiles * 128, tmem_base + kTiles * 128 + kD};

      auto issue_qk = [&](int i, int slot_k) {
        const uint32_t la = lo_of(sQ + i * Cfg::kTileBytes) | uint32_t(desc_k_major);
        const uint32_t lb = lo_of(sKV + slot_k * Cfg::kTileBytes) | uint32_t(desc_k_major);
#pragma unroll
        for (int kk = 0; kk < kD * kBytes / 32; ++kk) {  // one MMA per 32 bytes of the head dimension
          const uint32_t off = ((kk / 4) * Cfg::kPanelBytes + (kk % 4) * 32) >> 4;
          mma_ss<kFp8>(tS[i], mk(hi_k, la + off), mk(hi_k, lb + off), idesc_qk, kk > 0);
        }
      };
      auto issue_pv = [&](int i, int slot_v, bool acc, int half) {
        const uint32_t lb = lo_of(sKV + slot_v * Cfg::kTileBytes) | lo_mn;
        constexpr int kSteps = 4 * kBytes;              // MMAs per 128 keys (K = 16 keys for 16-bit, 32 for fp8)
        constexpr uint32_t kStepBytes = 4096 / kBytes;  // that many key rows of 128 B
        constexpr int kFirst = kSteps * kPSplit / 128;  // MMAs covered by the first P arrive
#pragma unroll
        for (int kk = 0; kk < kSteps; ++kk) {
          if ((kk < kFirst) != (half == 0)) continue;
          mma_ts<kFp8>(tO[i], tS[i] + kk * 8, mk(hi_mn, lb + kk * (kStepBytes >> 4)), idesc_pv, (acc || kk > 0) ? 1u : 0u);
        }
      };

      int ring = 0, vclean_cnt = 0;
      int q_items[2] = {0, 0};   // items so far in which tile i had work (Q_i loaded, O_i produced)
      int blk[2] = {0, 0};       // blocks so far processed by tile i: phase of s_full / p_full / o_full
      for (int L = sched.first(); L >= 0; L = sched.next(false, 0)) {
        FwdWork w;
        if (!fwd_decode_work<kRowsPerCta, kTiles>(p, L, w) || w.nblocks == 0) continue;
        const bool ts_item = q_items[0] + q_items[1] == 0;   // (the FTS stamps refer to the CTA's first item only)
        (void)ts_item;
        bool act[2], first_pv[2];
#pragma unroll
        for (int i = 0; i < kTiles; ++i) {
          act[i] = w.it_hi(i) > w.it_lo(i);
          first_pv[i] = true;
          if (act[i]) mbar_wait(bar_q_full(i), q_items[i] & 1, 200 + i);
        }
        if (w.v_tail) mbar_arrive(bar_v_start);   // every MMA of the earlier items has been issued behind its kv_full waits
        // prologue: S_i = Q_i K_0^T for the tiles that start with the first streamed block
        mbar_wait(bar_kv_full(ring % S), (ring / S) & 1, 210);
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < kTiles; ++i) {
          if (act[i] && w.it_lo(i) == 0) {
            issue_qk(i, ring % S);
            mma_commit(bar_s_full(i));
            if (w.it_hi(i) == 1) mma_commit(bar_q_empty(i));
          }
        }
        mma_commit(bar_kv_empty(ring % S));
        for (int it2 = 0; it2 < w.nblocks; ++it2) {
          const int ring_v = ring + 2 * it2 + 1, ring_k = ring + 2 * it2 + 2;
          const int sv = ring_v % S, sk_ = ring_k % S;
          const uint32_t phv = (ring_v / S) & 1, phk = (ring_k / S) & 1;
          const bool more = it2 + 1 < w.nblocks;
          const int it = ts_item ? it2 : -1;
          (void)it;
          if (it2 == 0 && w.v_tail) { mbar_wait(bar_v_clean, vclean_cnt & 1, 221); ++vclean_cnt; }  // rows beyond seqlen_k zeroed
          else mbar_wait(bar_kv_full(sv), phv, 220);
          bool k_waited = false;
          int last_pv = -1, last_qk = -1;
#pragma unroll
          for (int i = 0; i < kTiles; ++i) {
            if (act[i] && it2 >= w.it_lo(i) && it2 < w.it_hi(i)) last_pv = i;
            if (act[i] && it2 + 1 >= w.it_lo(i) && it2 + 1 < w.it_hi(i)) last_qk = i;
          }
#pragma unroll
          for (int i = 0; i < kTiles; ++i) {
            if (!act[i]) continue;
            if (it2 >= w.it_lo(i) && it2 < w.it_hi(i)) {
              if (first_pv[i]) {   // O_i of the previous item must have been read out by its epilogue
                if (q_items[i] > 0) mbar_wait(bar_o_free(i), (q_items[i] - 1) & 1, 225 + i);
              }
              mbar_wait(bar_p_full(i), blk[i] & 1, 230 + i);
              FTS(0 + 4 * i);
              tc_fence_after();
              issue_pv(i, sv, !first_pv[i], 0);
              mma_commit(bar_o_half(i));   // (only waited for when the softmax has to move its reference mid-block)
              mbar_wait(bar_p_full(i) + 16, blk[i] & 1, 232 + i);
              FTS(1 + 4 * i);
              tc_fence_after();
              issue_pv(i, sv, !first_pv[i], 1);
              mma_commit(bar_o_full(i));
              FTS(2 + 4 * i);
              first_pv[i] = false;
              ++blk[i];
              if (i == last_pv) mma_commit(bar_kv_empty(sv));
            }
            if (more && it2 + 1 >= w.it_lo(i) && it2 + 1 < w.it_hi(i)) {
              if (!k_waited) {
                mbar_wait(bar_kv_full(sk_), phk, 240);
                tc_fence_after();
                k_waited = true;
              }
              issue_qk(i, sk_);
              mma_commit(bar_s_full(i));
              FTS(3 + 4 * i);
              if (it2 + 2 == w.it_hi(i)) mma_commit(bar_q_empty(i));   // that was the item's last read of Q_i
              if (i == last_qk) mma_commit(bar_kv_empty(sk_));
            }
          }
        }
#pragma unroll
        for (int i = 0; i < kTiles; ++i)
          if (act[i]) ++q_items[i];
        ring += 2 * w.nblocks;
      }
    }
    __syncwarp();
  }
  // ================================================================== V-tail cleaner (otherwise idle warp)
  else if (warp == Cfg::kWarpClean) {
    int ring = 0, vstart_cnt = 0;
    for (int L = sched.first(); L >= 0; L = sched.next(true, lane)) {
      FwdWork w;
      if (!fwd_decode_work<kRowsPerCta, kTiles>(p, L, w) || w.nblocks == 0) continue;
      if (w.v_tail) {
        // Rows of the last V block beyond seqlen_k may hold garbage (kv-cache tail, next sequence): P is
        // exactly 0 there but 0 * NaN would poison O, so clear them before the first PV
        // (reference: Clear_OOB_MN on the V load, flash_fwd_kernel.h:254-262).
        const int rv = ring + 1;
        mbar_wait(bar_v_start, vstart_cnt & 1, 251);
        ++vstart_cnt;
        mbar_wait(bar_kv_full(rv % S), (rv / S) & 1, 250);
        const int valid = w.sk_b - (w.nblk_total - 1) * 128;
        const int nvec = (128 - valid) * 8;
        for (int pn = 0; pn < Cfg::kPanels; ++pn) {
          uint4* base = reinterpret_cast<uint4*>(smem_gen + Cfg::kSmemQ + (rv % S) * Cfg::kTileBytes + pn * Cfg::kPanelBytes +
                                                 valid * 128);
          for (int x = lane; x < nvec; x += 32) base[x] = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_v_clean);
      }
      ring += 2 * w.nblocks;
    }
  }
  } else {
  // ================================================================== softmax warpgroups
    reg_alloc<Cfg::kRegsSoftmax>();
    const int tile = warp >> 2;
    const int wq = warp & 3;
    const uint32_t lane_sel = uint32_t(wq * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel + tile * 128;
    const uint32_t tO = tmem_base + lane_sel + kTiles * 128 + tile * kD;
    int blk = 0;   // blocks processed so far by this tile (phase of s_full / o_full)
    int r = -1;   // items seen by this CTA (debug / timeline hooks refer to the first one)
    for (int L = sched.first(); L >= 0; L = sched.next(true, lane)) {
      ++r;
      FwdWork w;
      if (!fwd_decode_work<kRowsPerCta, kTiles>(p, L, w)) continue;
      if (tile >= w.ntiles) continue;
      FITEM(0);
      const int bidb = w.bidb, split = w.split, slot = w.slot, hk_idx = w.hk_idx, shift = w.shift;
      const int row0 = w.row0, rows_total = w.rows_total, sk_b = w.sk_b, q_off = w.q_off;
      const int it_lo = w.it_lo(tile), it_hi = w.it_hi(tile);
      const int nblocks = it_hi - it_lo;   // blocks of THIS tile
      const int rr = row0 + tile * 128 + wq * 32 + lane;  // packed row of this thread
      const int qpos = rr / p.gp;
      const int head = slot * p.gp + (rr - qpos * p.gp);
      const bool row_valid = rr < rows_total;
      // fp8: S = (q*qd).(k*kd) -> fold the de-scales into the softmax scale; O picks up vd in the epilogue
      float descale_qk = 1.f, descale_v = 1.f;
      if constexpr (kFp8) {
        if (p.q_descale) descale_qk *= p.q_descale[bidb * p.q_descale_bs + hk_idx];
        if (p.k_descale) descale_qk *= p.k_descale[bidb * p.k_descale_bs + hk_idx];
        if (p.v_descale) descale_v = p.v_descale[bidb * p.v_descale_bs + hk_idx];
      }
      const bool has_softcap = p.softcap_pre != 0.f;
      const float softcap_pre = p.softcap_pre * descale_qk;
      const float scale_eff = has_softcap ? p.scale : p.scale * descale_qk;       // natural-log scale of the raw scores
      const float scale_log2 = has_softcap ? p.scale_log2 : p.scale_log2 * descale_qk;
      // P is written as exp2(x - m_ref + kPOffset) and the reference max m_ref may lag the true row max by up to
      // kLagLog2 (exact after normalisation; avoids an O read-modify-write in TMEM on most blocks).  e4m3 keeps
      // 2^-9 .. 448, so fp8 shifts P up by 2^5 and allows a lag of 2^3: P <= 2^8, and probabilities down to
      // 2^-14 of the row max survive the cast (the reference scales by 2^8 with no lag, flash_fwd_sm100.py:2190;
      // with a zero lag almost every block rescales: 32 rows per warp, each a new max with probability ~1/j).
      // (OR-guard kernels: P = 2^(x - 13), so that "x >= 14" is the top exponent bit of the bf16 word; the power of two
      // cancels exactly in O / row_sum and bf16 has the exponent range to spare)
      constexpr float kPOffset = kFp8 ? 5.0f : (kOrGuard ? -13.0f : 0.0f);
      constexpr float kLagLog2 = kFp8 ? 3.0f : 8.0f;
      // tile-level limits for the "does this block need masking" test
      const int t_row_lo = row0 + tile * 128;
      int t_row_hi = t_row_lo + 127 < rows_total - 1 ? t_row_lo + 127 : rows_total - 1;
      const int t_qlo = t_row_lo / p.gp, t_qhi = t_row_hi / p.gp;
      // this row's visible column range [col_lo, col_hi)
      int col_hi = sk_b, col_lo = 0;
      if (p.has_right) { int l = qpos + 1 + shift + p.wr; col_hi = l < col_hi ? l : col_hi; }
      if (p.has_left) { int l = qpos + shift - p.wl; col_lo = l > 0 ? l : 0; }
      float slope = 0.f;
      if (p.alibi_slopes) slope = p.alibi_slopes[bidb * p.alibi_batch_stride + head] / scale_eff;

      constexpr bool kOnePass = kOnePassK;
      constexpr float kGuard = 16384.f;   // one-pass blocks: largest half-row sum of exp2(x) taken without moving the reference
                                          // (every term is then <= 2^14: finite in fp16 P, and far inside fp32 for the sums)
      const float inv_scale_log2 = 1.0f / scale_log2;
      float m_ref = kNegInf, row_sum = 0.f;
      for (int it = it_lo; it < it_hi; ++it, ++blk) {
        const int j = w.n_max - 1 - it;
        const int col0 = j * 128;
        mbar_wait(bar_s_full(tile), blk & 1, 300 + tile);
        if (wq == 0 && lane == 0 && r == 0) FTS(8 + 4 * tile);
        tc_fence_after();
        uint32_t su[128];
        float* s = reinterpret_cast<float*>(su);
        // which blocks need the slow path (score mods / masks)?  warp-uniform
        bool need_mask = col0 + 128 > sk_b;
        if (p.has_right) need_mask |= col0 + 128 > t_qlo + 1 + shift + p.wr;
        if (p.has_left) need_mask |= col0 < t_qhi + shift - p.wl;
        const bool slow = need_mask || has_softcap || p.alibi_slopes != nullptr ||
                          (p.debug_s != nullptr && it == 0 && r == 0);
        // One-pass blocks (unmasked, not the tile's first): no row maximum in front of the exponentials.  Softmax is
        // invariant to the reference as long as nothing overflows, so P = exp2((s - m_ref) * scale) is formed against the
        // reference left by the earlier blocks and a guard on the half-row sums (plain adds the loop has anyway) finds the
        // rows whose scores outgrew it; only those warps stop, take the maximum and redo the half (fix_half below).  The
        // row-max reduction -- 64 FMNMX3 and ~0.5k clk of the S -> P latency that paces the two-tile ping-pong -- is off the
        // common path (profiles/r02_fwd_timeline_r1_vs_pipelined.log: "S ok -> max done").
        const bool opt = kOnePass && !slow && it != it_lo;
        float mx = 0.f;
        if (opt) {
          tmem_ld_x32(tS + 0, su);
          tmem_ld_x32(tS + 32, su + 32);
          tmem_wait_ld();
          tmem_ld_x32(tS + 64, su + 64);   // waited for in front of the fifth chunk
          tmem_ld_x32(tS + 96, su + 96);
        } else if (!slow) {
          // fast path: second half of the row streams out of TMEM while the first half is reduced
          tmem_ld_x32(tS + 0, su);
          tmem_ld_x32(tS + 32, su + 32);
          tmem_wait_ld();
          tmem_ld_x32(tS + 64, su + 64);
          tmem_ld_x32(tS + 96, su + 96);
          float m0 = fmax3(s[0], s[1], s[2]), m1 = fmax3(s[3], s[4], s[5]);
          float m2 = fmax3(s[6], s[7], s[8]), m3 = fmax3(s[9], s[10], s[11]);
#pragma unroll
          for (int c = 12; c + 7 < 64; c += 8) {
            m0 = fmax3(m0, s[c], s[c + 1]);
            m1 = fmax3(m1, s[c + 2], s[c + 3]);
            m2 = fmax3(m2, s[c + 4], s[c + 5]);
            m3 = fmax3(m3, s[c + 6], s[c + 7]);
          }
          m0 = fmax3(m0, s[60], s[61]);
          m1 = fmax3(m1, s[62], s[63]);
          tmem_wait_ld();
#pragma unroll
          for (int c = 64; c + 7 < 128; c += 8) {
            m0 = fmax3(m0, s[c], s[c + 1]);
            m1 = fmax3(m1, s[c + 2], s[c + 3]);
            m2 = fmax3(m2, s[c + 4], s[c + 5]);
            m3 = fmax3(m3, s[c + 6], s[c + 7]);
          }
          mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        } else {
          tmem_ld_x32(tS + 0, su);
          tmem_ld_x32(tS + 32, su + 32);
          tmem_ld_x32(tS + 64, su + 64);
          tmem_ld_x32(tS + 96, su + 96);
          tmem_wait_ld();
          if (p.debug_s && blockIdx.x == 0 && r == 0 && tile == 0 && it == 0) {
#pragma unroll
            for (int c = 0; c < 128; ++c) p.debug_s[(wq * 32 + lane) * 128 + c] = s[c];
          }
          if (has_softcap) {
#pragma unroll
            for (int c = 0; c < 128; ++c) s[c] = tanh_approx(s[c] * softcap_pre);
          }
          if (p.alibi_slopes) {  // bias = -slope * |i + sk - sq - j|  (mask.h:148-165)
            const float rel0 = float(qpos + shift - col0);
#pragma unroll
            for (int c = 0; c < 128; ++c) s[c] = fmaf(-slope, fabsf(rel0 - float(c)), s[c]);
          }
          if (need_mask) {
            const int hi = col_hi - col0, lo = col_lo - col0;
#pragma unroll
            for (int c = 0; c < 128; ++c) s[c] = (c >= hi || c < lo) ? kNegInf : s[c];
          }
          float m0 = fmax3(s[0], s[1], s[2]), m1 = fmax3(s[3], s[4], s[5]);
#pragma unroll
          for (int c = 6; c + 3 < 128; c += 4) {
            m0 = fmax3(m0, s[c], s[c + 1]);
            m1 = fmax3(m1, s[c + 2], s[c + 3]);
          }
          mx = fmax3(m0, m1, fmaxf(s[126], s[127]));
        }
        // ---- online softmax: row max with a lazily updated reference (exact after normalisation)
        if (wq == 0 && lane == 0 && r == 0) FTS(9 + 4 * tile);
        const float m_new = fmaxf(m_ref, mx);
        if (opt) {
          // reference untouched
        } else if (it == it_lo) {
          m_ref = m_new;
        } else {
          bool need = (m_new - m_ref) * scale_log2 > kLagLog2;  // NaN (-inf - -inf) compares false
          if (m_ref == kNegInf) {  // nothing accumulated yet for this row: O row is 0, sum is 0
            m_ref = m_new;
            need = false;
          }
          if (__any_sync(0xffffffffu, need)) {
            mbar_wait(bar_o_full(tile), (blk - 1) & 1, 310 + tile);  // PV of the previous block has landed
            tc_fence_after();
            const float f = need ? ex2((m_ref - m_new) * scale_log2) : 1.0f;
#pragma unroll
            for (int cc = 0; cc < kD; cc += 32) {
              uint32_t o[32];
              tmem_ld_x32(tO + cc, o);
              tmem_wait_ld();
#pragma unroll
              for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * f);
              tmem_st_x32(tO + cc, o);
            }
            row_sum *= f;
            if (need) m_ref = m_new;
          }
        }
        const float neg_ms = (m_ref == kNegInf) ? kPOffset : kPOffset - m_ref * scale_log2;
        // ---- P = exp2(S*scale_log2 - m*scale_log2), written back over S as packed 16-bit.
        // Packed FFMA2 for the scaling; a share of the pairs takes the FMA-pipe polynomial instead
        // of MUFU.EX2 (the MUFU unit, 16 op/clk/SM, is otherwise as busy as the tensor pipe).
        const uint64_t sc2 = pack2(scale_log2, scale_log2);
        uint64_t ng2 = pack2(neg_ms, neg_ms);
        uint64_t acc_a = pack2(0.f, 0.f), acc_b = pack2(0.f, 0.f);
#if FA_FWD_EXP_PIPE
        if constexpr (!kDropout) {
        // Software-pipelined in chunks of 16 scores: stage A(k) scales the chunk (FMA pipe), stage B(k) takes the
        // exponentials (MUFU, and the polynomial share on the FMA pipe), adds the row sum and packs.  ptxas schedules the
        // fully unrolled block greedily and used to issue every scaling FFMA2 first (they are all ready when the row has
        // been loaded), which left a MUFU-bound tail of ~75 back-to-back MUFU.EX2 with an idle FMA pipe: the phases ran in
        // sequence, not in parallel (profiles/r01_sass_*; 1.66k clk for 128 exponentials vs ~0.8k of either pipe).  The
        // scale factor of chunk k+1 is therefore made to DEPEND on an exponential of chunk k-1 (one extra FFMA2 with a zero
        // multiplicand, which the assembler cannot fold): A(k+1) can only issue while B(k) is in flight, so both pipes stay busy.
        const uint64_t zero2 = pack2(0.f, 0.f);
        uint32_t pk[16];
        uint32_t g_or = 0, g_or2 = 0;   // OR of the half's packed P words (two chains)
        uint64_t tok_prev = zero2, tok_cur = zero2;   // an exponential pair of chunk k-1 / k
        auto stage_a = [&](int k, uint64_t sc) {
#pragma unroll
          for (int c = 0; c < 16; c += 2)
            unpack2(fma2(pack2(s[16 * k + c], s[16 * k + c + 1]), sc, ng2), s[16 * k + c], s[16 * k + c + 1]);
        };
        // one-pass blocks: move the reference of the rows that need it to the maximum of half h (columns 64h .. 64h+63) and
        // redo that half.  s[] holds x = (s - m_ref) * scale_log2 + kPOffset for columns [0, hi) (m_ref = -inf: x = s *
        // scale_log2 + kPOffset), so the raw scores are never needed again.  Cold path: plain loops, MUFU only.
        float hs0 = 0.f, xm_poly = -INFINITY;
        auto fix_half = [&](const int h, const float hs) {
          float mxx = s[64 * h];
#pragma unroll
          for (int c = 1; c < 64; ++c) mxx = fmaxf(mxx, s[64 * h + c]);
          const bool had_ref = m_ref != kNegInf;
          const bool need = !had_ref || !(hs <= kGuard) || xm_poly > 100.f;
          // (with a reference in place it only ever moves up: the half's own maximum may lie below it)
          const float delta = need ? (had_ref ? fmaxf(mxx - kPOffset, 0.f) : mxx - kPOffset) : 0.f;
          const float f = (need && had_ref) ? ex2(-delta) : 1.0f;
          // everything accumulated so far -- O (earlier blocks, and the first half of this one if h == 1) and the row sum --
          // moves to the new reference
          if (h == 0) mbar_wait(bar_o_full(tile), (blk - 1) & 1, 311 + tile);
          else mbar_wait(bar_o_half(tile), blk & 1, 313 + tile);
          tc_fence_after();
          tmem_wait_st();
#pragma unroll
          for (int cc = 0; cc < kD; cc += 32) {
            uint32_t o[32];
            tmem_ld_x32(tO + cc, o);
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * f);
            tmem_st_x32(tO + cc, o);
          }
          row_sum *= f;
          if (need) m_ref = (had_ref ? m_ref : 0.f) + delta * inv_scale_log2;
          const int hi = h == 0 ? 80 : 128;   // chunks already scaled (the software pipeline runs one chunk ahead)
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (c >= 64 * h && c < hi) s[c] -= delta;
          ng2 = add2(ng2, pack2(-delta, -delta));
          float sum = 0.f;
#pragma unroll
          for (int cc = 0; cc < 64; cc += 32) {
            uint32_t pk2[16];
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              const float e0 = ex2(s[64 * h + cc + c]), e1 = ex2(s[64 * h + cc + c + 1]);
              sum += e0 + e1;
              pk2[c / 2] = Elt<T>::pack(e0, e1);
            }
            tmem_st_x16(tS + (64 * h + cc) / 2, pk2);
          }
          xm_poly = -INFINITY;
          acc_a = pack2((h == 0 ? 0.f : hs0 * f) + sum, 0.f);
          acc_b = pack2(0.f, 0.f);
          return (h == 0 ? 0.f : hs0 * f) + sum;
        };
        // OR-guard variant of the same cold path.  The exponentials have replaced the scaled scores in s[] (late row sum), so
        // the half's scores come out of TMEM again: P occupies TMEM columns 0..63, so columns 64..127 of S are never
        // overwritten, and at the guard of the first half only the P words of columns 0..31 have been stored (TMEM columns
        // 0..15) -- the scaled scores of columns 0..15 are still in s[] (chunk 0 is summed early for that reason).
        // Fills lo[] / hi[] with the redone P words of columns 64h..64h+31 / 64h+32..64h+63.
        auto fix_half_or = [&](auto hc, uint32_t* lo, uint32_t* hi) {
          constexpr int h = decltype(hc)::value;   // (a run-time half index would turn s[] into a local-memory array)
          float nm0, nm1;
          unpack2(ng2, nm0, nm1);
          if (h == 0) {
            tmem_ld_x16(tS + 16, su + 16);
            tmem_ld_x32(tS + 32, su + 32);
          } else {
            tmem_ld_x32(tS + 64, su + 64);
            tmem_ld_x32(tS + 96, su + 96);
          }
          tmem_wait_ld();
          float* x = s + 64 * h;   // scaled scores of the half against the OLD reference
#pragma unroll
          for (int c = (h == 0 ? 16 : 0); c < 64; ++c) x[c] = fmaf(x[c], scale_log2, nm0);
          float xmax = x[0];
#pragma unroll
          for (int c = 1; c < 64; ++c) xmax = fmaxf(xmax, x[c]);
          const bool had_ref = m_ref != kNegInf;
          const float xm = xmax - kPOffset;            // the half's maximum above the reference, log2 units
          const bool need = !had_ref || xm > kLagLog2;
          const float delta = need ? xm : 0.f;
          const float f = (need && had_ref) ? ex2(-delta) : 1.0f;
          // everything accumulated so far -- O (earlier blocks, and the first half of this one if h == 1) and the sums --
          // moves to the new reference
          if (h == 0) mbar_wait(bar_o_full(tile), (blk - 1) & 1, 311 + tile);
          else mbar_wait(bar_o_half(tile), blk & 1, 313 + tile);
          tc_fence_after();
          tmem_wait_st();
#pragma unroll
          for (int cc = 0; cc < kD; cc += 32) {
            uint32_t o[32];
            tmem_ld_x32(tO + cc, o);
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * f);
            tmem_st_x32(tO + cc, o);
          }
          row_sum *= f;
          if (need) m_ref = (had_ref ? m_ref : 0.f) + delta * inv_scale_log2;
          ng2 = add2(ng2, pack2(-delta, -delta));
          float sum0 = 0.f;
          if (h == 0) {
#pragma unroll
            for (int c = 64; c < 80; ++c) s[c] -= delta;   // the software pipeline has scaled one chunk ahead
          } else {
#pragma unroll
            for (int c = 16; c < 64; ++c) s[c] *= f;         // the first half's exponentials wait for the late row sum
            float a0, a1, b0, b1;                            // (chunk 0 has been summed already)
            unpack2(acc_a, a0, a1);
            unpack2(acc_b, b0, b1);
            sum0 = ((a0 + a1) + (b0 + b1)) * f;
          }
#pragma unroll
          for (int c = 0; c < 64; c += 2) {
            const float e0 = ex2(x[c] - delta), e1 = ex2(x[c + 1] - delta);
            if (h == 0 && c < 16) {
              sum0 += e0 + e1;
            } else {
              x[c] = e0;
              x[c + 1] = e1;
            }
            if (c < 32) lo[c / 2] = Elt<T>::pack(e0, e1);
            else hi[(c - 32) / 2] = Elt<T>::pack(e0, e1);
          }
          acc_a = pack2(sum0, 0.f);
          acc_b = pack2(0.f, 0.f);
          xm_poly = -INFINITY;
        };
        stage_a(0, sc2);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (k == 3) tmem_wait_ld();   // one-pass blocks: columns 64..127 were still streaming out of TMEM
          if (k + 1 < 8) stage_a(k + 1, k == 0 ? sc2 : fma2(tok_prev, zero2, sc2));
          float ee[4];
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const int col = 16 * k + c;
            float e0, e1;
            if ((c >> 1) < 8 - (kFp8 ? FA_FWD_POLY8_FP8 : FA_FWD_POLY8)) {
              e0 = ex2(s[col]);
              e1 = ex2(s[col + 1]);
            } else {
              // (the polynomial wraps around above 2^127 instead of returning inf: the guard watches its inputs)
              if constexpr (kOnePass) xm_poly = fmax3(xm_poly, s[col], s[col + 1]);
              ex2_poly2_nz(pack2(s[col], s[col + 1]), e0, e1);
            }
            const uint64_t e2 = pack2(e0, e1);
            if (kSumLate && !(kOrGuard && col < 16)) {   // (OR-guard kernels: chunk 0 is summed here and keeps its scaled
              s[col] = e0;                               //  scores -- its S columns are the ones the first P store overwrites)
              s[col + 1] = e1;
            } else {
              if (c & 2) acc_b = add2(acc_b, e2);
              else acc_a = add2(acc_a, e2);
            }
            if (c == 10) {
              tok_prev = tok_cur;
              if constexpr (kOnePass) {   // an exponential against a stale reference may be inf: inf * 0 would poison the scale
                const float tk = fminf(e0, 1.0f);
                tok_cur = pack2(tk, tk);
              } else {
                tok_cur = e2;
              }
            }
            if constexpr (kFp8) {
              ee[c & 2] = e0;
              ee[(c & 2) + 1] = e1;
              if (c & 2) pk[(col & 31) / 4] = pack_e4m3x4(ee[0], ee[1], ee[2], ee[3]);
            } else {
              const uint32_t w = Elt<T>::pack(e0, e1);
              pk[(col & 31) / 2] = w;
              if constexpr (kOrGuard) {
                if (c & 2) g_or2 |= w;
                else g_or |= w;
              }
            }
          }
          if (k & 1) {   // 32 columns done
            const int cc = 32 * (k >> 1);
            if constexpr (kOrGuard) {
              if (cc == 32 || cc == 96) {
                if (opt) {   // guard: no exponential of this half reached 2^14 (inf included), and the polynomial saw no huge input
#ifdef FA_FWD_OG_NOFIX   // TIMING EXPERIMENT ONLY: the guard never fires
                  const bool bad = false;
#else
                  const bool bad = m_ref == kNegInf || ((g_or | g_or2) & 0x40004000u) != 0u || xm_poly > 100.f;
#endif
                  if (__any_sync(0xffffffffu, bad)) {
                    uint32_t lo[16];
                    if (cc == 32) {
                      fix_half_or(HalfIdx<0>{}, lo, pk);
                      tmem_st_x16(tS + 0, lo);
                    } else {
                      fix_half_or(HalfIdx<1>{}, lo, pk);
                      tmem_st_x16(tS + 32, lo);
                    }
                  }
                }
                g_or = 0;
                g_or2 = 0;
              }
              tmem_st_x16(tS + cc / 2, pk);
            } else if constexpr (kFp8) tmem_st_x8(tS + cc / 4, pk);
            else tmem_st_x16(tS + cc / 2, pk);
            if (cc + 32 == kPSplit || cc == 96) {  // first part / all of P is in TMEM: let the first / second part of P V start
              if constexpr (kOnePass && !kOrGuard) {
                if (opt) {   // guard: is every exponential of this half finite and small enough to keep the reference?
                  float a0, a1, b0, b1;
                  unpack2(acc_a, a0, a1);
                  unpack2(acc_b, b0, b1);
                  float hs = (a0 + a1) + (b0 + b1);
                  if (cc == 96) hs -= hs0;            // this half's own sum (inf / NaN survive the subtraction)
                  const bool bad = m_ref == kNegInf || !(hs <= kGuard) || xm_poly > 100.f;
                  if (__any_sync(0xffffffffu, bad)) hs = fix_half(cc == 96 ? 1 : 0, hs);
                  if (cc == 32) hs0 = hs;
                }
              }
              tmem_wait_st();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_p_full(tile) + (cc == 96 ? 16u : 0u));
              if (wq == 0 && lane == 0 && r == 0) FTS((cc == 96 ? 11 : 10) + 4 * tile);
            }
          }
        }
        if constexpr (kSumLate) {   // P is on its way: the row sum runs while the MMA thread works for this tile
          uint64_t acc_c = pack2(0.f, 0.f), acc_d = pack2(0.f, 0.f);
#pragma unroll
          for (int c = kOrGuard ? 16 : 0; c < 128; c += 8) {
            acc_a = add2(acc_a, pack2(s[c], s[c + 1]));
            acc_b = add2(acc_b, pack2(s[c + 2], s[c + 3]));
            acc_c = add2(acc_c, pack2(s[c + 4], s[c + 5]));
            acc_d = add2(acc_d, pack2(s[c + 6], s[c + 7]));
          }
          acc_a = add2(acc_a, acc_c);
          acc_b = add2(acc_b, acc_d);
        }
        } else
#endif
        {
#pragma unroll
        for (int cc = 0; cc < 128; cc += 32) {
          uint32_t pk[16];
          float ee[4];
          uint32_t keep_w = 0;
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float x0, x1, e0, e1;
            unpack2(fma2(pack2(s[cc + c], s[cc + c + 1]), sc2, ng2), x0, x1);
            if (((c >> 1) & 3) < 4 - kPolyPairs) {
              e0 = ex2(x0);
              e1 = ex2(x1);
            } else {
              if constexpr (kPolyDeg == 2) ex2_poly2_deg2(x0, x1, e0, e1);
              else ex2_poly2(x0, x1, e0, e1);
            }
            if (c & 2) acc_b = add2(acc_b, pack2(e0, e1));
            else acc_a = add2(acc_a, pack2(e0, e1));
            if constexpr (kDropout) {
              // the row sum keeps the undropped probabilities (softmax.h:128-187 runs before dropout.h:apply_dropout);
              // only the P that feeds P V is thinned, and 1/p_keep is folded into the final normalisation
              if (c == 0) keep_w = dropout_keep_row32(p.drop, uint32_t(qpos), uint32_t(col0 + cc) >> 5, uint32_t(bidb * p.h + head));
              e0 = ((keep_w >> c) & 1u) ? e0 : 0.f;
              e1 = ((keep_w >> (c + 1)) & 1u) ? e1 : 0.f;
            }
            if constexpr (kFp8) {
              ee[c & 2] = e0;
              ee[(c & 2) + 1] = e1;
              if (c & 2) pk[c / 4] = pack_e4m3x4(ee[0], ee[1], ee[2], ee[3]);
            } else {
              pk[c / 2] = Elt<T>::pack(e0, e1);
            }
          }
          if constexpr (kFp8) tmem_st_x8(tS + cc / 4, pk);
          else tmem_st_x16(tS + cc / 2, pk);
          if (cc + 32 == kPSplit || cc == 96) {  // first part / all of P is in TMEM: let the first / second part of P V start
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p_full(tile) + (cc == 96 ? 16u : 0u));
            if (wq == 0 && lane == 0 && r == 0) FTS((cc == 96 ? 11 : 10) + 4 * tile);
          }
        }
        }
        {
          float a0, a1, b0, b1;
          unpack2(acc_a, a0, a1);
          unpack2(acc_b, b0, b1);
          row_sum += (a0 + a1) + (b0 + b1);
        }
      }

      // ---------------------------------------------------------------- epilogue
      float inv = 0.f, lse_val;
      // no visible key at all: the reference max never left -inf (the polynomial exp2 of the pipelined loop returns
      // 2^-125 * p, not 0, for a masked score, so the row sum alone does not tell); !(row_sum > 0) also catches NaN
      const bool dead = m_ref == kNegInf || !(row_sum > 0.f);
      if (!dead) {
        inv = descale_v / row_sum;
        if constexpr (kDropout) inv *= p.drop.rp;  // dropout.h:apply_dropout rescale, folded into the normalisation
        lse_val = m_ref * scale_eff + __logf(row_sum) - kPOffset * 0.6931471805599453f;
      } else {
        // fully masked row: +inf for a final result, -inf for a split partial
        // (softmax.h:180, flash_fwd_kernel.h:544-582)
        lse_val = p.num_splits > 1 ? kNegInf : INFINITY;
      }
      // The O row leaves TMEM before it is stored, so that the next item's P V may overwrite it while the stores drain
      // (head dims above 128 in two halves of 128 columns: 256 live values would not fit the register budget).
      constexpr int kH = kD > 128 ? 128 : kD;
      if (nblocks > 0) {
        mbar_wait(bar_o_full(tile), (blk - 1) & 1, 320 + tile);
        tc_fence_after();
      }
#pragma unroll
      for (int h0 = 0; h0 < kD; h0 += kH) {
        uint32_t o[kH] = {};   // zeros if the tile had no visible block (a maybe-uninitialised array would be demoted to local memory)
        if (nblocks > 0) {
#pragma unroll
          for (int cc = 0; cc < kH; cc += 32) tmem_ld_x32(tO + h0 + cc, o + cc);
          tmem_wait_ld();
          if (h0 + kH == kD) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_o_free(tile));
          }
        }
        if (p.num_splits > 1) {
          const int64_t prow = ((int64_t(split) * p.b + bidb) * p.h + head) * p.sq + qpos;
          float* dst = p.o_accum + prow * kD + h0;
          if (row_valid) {
#pragma unroll
            for (int c = 0; c < kH; c += 4) {
              float4 v4;
              v4.x = __uint_as_float(o[c]) * inv;
              v4.y = __uint_as_float(o[c + 1]) * inv;
              v4.z = __uint_as_float(o[c + 2]) * inv;
              v4.w = __uint_as_float(o[c + 3]) * inv;
              *reinterpret_cast<float4*>(dst + c) = v4;
            }
            if (h0 == 0) p.lse_accum[prow] = lse_val;
          }
        } else {
          OutT* dst = reinterpret_cast<OutT*>(p.o) + (p.cu_seqlens_q ? 0 : bidb * p.o_batch_stride) +
                   int64_t(q_off + qpos) * p.o_row_stride + int64_t(head) * p.o_head_stride + h0;
          if (row_valid) {
#pragma unroll
            for (int c = 0; c < kH; c += 8) {
              if (h0 + c < p.d) {  // d is a multiple of 8
                uint4 v4;
                v4.x = Elt<T>::pack(__uint_as_float(o[c]) * inv, __uint_as_float(o[c + 1]) * inv);
                v4.y = Elt<T>::pack(__uint_as_float(o[c + 2]) * inv, __uint_as_float(o[c + 3]) * inv);
                v4.z = Elt<T>::pack(__uint_as_float(o[c + 4]) * inv, __uint_as_float(o[c + 5]) * inv);
                v4.w = Elt<T>::pack(__uint_as_float(o[c + 6]) * inv, __uint_as_float(o[c + 7]) * inv);
                *reinterpret_cast<uint4*>(dst + c) = v4;
              }
            }
            if (h0 == 0) {
              if (p.cu_seqlens_q) p.lse[int64_t(head) * p.total_q + q_off + qpos] = lse_val;
              else p.lse[(int64_t(bidb) * p.h + head) * p.sq + qpos] = lse_val;
            }
          }
        }
      }
      FITEM(1);
    }
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if (warp == Cfg::kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace fa
