// fa_fwd_wide_sm100.cuh — forward attention, head dim 128, S computed 256 keys at a time ("wide S").
//
// Why: the measured cost of a tcgen05.mma on B200 is `10 + N/2` clk (+31 when A comes from shared memory)
// per K = 16 step (experiments/mma_rate_probe.cu, profiles/r02_mma_rate_probe_n256.txt): the fixed part is paid per
// INSTRUCTION.  The two-tile kernel (fa_fwd_sm100.cuh) issues S = Q K^T with N = 128 (105 clk, 61 % of peak) and sits
// within 10 % of the resulting issue floor of 2864 clk per (256 queries x 128 keys) -- as does the reference's Blackwell
// kernel (flash_attn/cute/flash_fwd_sm100.py, same shape).  With N = 256 the same S costs 169 clk per 256 keys (76 %),
// i.e. 2536 clk for the same work (-11.5 %).  256 S columns per query tile leave TMEM room for ONE query tile per CTA
// (S 256 + P 128 + O 128 = 512 columns), so the overlap of softmax and tensor pipe comes from a different place than
// the two-tile ping-pong:
//
//   * P does NOT alias S here: S is free again as soon as the softmax warps have read it into registers, so the MMA
//     thread issues  QK(g+1)  BEFORE  PV(g)  and the tensor pipe works on the next block pair while the softmax of the
//     current one runs (a software pipeline of depth two over the stream of 256-key block pairs, ACROSS work items).
//   * the 256 columns of a block pair are split between kNG softmax warpgroups (same 128 query rows: thread t of
//     every group owns row t).  The kNG warps that own the same 32 rows agree on the row maximum through a small
//     shared-memory exchange + one named barrier of their own per block pair, keep partial row sums (joined in the
//     epilogue) and each rescales / writes out 1/kNG of the O columns.  kNG = 4 (64 columns per thread, 16 softmax
//     warps): with two groups the softmax of a block pair took 3.0k clk against the 2.5k of the tensor pipe
//     (profiles/r02_q_wide_timeline_2groups.log: load + max 0.4k, exchange 0.25k, exponentials 2.0k, hand-over
//     0.25k, both warps of a scheduler in the same phase at the same time); four warps per scheduler hide the latency
//     phases behind each other's arithmetic.
//
//   warp 0 .. 4 kNG - 1    softmax group h = warp / 4: S columns [kC h, kC h + kC) of the pair, kC = 256 / kNG
//                          -> P columns of the same keys, O columns [128 h / kNG, 128 (h + 1) / kNG)
//   warp 4 kNG             MMA issuer (one elected thread)
//   warp 4 kNG + 1         TMA producer: Q tile, K / V block pairs (256 rows x 128, 64 KB) through a 3-slot ring,
//                          in consumption order  K(g), V(g-1), K(g+1), V(g), ...
//
// Scope (the host picks this kernel only then, plan_fwd): dense (batch, seq, head, dim) tensors, fp16 / bf16, head
// dim <= 128, causal / local / no mask, no dropout / softcap / ALiBi / split-KV / paged KV / packed GQA.  Semantics
// as in fa_fwd_sm100.cuh (reference: csrc/flash_attn/src/flash_fwd_kernel.h, mask.h:172-173, softmax.h:169-186).
#pragma once
#include "fa_fwd_sm100.cuh"

#ifndef FA_WIDE_POLY8
#define FA_WIDE_POLY8 2   // polynomial pairs out of every 8 pairs (16 scores)
#endif
#ifndef FA_WIDE_GROUPS
#define FA_WIDE_GROUPS 2   // measured: 2 groups 1256, 4 groups 1212 TFLOP/s at the C2 shape (profiles/r02_r_wide_4groups_timing.log)
#endif
#ifndef FA_WIDE_ONEPASS
#define FA_WIDE_ONEPASS 0   // 1 (bf16): unmasked block pairs after an item's first take their exponentials against the row maximum of
                            // the blocks BEFORE them (the groups exchange a block pair's maximum after handing its P over; rows
                            // whose accumulators could overflow are redone exactly by fwd_wide_slow_row).  Correct (48 tests) but
                            // slower: 1222 vs 1280 TFLOP/s at the C2 shape -- the row maximum still has to be taken, and taken
                            // inside the exponential loop it costs more than in front of it (profiles/r02_u_wide_onepass.log)
#endif
#ifndef FA_WIDE_DIAG
#define FA_WIDE_DIAG 0   // TIMING EXPERIMENTS ONLY (wrong results): 1 = exponentials for half of the columns, 3 = every CTA streams the K / V of (batch 0, head 0) (L2-resident)
#endif

// timeline stamps of CTA 0's first work item (FA_FWD_TS builds; tools/probe_fwd_wide_ts.py)
#define WTS(cond_, blk_, ev_)                     \
  do {                                            \
    if (cond_) {                                  \
      const int it = (blk_);                      \
      (void)it;                                   \
      FTS(ev_);                                   \
    }                                             \
  } while (0)

namespace fa {

template <int kNG>
struct FwdWideCfg {
  static_assert(kNG == 2 || kNG == 4, "softmax groups");
  static constexpr int kC = 256 / kNG;            // S columns per softmax thread
  static constexpr int kOC = 128 / kNG;           // O columns per softmax thread
  static constexpr int kW = 4 * kNG;              // softmax warps
  static constexpr int kThreads = 128 * (kNG + 1);
  // register pool = kThreads x (65536 / kThreads rounded down to 8): 384 x 168 or 640 x 96
  static constexpr int kRegsSoftmax = kNG == 2 ? 216 : 104, kRegsOther = kNG == 2 ? 72 : 64;
  static constexpr int kSlots = 3;
  static constexpr int kSlotBytes = 65536;        // a K or V block pair: 2 panels x (256 rows x 128 B)
  static constexpr int kPanelBytesKV = 32768;
  static constexpr int kPanelBytesQ = 16384;
  static constexpr int kSQ = 0;
  static constexpr int kSKV = 32768;
  static constexpr int kSBar = kSKV + kSlots * kSlotBytes;
  static constexpr int kNumBars = 20;
  static constexpr int kSMisc = kSBar + 8 * kNumBars;     // tmem pointer (4 B), pad, 2 scheduler items
  static constexpr int kSExch = kSMisc + 32;              // 4 slices (one per `warp & 3`) of kNG * 128 bytes: row maxima / row sums
  static constexpr int kSmemBytes = kSExch + 2048;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// Exact per-row fallback of the one-pass mode (cold: a score more than 2^214 above the row maximum of the blocks before it
// overflowed the accumulators).  One thread recomputes its query row against keys [col_lo, col_hi) straight from global
// memory, two passes, fp32, and writes kOC output columns from column oc0 (+ the LSE if write_lse).
template <typename T, int kOC>
__device__ __noinline__ void fwd_wide_slow_row(const FwdKernelParams& p, int b, int head, int hk, int qpos, int col_lo, int col_hi,
                                               int oc0, bool write_lse) {
  const T* q = reinterpret_cast<const T*>(p.q_raw) + b * p.q_bs + int64_t(qpos) * p.q_rs + int64_t(head) * p.q_hs;
  const T* kb = reinterpret_cast<const T*>(p.k_raw) + b * p.k_bs + int64_t(hk) * p.k_hs;
  const T* vb = reinterpret_cast<const T*>(p.v_raw) + b * p.v_bs + int64_t(hk) * p.v_hs;
  auto dot = [&](int j) {
    const T* k = kb + int64_t(j) * p.k_rs;
    float acc = 0.f;
    for (int c = 0; c < p.d; ++c) acc = fmaf(float(q[c]), float(k[c]), acc);
    return acc * p.scale;
  };
  float m = -INFINITY;
  for (int j = col_lo; j < col_hi; ++j) m = fmaxf(m, dot(j));
  float o[kOC];
#pragma unroll
  for (int c = 0; c < kOC; ++c) o[c] = 0.f;
  float l = 0.f;
  for (int j = col_lo; j < col_hi; ++j) {
    const float pj = __expf(dot(j) - m);
    l += pj;
    const T* v = vb + int64_t(j) * p.v_rs + oc0;
#pragma unroll
    for (int c = 0; c < kOC; ++c)
      if (oc0 + c < p.d) o[c] = fmaf(pj, float(v[c]), o[c]);
  }
  const bool dead = !(l > 0.f) || m == -INFINITY;
  const float inv = dead ? 0.f : 1.0f / l;
  T* dst = reinterpret_cast<T*>(p.o) + b * p.o_batch_stride + int64_t(qpos) * p.o_row_stride + int64_t(head) * p.o_head_stride + oc0;
#pragma unroll
  for (int c = 0; c < kOC; ++c)
    if (oc0 + c < p.d) dst[c] = T(o[c] * inv);
  if (write_lse) p.lse[(int64_t(b) * p.h + head) * p.sq + qpos] = dead ? INFINITY : m + __logf(l);
}

template <typename T, int kNG>
__global__ void __launch_bounds__(FwdWideCfg<kNG>::kThreads, 1) fa_fwd_wide_kernel(const __grid_constant__ FwdKernelParams p) {
  using Cfg = FwdWideCfg<kNG>;
  constexpr int kD = 128;
  constexpr int kC = Cfg::kC, kOC = Cfg::kOC, kW = Cfg::kW;
  constexpr float kNegInf = -INFINITY;
  constexpr float kLagLog2 = 8.0f;
  // One-pass mode (bf16 only: P needs the exponent range).  P is written as 2^(x - 100), x = (s - m_ref) * scale_log2: the power of
  // two cancels exactly between O and the row sum, scores within 2^-26 of the reference stay normal bf16 numbers, and a block pair may
  // exceed the reference (the maximum of the blocks BEFORE it) by up to 2^214 before an accumulator overflows -- rows beyond that are
  // caught in the epilogue (non-finite row sum) and recomputed exactly by fwd_wide_slow_row.
  constexpr bool kOnePass = FA_WIDE_ONEPASS && Elt<T>::kFmt == 1;
  constexpr float kPOffset = kOnePass ? -100.0f : 0.0f;

  const int G = int(gridDim.x), cta = int(blockIdx.x);
  const int total_items = p.total_items;
  const int rounds = (total_items + G - 1) / G;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t sQ = smem_base + Cfg::kSQ, sKV = smem_base + Cfg::kSKV, sBar = smem_base + Cfg::kSBar;
  enum { Q_FULL = 0, Q_EMPTY = 1, KV_FULL0 = 2, KV_EMPTY0 = 5, S_FULL = 8, S_READ = 9, P_FULL = 10, PV_DONE = 11,
         O_FREE = 12, SCHED_FULL0 = 13, SCHED_EMPTY0 = 15 };
  auto bar = [&](int i) { return sBar + 8u * i; };
  const uint32_t sTmemPtr = smem_base + Cfg::kSMisc;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + Cfg::kSMisc);
  // exchange area: one slice of 512 bytes per group of partner warps (same `warp & 3`: the warps that own the same 32
  // rows) -- a slice is only ever touched by warps that synchronise on the same named barrier.  Inside a slice the row maxima
  // ([2 parities][kNG][32 lanes]: fp32 for two groups, bf16 for four) and the row sums of the epilogue (float [kNG][32 lanes])
  // share the bytes.

  FwdSched sched;
  sched.G = G; sched.cta = cta; sched.total = total_items; sched.rounds = rounds; sched.k = 0;
  sched.bar_full0 = bar(SCHED_FULL0); sched.bar_empty0 = bar(SCHED_EMPTY0); sched.item_addr = sTmemPtr + 8u;
  sched.dyn = p.sched_counter != nullptr;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarpMma = kW, kWarpTma = kW + 1;
  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&p.tm_q);
    tma_prefetch_desc(&p.tm_k256);
    tma_prefetch_desc(&p.tm_v256);
  }
  if (warp == kWarpMma) {
    if (lane == 0) {
      mbar_init(bar(Q_FULL), 1);
      mbar_init(bar(Q_EMPTY), 1);
      for (int s = 0; s < Cfg::kSlots; ++s) {
        mbar_init(bar(KV_FULL0 + s), 1);
        mbar_init(bar(KV_EMPTY0 + s), 1);
      }
      mbar_init(bar(S_FULL), 1);
      mbar_init(bar(S_READ), kW);     // every softmax warp has its S columns in registers
      mbar_init(bar(P_FULL), kW);
      mbar_init(bar(PV_DONE), 1);
      mbar_init(bar(O_FREE), kW);
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar(SCHED_FULL0 + i), 1);
        mbar_init(bar(SCHED_EMPTY0 + i), 1 + kW);   // MMA thread + softmax warps
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(sTmemPtr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // block-pair range of a work item: 128-key blocks [n_min, n_max) -> pairs [p_lo, p_hi), streamed right to left
  auto pair_range = [](const FwdWork& w, int& p_lo, int& p_hi) {
    p_lo = w.n_min >> 1;
    p_hi = (w.n_max + 1) >> 1;
    if (w.nblocks == 0) p_lo = p_hi = 0;
  };

  if (warp >= kW) {
    reg_dealloc<Cfg::kRegsOther>();
    // ================================================================== TMA producer
    if (warp == kWarpTma) {
      if (elect_one()) {
        int ring = 0, q_cnt = 0;
        bool pend = false;
        int pv_row = 0, pv_hk = 0, pv_b = 0;
        int nxt = -1;
        bool drawn = false;
        auto load_pair = [&](const CUtensorMap* tm, int row, int hk, int b) {
          const int s = ring % Cfg::kSlots;
          const uint32_t ph = (ring / Cfg::kSlots) & 1;
          mbar_wait(bar(KV_EMPTY0 + s), ph ^ 1, 100 + s);
          mbar_arrive_expect_tx(bar(KV_FULL0 + s), Cfg::kSlotBytes);
#pragma unroll
          for (int pn = 0; pn < 2; ++pn)
            tma_load_4d(sKV + s * Cfg::kSlotBytes + pn * Cfg::kPanelBytesKV, tm, bar(KV_FULL0 + s), pn * 64, row, hk, b);
          ++ring;
        };
        auto advance = [&]() -> int {
          if (!sched.dyn) return sched.next(false, 0);
          if (!drawn) nxt = sched.draw(p.sched_counter);
          drawn = false;
          return sched.publish(nxt);
        };
        for (int L = sched.first(); L >= 0; L = advance()) {
          FwdWork w;
          if (!fwd_decode_work<128, 1>(p, L, w)) continue;
          int p_lo, p_hi;
          pair_range(w, p_lo, p_hi);
          const int N = p_hi - p_lo;
          if (N == 0) continue;
          if (q_cnt > 0) mbar_wait(bar(Q_EMPTY), (q_cnt - 1) & 1, 110);
          ++q_cnt;
          mbar_arrive_expect_tx(bar(Q_FULL), 32768);
#pragma unroll
          for (int pn = 0; pn < 2; ++pn)
            tma_load_5d(sQ + pn * Cfg::kPanelBytesQ, &p.tm_q, bar(Q_FULL), pn * 64, 0, w.row0, w.slot, w.bidb);
          for (int n = 0; n < N; ++n) {
            if (sched.dyn && !drawn && n + 2 >= N) {
              nxt = sched.draw(p.sched_counter);
              drawn = true;
            }
            const int row = (p_hi - 1 - n) * 256;
#if FA_WIDE_DIAG == 3
            w.hk_idx = 0; w.bidb = 0;
#endif
            load_pair(&p.tm_k256, row, w.hk_idx, w.bidb);
            if (pend) load_pair(&p.tm_v256, pv_row, pv_hk, pv_b);
            pend = true; pv_row = row; pv_hk = w.hk_idx; pv_b = w.bidb;
          }
        }
        if (pend) load_pair(&p.tm_v256, pv_row, pv_hk, pv_b);
      }
      __syncwarp();
    }
    // ================================================================== MMA issuer
    else if (warp == kWarpMma) {
      if (elect_one()) {
        constexpr uint32_t fmt = Elt<T>::kFmt;
        constexpr uint32_t idesc_qk = make_idesc(fmt, fmt, 0, 0, 128, 256);
        constexpr uint32_t idesc_pv = make_idesc(fmt, fmt, 0, 1, 128, kD);
        constexpr uint64_t desc_k_major = smem_desc_sw128(16, 1024);
        constexpr uint64_t desc_mn_major = smem_desc_sw128(Cfg::kPanelBytesKV, 1024);
        constexpr uint32_t hi_k = uint32_t(desc_k_major >> 32), hi_mn = uint32_t(desc_mn_major >> 32);
        constexpr uint32_t lo_mn = uint32_t(desc_mn_major);
        auto lo_of = [](uint32_t addr) { return (addr & 0x3FFFFu) >> 4; };
        auto mk = [](uint32_t hi, uint32_t lo) { return (uint64_t(hi) << 32) | lo; };
        const uint32_t tS = tmem_base, tP = tmem_base + 256, tO = tmem_base + 384;

        int ring = 0, g = 0, items = 0;
        bool pend = false;
        bool pd_first = false;
        int pd_g = 0, pd_item = 0, pd_n = 0;
        auto do_pv = [&]() {
          const int rv = ring++;
          const int sv = rv % Cfg::kSlots;
          mbar_wait(bar(KV_FULL0 + sv), (rv / Cfg::kSlots) & 1, 220);
          if (pd_first && pd_item > 0) mbar_wait(bar(O_FREE), (pd_item - 1) & 1, 225);   // the previous item's O has been read out
          const uint32_t lb = lo_of(sKV + sv * Cfg::kSlotBytes) | lo_mn;
          mbar_wait(bar(P_FULL), pd_g & 1, 230);
          WTS(pd_item == 0, pd_n, 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 16; ++kk)   // 16 keys per step
            mma_ts<false>(tO, tP + kk * 8, mk(hi_mn, lb + kk * (2048 >> 4)), idesc_pv, (!pd_first || kk > 0) ? 1u : 0u);
          mma_commit(bar(PV_DONE));
          mma_commit(bar(KV_EMPTY0 + sv));
          WTS(pd_item == 0, pd_n, 2);
        };
        for (int L = sched.first(); L >= 0; L = sched.next(false, 0)) {
          FwdWork w;
          if (!fwd_decode_work<128, 1>(p, L, w)) continue;
          int p_lo, p_hi;
          pair_range(w, p_lo, p_hi);
          const int N = p_hi - p_lo;
          if (N == 0) continue;
          mbar_wait(bar(Q_FULL), items & 1, 200);
          for (int n = 0; n < N; ++n) {
            const int rk = ring++;
            const int sk_ = rk % Cfg::kSlots;
            mbar_wait(bar(KV_FULL0 + sk_), (rk / Cfg::kSlots) & 1, 210);
            if (g > 0) mbar_wait(bar(S_READ), (g - 1) & 1, 211);   // S of the previous block pair is in registers
            tc_fence_after();
            {
              const uint32_t la = lo_of(sQ) | uint32_t(desc_k_major);
              const uint32_t lb = lo_of(sKV + sk_ * Cfg::kSlotBytes) | uint32_t(desc_k_major);
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {   // 16 head-dim elements (32 B) per step
                const uint32_t offa = ((kk / 4) * Cfg::kPanelBytesQ + (kk % 4) * 32) >> 4;
                const uint32_t offb = ((kk / 4) * Cfg::kPanelBytesKV + (kk % 4) * 32) >> 4;
                mma_ss<false>(tS, mk(hi_k, la + offa), mk(hi_k, lb + offb), idesc_qk, kk > 0);
              }
            }
            mma_commit(bar(S_FULL));
            mma_commit(bar(KV_EMPTY0 + sk_));
            WTS(items == 0, n, 0);
            if (n == N - 1) mma_commit(bar(Q_EMPTY));
            if (pend) do_pv();
            pend = true; pd_first = n == 0; pd_g = g; pd_item = items; pd_n = n;
            ++g;
          }
          ++items;
        }
        if (pend) do_pv();
      }
      __syncwarp();
    }
  } else {
    // ================================================================== softmax warpgroups
    reg_alloc<Cfg::kRegsSoftmax>();
    const int h = warp >> 2;          // which kC columns of the block pair
    const int wq = warp & 3;
    const int trow = wq * 32 + lane;  // row of the query tile == TMEM lane
    const uint32_t lane_sel = uint32_t(wq * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel + kC * h;
    const uint32_t tP = tmem_base + lane_sel + 256 + (kC / 2) * h;
    const uint32_t tO = tmem_base + lane_sel + 384 + kOC * h;
    uint8_t* exch_slice = smem_raw + Cfg::kSExch + wq * 512;
    float* exch_sum = reinterpret_cast<float*>(exch_slice);
    const uint32_t pbar = 2 + wq;     // named barrier of the kNG warps that own these 32 rows
    auto partner_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(pbar), "n"(32 * kNG) : "memory"); };
    int g = 0;
    for (int L = sched.first(); L >= 0; L = sched.next(true, lane)) {
      FwdWork w;
      if (!fwd_decode_work<128, 1>(p, L, w)) continue;
      int p_lo, p_hi;
      pair_range(w, p_lo, p_hi);
      const int N = p_hi - p_lo;
      const int bidb = w.bidb, head = w.slot, shift = w.shift, sk_b = w.sk_b;
      const int qpos = w.row0 + trow;
      const bool row_valid = qpos < w.rows_total;
      const float scale_log2 = p.scale_log2;
      const int t_qlo = w.row0;
      const int t_qhi = w.row0 + 127 < w.rows_total - 1 ? w.row0 + 127 : w.rows_total - 1;
      int col_hi = sk_b, col_lo = 0;
      if (p.has_right) { int l = qpos + 1 + shift + p.wr; col_hi = l < col_hi ? l : col_hi; }
      if (p.has_left) { int l = qpos + shift - p.wl; col_lo = l > 0 ? l : 0; }

      float m_ref = kNegInf, row_sum = 0.f;
      float mx_pending = kNegInf;   // one-pass mode: row maximum of the last block pair, not yet folded into the reference
      bool sticky = false;          // one-pass mode: an accumulator of this row may have overflowed
      for (int n = 0; n < N; ++n, ++g) {
        const int col0 = (p_hi - 1 - n) * 256 + kC * h;
        mbar_wait(bar(S_FULL), g & 1, 300 + h);
        const bool ts_on = wq == 0 && lane == 0 && g == n && h < 2;   // (g == n: the CTA's first item)
        WTS(ts_on, n, 4 + 6 * h);
        tc_fence_after();
        uint32_t su[kC];
        float* s = reinterpret_cast<float*>(su);
        // masks: warp-uniform per group; a block pair is one-pass only if NO group has to mask (uniform across the partners)
        bool need_mask = col0 + kC > sk_b;
        if (p.has_right) need_mask |= col0 + kC > t_qlo + 1 + shift + p.wr;
        if (p.has_left) need_mask |= col0 < t_qhi + shift - p.wl;
        bool any_mask;
        {
          const int pc0 = (p_hi - 1 - n) * 256;
          any_mask = pc0 + 256 > sk_b;
          if (p.has_right) any_mask |= pc0 + 256 > t_qlo + 1 + shift + p.wr;
          if (p.has_left) any_mask |= pc0 < t_qhi + shift - p.wl;
        }
        // ... and only once every row of the warp has a finite reference (a row's first visible block pair is two-pass: against
        // "reference 0" strongly negative scores would underflow).  Same answer in every partner: same m_ref / mx_pending.
        const bool have_ref = fmaxf(m_ref, mx_pending) != kNegInf;
        const bool onepass = kOnePass && n > 0 && !any_mask && __all_sync(0xffffffffu, have_ref);
#pragma unroll
        for (int c = 0; c < kC / 2; c += 32) tmem_ld_x32(tS + c, su + c);
        tmem_wait_ld();
        WTS(ts_on, n, 5 + 6 * h);
#pragma unroll
        for (int c = kC / 2; c < kC; c += 32) tmem_ld_x32(tS + c, su + c);   // streams out of TMEM while the first half is worked on
        float mx = kNegInf;
        if (onepass) {
          // the reference moves to the maximum of the blocks before this one (exchanged after the previous hand-over)
          mx = mx_pending;
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(S_READ));
        } else if (!need_mask) {
          float m0 = fmax3(s[0], s[1], s[2]), m1 = fmax3(s[3], s[4], s[5]);
          float m2 = fmax3(s[6], s[7], s[8]), m3 = fmax3(s[9], s[10], s[11]);
#pragma unroll
          for (int c = 12; c + 7 < kC / 2; c += 8) {
            m0 = fmax3(m0, s[c], s[c + 1]);
            m1 = fmax3(m1, s[c + 2], s[c + 3]);
            m2 = fmax3(m2, s[c + 4], s[c + 5]);
            m3 = fmax3(m3, s[c + 6], s[c + 7]);
          }
          m0 = fmax3(m0, s[kC / 2 - 4], s[kC / 2 - 3]);
          m1 = fmax3(m1, s[kC / 2 - 2], s[kC / 2 - 1]);
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(S_READ));   // the MMA thread may overwrite S with the next block pair
#pragma unroll
          for (int c = kC / 2; c + 7 < kC; c += 8) {
            m0 = fmax3(m0, s[c], s[c + 1]);
            m1 = fmax3(m1, s[c + 2], s[c + 3]);
            m2 = fmax3(m2, s[c + 4], s[c + 5]);
            m3 = fmax3(m3, s[c + 6], s[c + 7]);
          }
          mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        } else {
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(S_READ));
          const int hi = col_hi - col0, lo = col_lo - col0;
#pragma unroll
          for (int c = 0; c < kC; ++c) s[c] = (c >= hi || c < lo) ? kNegInf : s[c];
          float m0 = fmax3(s[0], s[1], s[2]), m1 = fmax3(s[3], s[4], s[5]);
#pragma unroll
          for (int c = 6; c + 3 < kC; c += 4) {
            m0 = fmax3(m0, s[c], s[c + 1]);
            m1 = fmax3(m1, s[c + 2], s[c + 3]);
          }
          mx = fmax3(m0, m1, fmaxf(s[kC - 2], s[kC - 1]));
        }
        // ---- the kNG warps that own these 32 rows agree on the row maximum of a block pair through their slice of the exchange
        // area, two parities deep: one named barrier per block pair is enough.
        // Two-pass block pairs exchange BEFORE the exponentials, one-pass block pairs AFTER the hand-over of P (below).
        auto exchange = [&](float mine) {
          float mr = kNegInf;
          if constexpr (kNG == 2) {   // exact maxima: the reference of a two-pass block pair IS its row maximum (P of it = 1)
            float* ex = reinterpret_cast<float*>(exch_slice) + (g & 1) * (kNG * 32);
            ex[h * 32 + lane] = mine;
            partner_sync();
#pragma unroll
            for (int i = 0; i < kNG; ++i) mr = fmaxf(mr, ex[i * 32 + lane]);
          } else {                    // four groups: 512 bytes per slice only hold bf16, rounded UP (a reference slightly above the
            uint16_t* ex = reinterpret_cast<uint16_t*>(exch_slice) + (g & 1) * (kNG * 32);   // maximum is still a valid reference)
            const uint32_t bits = __float_as_uint(mine);
            ex[h * 32 + lane] = uint16_t((mine >= 0.f ? bits + 0xFFFFu : bits) >> 16);
            partner_sync();
#pragma unroll
            for (int i = 0; i < kNG; ++i) mr = fmaxf(mr, __uint_as_float(uint32_t(ex[i * 32 + lane]) << 16));
          }
          return mr;
        };
        WTS(ts_on, n, 6 + 6 * h);
        if (!onepass) {
          mx = fmaxf(exchange(mx), mx_pending);   // (a one-pass block pair before this one may have left its maximum pending)
          mx_pending = kNegInf;
        }
        WTS(ts_on, n, 7 + 6 * h);
        // ---- online softmax with a lazily updated reference (all groups take identical decisions: same inputs)
        const float m_new = fmaxf(m_ref, mx);
        if (n == 0) {
          m_ref = m_new;
        } else {
          bool need = (m_new - m_ref) * scale_log2 > kLagLog2;   // NaN (-inf - -inf) compares false
          if (m_ref == kNegInf) {
            m_ref = m_new;
            need = false;
          }
          if (__any_sync(0xffffffffu, need)) {
            mbar_wait(bar(PV_DONE), (g - 1) & 1, 310 + h);   // P V of the previous block pair has landed
            tc_fence_after();
            // The factor is applied in two halves: in one-pass mode O may hold numbers as large as 2^114 that a reference jump of
            // 2^200 brings back to 2^-86 -- the product is representable, 2^-200 itself is not (flushed to zero).
            const float f = need ? ex2(0.5f * (m_ref - m_new) * scale_log2) : 1.0f;
#pragma unroll
            for (int cc = 0; cc < kOC; cc += 32) {   // this group's share of the O columns
              uint32_t o[32];
              tmem_ld_x32(tO + cc, o);
              tmem_wait_ld();
#pragma unroll
              for (int c = 0; c < 32; ++c) o[c] = __float_as_uint((__uint_as_float(o[c]) * f) * f);
              tmem_st_x32(tO + cc, o);
            }
            tmem_wait_st();
            tc_fence_before();
            // every share of the O rows must be rescaled before ANY group hands its P to the MMA thread (the first P V step
            // accumulates into all 128 columns)
            partner_sync();
            row_sum = (row_sum * f) * f;
            if (need) m_ref = m_new;
          }
        }
        const float neg_ms = (m_ref == kNegInf) ? kPOffset : kPOffset - m_ref * scale_log2;
        // ---- P = exp2(S * scale_log2 - m * scale_log2): software-pipelined as in fa_fwd_sm100.cuh (scaling of chunk k+1
        // made to depend on an exponential of chunk k-1, so that MUFU and FMA pipes run side by side); the packed words
        // stay in registers (they replace the scores they came from) until the P buffer is free.  One-pass block pairs pick
        // up their own row maximum on the way (raw scores, before the scaling).
        const uint64_t sc2 = pack2(scale_log2, scale_log2);
        const uint64_t ng2 = pack2(neg_ms, neg_ms);
        const uint64_t zero2 = pack2(0.f, 0.f);
        uint64_t acc_a = zero2, acc_b = zero2;
        uint32_t pk[kC / 2];
        uint64_t tok_prev = zero2, tok_cur = zero2;
        float lm0 = kNegInf, lm1 = kNegInf;
        auto stage_a = [&](int k, uint64_t sc) {
          if constexpr (kOnePass) {   // (unconditional: a branch here would cut the software pipeline into basic blocks)
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
              lm0 = fmax3(lm0, s[16 * k + c], s[16 * k + c + 1]);
              lm1 = fmax3(lm1, s[16 * k + c + 2], s[16 * k + c + 3]);
            }
          }
#pragma unroll
          for (int c = 0; c < 16; c += 2)
            unpack2(fma2(pack2(s[16 * k + c], s[16 * k + c + 1]), sc, ng2), s[16 * k + c], s[16 * k + c + 1]);
        };
        stage_a(0, sc2);
#pragma unroll
        for (int k = 0; k < kC / 16; ++k) {
#if FA_WIDE_DIAG == 1
          if (k & 1) {
#pragma unroll
            for (int c = 0; c < 8; ++c) pk[8 * k + c] = 0;
            continue;
          }
#endif
          if (k + 1 < kC / 16) stage_a(k + 1, k == 0 ? sc2 : fma2(tok_prev, zero2, sc2));
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const int col = 16 * k + c;
            float e0, e1;
            if ((c >> 1) < 8 - FA_WIDE_POLY8) {
              e0 = ex2(s[col]);
              e1 = ex2(s[col + 1]);
            } else {
              ex2_poly2_nz(pack2(s[col], s[col + 1]), e0, e1);
            }
            const uint64_t e2 = pack2(e0, e1);
            if (c & 2) acc_b = add2(acc_b, e2);
            else acc_a = add2(acc_a, e2);
            if (c == 10) {
              tok_prev = tok_cur;
              if constexpr (kOnePass) {   // (an exponential against a lagging reference may be inf: inf * 0 would poison the scale)
                const float tk = fminf(e0, 1.0f);
                tok_cur = pack2(tk, tk);
              } else {
                tok_cur = e2;
              }
            }
            pk[col / 2] = Elt<T>::pack(e0, e1);
          }
        }
        WTS(ts_on, n, 8 + 6 * h);
        if (g > 0) mbar_wait(bar(PV_DONE), (g - 1) & 1, 320 + h);   // the previous block pair's P V has read P
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kC / 2; c += 16) tmem_st_x16(tP + c, pk + c);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(P_FULL));
        WTS(ts_on, n, 9 + 6 * h);
        {
          float a0, a1, b0, b1;
          unpack2(acc_a, a0, a1);
          unpack2(acc_b, b0, b1);
          const float bsum = (a0 + a1) + (b0 + b1);
          if constexpr (kOnePass) sticky |= !(bsum <= 1.0e18f);   // a score > 2^160 above the lagging reference: the row is redone exactly
          row_sum += bsum;
        }
        if (onepass) {
          const float lm = fmaxf(lm0, lm1);
          // (the polynomial exp2 wraps around above 2^127 instead of returning inf: watch the largest exponent itself)
          sticky |= fmaf(lm, scale_log2, neg_ms) > 60.0f;
          mx_pending = exchange(lm);   // off the S -> P chain; applied at the next block pair
        }
      }

      // ---------------------------------------------------------------- epilogue: join the row sums, write this group's O columns
      partner_sync();   // (the sums reuse the buffer of the maxima: every partner has read the last block pair's)
      if (sticky) row_sum = INFINITY;   // every partner then sees a non-finite joined sum
      exch_sum[h * 32 + lane] = row_sum;
      partner_sync();
      row_sum = 0.f;
#pragma unroll
      for (int i = 0; i < kNG; ++i) row_sum += exch_sum[i * 32 + lane];
      partner_sync();   // (the next item's first exchange reuses the buffer)
      float inv = 0.f, lse_val;
      const bool dead = m_ref == kNegInf || !(row_sum > 0.f);
      if (!dead) {
        inv = 1.0f / row_sum;
        lse_val = m_ref * p.scale + __logf(row_sum) - kPOffset * 0.6931471805599453f;
      } else {
        lse_val = INFINITY;   // fully masked row (softmax.h:180)
      }
      uint32_t o[kOC] = {};
      if (N > 0) {
        mbar_wait(bar(PV_DONE), (g - 1) & 1, 330 + h);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kOC; c += 32) tmem_ld_x32(tO + c, o + c);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(O_FREE));
      }
      bool overflowed = false;
      if constexpr (kOnePass) overflowed = !(row_sum <= 1.0e18f);   // inf, NaN or absurd: the same joined sum in every partner
      if (overflowed) {
        if (row_valid) fwd_wide_slow_row<T, kOC>(p, bidb, head, w.hk_idx, qpos, col_lo, col_hi, kOC * h, h == 0);
      } else if (row_valid) {
        T* dst = reinterpret_cast<T*>(p.o) + bidb * p.o_batch_stride + int64_t(qpos) * p.o_row_stride +
                 int64_t(head) * p.o_head_stride + kOC * h;
#pragma unroll
        for (int c = 0; c < kOC; c += 8) {
          if (kOC * h + c < p.d) {   // d is a multiple of 8
            uint4 v4;
            v4.x = Elt<T>::pack(__uint_as_float(o[c]) * inv, __uint_as_float(o[c + 1]) * inv);
            v4.y = Elt<T>::pack(__uint_as_float(o[c + 2]) * inv, __uint_as_float(o[c + 3]) * inv);
            v4.z = Elt<T>::pack(__uint_as_float(o[c + 4]) * inv, __uint_as_float(o[c + 5]) * inv);
            v4.w = Elt<T>::pack(__uint_as_float(o[c + 6]) * inv, __uint_as_float(o[c + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + c) = v4;
          }
        }
        if (h == 0) p.lse[(int64_t(bidb) * p.h + head) * p.sq + qpos] = lse_val;
      }
    }
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace fa
