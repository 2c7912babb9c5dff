// fa_bwd2_sm100.cuh — backward main loop on a CTA PAIR (cluster of 2, tcgen05 cta_group::2), head dim 128.
//
// Why a pair: the single-CTA kernel (fa_bwd_sm100.cuh) handles one tile at a time in a strictly serial chain
// (S^T -> P^T -> dV, dP^T -> dS^T -> dK, dQ -> drain -> next dP^T): measured period 5.3k clk per 128x128 tile for 3.7k clk
// of tcgen05.mma issue cost (experiments/mma_rate_probe.cu: TS 74 clk, SS 105 clk per 128x128x16), because its 512 TMEM
// columns are full (S^T/P^T/dS^T 128, dP^T/dQ 128, dV 128, dK 128) and no MMA of the next tile can start early.
// A pair produces ONE dQ tile per 256 keys, so dQ needs 64 TMEM columns per CTA instead of 128 -- that is the
// room the schedule below needs (timeline in profiles/r01_pair_timeline_scheduleD.txt).
//
//   two CTAs = two adjacent 128-key blocks of one (batch, kv head), iterating the same (q head, query block) list
//   S^T  = K Q^T       pair SS  M=256 (128 keys / CTA)  N=128 queries   B: CTA r holds query rows [64r, 64r+64)
//   dP^T = V dO^T      pair SS  same shape
//   dV  += P^T dO      pair TS  M=256  N=128 (d)                        B: CTA r holds d columns [64r, 64r+64)
//   dK  += dS^T Q      pair TS  same shape
//   dQ   = dS K        pair SS  M=128 (64 queries / CTA)  N=128 (d)  K-dim = 256 keys of the pair
//                      A: CTA r holds dS^T[256 keys, query half r] -- the half computed by the peer is staged and shipped
//                      with cp.async.bulk.shared::cluster (tx-counted on the receiver's mbarrier);
//                      B: CTA r holds K[256 keys, d half r]
//   issue order per tile n:  dV(n), S^T(n+1), dK(n), dP^T(n+1), dQ(n)   (TMEM tenants: see the MMA issuer below)
//
// Same semantics, statistics, masks, dropout, GQA loop and epilogue as fa_bwd_kernel (reference:
// csrc/flash_attn/src/flash_bwd_kernel.h:80-795); TMEM layouts of pair MMAs measured with
// experiments/cta_pair_probe.cu (see fa_ptx.cuh).
#pragma once
#include "fa_bwd_sm100.cuh"

namespace fa {

#ifndef FA_BWD2_POLY
#define FA_BWD2_POLY 0   // n > 0: of every n groups of 4 scores, one group computes 2 of its exp2 on the FMA pipe
#endif

#ifndef FA_BWD2_LOCAL_DS
#define FA_BWD2_LOCAL_DS 0   // 1: TIMING EXPERIMENT ONLY (wrong dQ): every warp stores its dS^T rows into its own CTA
#endif
#ifndef FA_BWD2_TS
#define FA_BWD2_TS 0     // 1: CTA pair (0,0,0) records clock64 stamps of its pipeline events (timeline debugging build)
#endif
#if FA_BWD2_TS
__device__ unsigned long long g_bwd2_span[3 * 8192];    // per CTA: smid, globaltimer at entry, at exit
__device__ unsigned int g_bwd2_span_n;
__device__ unsigned long long g_bwd2_ts[2 * 16 * 32];   // [rank][iteration - 8][event]
#define TS(ev)                                                                                          \
  do {                                                                                                  \
    if (ts_on && it >= 8 && it < 24)                                                   \
      g_bwd2_ts[(int(rank) * 16 + (it - 8)) * 32 + (ev)] = clock64();                                    \
  } while (0)
// once-per-CTA stamps (prologue / epilogue): slot [rank][15][ev]
#define TSX(ev)                                                                                         \
  do {                                                                                                  \
    if (blockIdx.x < 2) g_bwd2_ts[(int(rank) * 16 + 15) * 32 + (ev)] = clock64();                         \
  } while (0)
#else
#define TS(ev) do {} while (0)
#define TSX(ev) do {} while (0)
#endif

struct Bwd2Cfg {
  static constexpr int kThreads = 512;
  static constexpr int kRegsCompute = 160, kRegsDrain = 96, kRegsOther = 96;
  static constexpr int kSK = 0;                  // own K block   (128 keys x 128 d)   K-major A of S^T
  static constexpr int kSV = kSK + 32768;        // own V block                         K-major A of dP^T
  static constexpr int kSKt = kSV + 32768;       // K[256 pair keys, d half r]          MN-major B of dQ
  static constexpr int kSQb = kSKt + 32768;      // Q[query half r, all d] 2 x (64 x 128 B)   K-major B of S^T
  static constexpr int kSQg = kSQb + 16384;      // Q[128 queries, d half r]                   MN-major B of dK
  static constexpr int kSDOb = kSQg + 16384;     // dO[query half r, all d]                    K-major B of dP^T
  static constexpr int kSDOg = kSDOb + 16384;    // dO[128 queries, d half r]                  MN-major B of dV
  static constexpr int kSDS = kSDOg + 16384;     // dS^T[256 pair keys, query half r]          MN-major A of dQ
  static constexpr int kSDQ = kSDS + 32768;      // dQ staging: 2 chunks of 64 rows x 32 fp32 (two rounds per tile)
  static constexpr int kSDX = kSDQ + 16384;      // export staging: the dS^T rows owed to the peer (128 keys x 64 queries)
  static constexpr int kSStat = kSDX + 16384;    // lse_log2[2][128], dpsum[2][128]
  static constexpr int kSBar = kSStat + 2048;
  static constexpr int kNumBars = 33;   // 29..30 SCHED_FULL[2], 31..32 SCHED_EMPTY[2] (dynamic scheduler)
  static constexpr int kSmemBytes = kSBar + 8 * kNumBars + 16;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// One work item of the persistent pair kernel: a pair of adjacent 128-key blocks of one (batch, kv head).
struct Bwd2Work {
  int pair, hk, bidb;
  int q_off, sq_b, k_off, sk_b, pkey0, shift;
  int m_min, n_m, n_iter;
};
// Item index -> work item; false if the pair lies beyond its sequence (varlen) or the index is past the end.  Order:
// groups of `sched_group` (kv head, batch) combinations whose Q / dO / dq_accum stream shares L2; inside a group the pair
// index is the slow dimension, ASCENDING: with a causal mask the low pairs are the heavy ones (longest-processing-time
// first), and the deterministic mode's "lower pairs add first" chains (dq_semaphore) only ever wait for items that were
// dealt out earlier, which keeps them deadlock-free under static scheduling.
__device__ __forceinline__ bool bwd2_decode_work(const BwdKernelParams& p, int L, Bwd2Work& w) {
  const int hz_total = p.h_k * p.b;
  const int per_group = p.num_pairs * p.sched_group;
  const int grp = L / per_group;
  const int within = L - grp * per_group;
  int gsz = hz_total - grp * p.sched_group;
  gsz = gsz < p.sched_group ? gsz : p.sched_group;
  w.pair = within / gsz;
  const int hz = grp * p.sched_group + (within - w.pair * gsz);
  w.hk = hz % p.h_k;
  w.bidb = hz / p.h_k;
  w.q_off = 0; w.sq_b = p.sq; w.k_off = 0; w.sk_b = p.sk;
  if (p.cu_seqlens_q) { w.q_off = p.cu_seqlens_q[w.bidb]; w.sq_b = p.cu_seqlens_q[w.bidb + 1] - w.q_off; }
  if (p.cu_seqlens_k) { w.k_off = p.cu_seqlens_k[w.bidb]; w.sk_b = p.cu_seqlens_k[w.bidb + 1] - w.k_off; }
  w.pkey0 = w.pair * 256;
  if (w.pkey0 >= w.sk_b) return false;  // the whole pair lies beyond this sequence
  w.shift = w.sk_b - w.sq_b;
  const int m_total = (w.sq_b + 127) / 128;
  // query-block range of the PAIR = union of its key blocks' ranges; a block outside its own range sees a fully
  // masked tile (P = dS = 0), which keeps the two CTAs in lock step
  int m_min = 0, m_max = m_total;
  if (p.has_right) {
    const int lo = w.pkey0 - w.shift - p.wr;
    m_min = lo <= 0 ? 0 : lo / 128;
  }
  if (p.has_left) {
    const int last_key0 = w.pkey0 + 128 < w.sk_b ? w.pkey0 + 128 : w.pkey0;
    const int hi = last_key0 + 127 - w.shift + p.wl;
    const int nb = hi < 0 ? 0 : hi / 128 + 1;
    m_max = nb < m_max ? nb : m_max;
  }
  w.m_min = m_min;
  w.n_m = m_max > m_min ? m_max - m_min : 0;
  w.n_iter = w.n_m * p.g;
  return true;
}


// Work-item sequence of one role of the persistent pair kernel (cf. FwdSched).  Dynamic mode: cluster c starts with item c;
// the LEADER CTA's TMA thread draws the following ones from the global counter and hands each to both CTAs with a 4-byte
// st.async into a two-slot ring (SCHED_FULL counts its tx-bytes; each CTA's own TMA thread arms the barrier); every role of
// both CTAs acknowledges on the leader's SCHED_EMPTY.  Items come out of the counter in list order, so an item only ever
// waits (deterministic dQ chains) for items that were started before it.
struct Bwd2Sched {
  int G, cl, total, rounds, k;
  uint32_t full0, empty0, item_addr;   // this CTA's shared::cta addresses
  uint32_t empty0_leader;              // shared::cluster address of the leader's SCHED_EMPTY[0]
  bool dyn;
  __device__ __forceinline__ int static_item(int r) const { return r * G + ((r & 1) ? (G - 1 - cl) : cl); }
  __device__ __forceinline__ int first() {
    k = 0;
    return cl < total ? cl : -1;
  }
  __device__ __forceinline__ int next_static() {
    if (k >= rounds) return -1;
    const int L = static_item(k);
    return L < total ? L : -1;
  }
  // every role except the leader's TMA thread.  arm: the caller is the (peer) CTA's TMA thread, which arms SCHED_FULL
  __device__ __forceinline__ int next(bool whole_warp, int lane, bool arm = false) {
    ++k;
    if (!dyn) return next_static();
    const int j = k - 1;
    const uint32_t s = uint32_t(j & 1);
    if (arm) mbar_arrive_expect_tx(full0 + 8u * s, 4);
    mbar_wait(full0 + 8u * s, (j >> 1) & 1, 700);
    int L;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(L) : "r"(item_addr + 4u * s) : "memory");
    if (whole_warp) __syncwarp();
    if (!whole_warp || lane == 0) mbar_arrive_cluster(empty0_leader + 8u * s);
    return L < total ? L : -1;
  }
  __device__ __forceinline__ int draw(int* counter) const { return G + atomicAdd(counter, 1); }
  // the leader's TMA thread
  __device__ __forceinline__ int publish(int L) {
    const int j = k++;
    const uint32_t s = uint32_t(j & 1);
    mbar_wait(empty0 + 8u * s, ((j >> 1) & 1) ^ 1, 701);
    mbar_arrive_expect_tx(full0 + 8u * s, 4);
#pragma unroll
    for (uint32_t rk = 0; rk < 2; ++rk)
      st_async_b32(mapa_rank(item_addr + 4u * s, rk), uint32_t(L), mapa_rank(full0 + 8u * s, rk));
    return L < total ? L : -1;
  }
};

// PERSISTENT: the grid is one CTA pair (cluster) per two SMs; every pair loops over work items (bwd2_decode_work), drawn
// dynamically from a global counter (Bwd2Sched) or, without one, dealt out in snake order (round r: cluster c takes item
// r*G + c, or r*G + G-1-c on odd rounds).  Barrier phases, the Q / dO
// stages and TMEM stay live across items; what an item needs from its predecessor is tracked by two extra barriers
// (ACC_FREE: dV / dK accumulators read out by both CTAs' epilogues; STAGE_FREE: the epilogue's dV / dK staging -- the dS^T and
// K^T buffers, idle at the end of an item -- has been read by its TMA stores in both CTAs).  K / V are free as soon as the
// item's last MMA is done (DKV_F), so the next item's K / V / Q / dO loads and its first S^T, dP^T run under the current
// item's dK / dV epilogue; only its K^T panel (first needed by dQ, late in the first tile) waits for STAGE_FREE.
template <typename T, bool kSoftcap, bool kDropout>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Bwd2Cfg::kThreads, 1)
fa_bwd2_kernel(const __grid_constant__ BwdKernelParams p) {
  using Cfg = Bwd2Cfg;
  constexpr int kD = 128;
  const uint32_t rank = cluster_ctarank();
  const int G = int(gridDim.x >> 1), cl = int(blockIdx.x >> 1);
  const int total_items = p.total_items;
  const int rounds = (total_items + G - 1) / G;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  uint8_t* smem_gen = smem_raw;
  const uint32_t sK = smem_base + Cfg::kSK, sV = smem_base + Cfg::kSV, sKt = smem_base + Cfg::kSKt;
  const uint32_t sQb = smem_base + Cfg::kSQb, sQg = smem_base + Cfg::kSQg;
  const uint32_t sDOb = smem_base + Cfg::kSDOb, sDOg = smem_base + Cfg::kSDOg;
  const uint32_t sDS = smem_base + Cfg::kSDS, sDQ = smem_base + Cfg::kSDQ, sDX = smem_base + Cfg::kSDX;
  const uint32_t sStat = smem_base + Cfg::kSStat, sBar = smem_base + Cfg::kSBar;
  // F_*: "operand landed" barriers, live in the LEADER (rank 0) and collect the TMA bytes of both CTAs
  // E_*: "operand consumed" barriers, one per CTA, signalled by multicast tcgen05.commit
  enum { F_KV = 0, F_QB = 1, F_DOB = 2, F_DOG = 3, F_QG = 4, E_QB = 5, E_DOB = 6, E_DOG = 7, E_QG = 8, F_ST0 = 9,
         F_ST1 = 10, S_F = 11, DP_F = 12, P_F = 13, DS_F = 14, DQ_F = 15, DQ_E = 16, DKV_F = 17, DS_IN = 18, DS_PEER = 19, T_DV = 20, T_S = 21, T_DK = 22, T_DQ = 23, T_DP = 24, S_RD = 25, ACC_FREE = 26, STAGE_FREE = 27, F_KT = 28, SCHED_FULL0 = 29, SCHED_EMPTY0 = 31 };
  auto bar = [&](int i) { return sBar + 8u * i; };
  const uint32_t sTmemPtr = sBar + 8u * Cfg::kNumBars;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::kSBar + 8 * Cfg::kNumBars);

  Bwd2Sched sched;
  sched.G = G; sched.cl = cl; sched.total = total_items; sched.rounds = rounds; sched.k = 0;
  sched.full0 = bar(SCHED_FULL0); sched.empty0 = bar(SCHED_EMPTY0); sched.item_addr = sTmemPtr + 8u;
  sched.empty0_leader = mapa_rank(bar(SCHED_EMPTY0), 0);
  sched.dyn = p.sched_counter != nullptr;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TSX(0);
#if FA_BWD2_TS
  unsigned int span_slot = 0;
  if (threadIdx.x == 0) {
    span_slot = atomicAdd(&g_bwd2_span_n, 1u);
    if (span_slot < 8192) {
      unsigned int smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      g_bwd2_span[3 * span_slot] = smid;
      g_bwd2_span[3 * span_slot + 1] = global_timer_ns();
    }
  }
#endif
  if (warp == 13 && lane == 0) {
    tma_prefetch_desc(&p.tm_q); tma_prefetch_desc(&p.tm_do); tma_prefetch_desc(&p.tm_k); tma_prefetch_desc(&p.tm_v);
    tma_prefetch_desc(&p.tm_q64); tma_prefetch_desc(&p.tm_do64); tma_prefetch_desc(&p.tm_dq64);
  }
  if (warp == 12) {
    if (lane == 0) {
      mbar_init(bar(F_KV), 1);
      mbar_init(bar(F_QB), 1); mbar_init(bar(F_DOB), 1); mbar_init(bar(F_DOG), 1); mbar_init(bar(F_QG), 1);
      mbar_init(bar(E_QB), 1); mbar_init(bar(E_DOB), 1); mbar_init(bar(E_DOG), 1); mbar_init(bar(E_QG), 1);
      mbar_init(bar(F_ST0), 1); mbar_init(bar(F_ST1), 1);
      mbar_init(bar(S_F), 1); mbar_init(bar(DP_F), 1);
      mbar_init(bar(P_F), 16); mbar_init(bar(DS_F), 16);   // 8 compute warps of each CTA
      mbar_init(bar(DQ_F), 1); mbar_init(bar(DQ_E), 8);    // 4 drain warps of each CTA
      mbar_init(bar(DKV_F), 1);
      mbar_init(bar(DS_IN), FA_BWD2_LOCAL_DS ? 8 : 4);     // the 4 local writer warps of sDS (+ 16 KB of st.async tx-bytes from the peer)
#if FA_BWD2_TS
      mbar_init(bar(T_DV), 1); mbar_init(bar(T_S), 1); mbar_init(bar(T_DK), 1); mbar_init(bar(T_DQ), 1); mbar_init(bar(T_DP), 1);
#endif
      mbar_init(bar(S_RD), 16);     // leader only: all compute warps of the pair have read S^T out of TMEM
      mbar_init(bar(DS_PEER), 1);   // leader only: the peer's sDS is complete (relayed by the peer's warp 12)
      mbar_init(bar(ACC_FREE), 16); // leader only: all compute warps of the pair have read dV / dK out of TMEM (item epilogue)
      mbar_init(bar(STAGE_FREE), 2); // per CTA: the dV / dK staging (the dS^T and K^T buffers, idle at the end of an item) of BOTH
                                     // CTAs has been read by the epilogue's TMA stores (one arrive from each CTA's epilogue thread)
      mbar_init(bar(F_KT), 1);       // leader: the K^T panels of both CTAs have landed (loaded after STAGE_FREE, needed by dQ only)
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar(SCHED_FULL0 + i), 1);    // armed by this CTA's TMA thread (+ 4 tx-bytes: the item index)
        mbar_init(bar(SCHED_EMPTY0 + i), 27);  // leader only: MMA thread + 4 drain + 8 compute warps; peer: + TMA and relay threads
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_pair(sTmemPtr, 512);
    tmem_relinquish_pair();
  }
  if (threadIdx.x == 0) TSX(1);
  tc_fence_before();
  cluster_sync_all();   // barriers of both CTAs initialised, both TMEM allocations done
  tc_fence_after();
  if (threadIdx.x == 0) TSX(2);
  const uint32_t tmem_base = *tmem_ptr_gen;
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 384;

  if (warp >= 12) {
    reg_dealloc<Cfg::kRegsOther>();
    // ================================================================ TMA producer (one per CTA)
    if (warp == 13) {
      if (elect_one()) {
        const uint32_t fKV = mapa_rank(bar(F_KV), 0), fQB = mapa_rank(bar(F_QB), 0), fDOB = mapa_rank(bar(F_DOB), 0);
        const uint32_t fDOG = mapa_rank(bar(F_DOG), 0), fQG = mapa_rank(bar(F_QG), 0);
        const int dcol = 64 * int(rank);
        int git = 0;      // tiles processed so far by this pair, over all its items: phase of the per-tile barriers
        int kitem = 0;    // items with work processed so far: phase of F_KV / F_KT / DKV_F / STAGE_FREE
        int nxt = -1;
        bool drawn = false;
        auto advance = [&]() -> int {
          if (!sched.dyn || rank != 0) return sched.next(false, 0, /*arm=*/true);
          if (!drawn) nxt = sched.draw(p.sched_counter);
          drawn = false;
          return sched.publish(nxt);
        };
        for (int L = sched.first(); L >= 0; L = advance()) {
          Bwd2Work w;
          if (!bwd2_decode_work(p, L, w) || w.n_iter == 0) continue;
          const int bk = p.cu_seqlens_k ? 0 : w.bidb, bq = p.cu_seqlens_q ? 0 : w.bidb;
          const int key0 = w.pkey0 + 128 * int(rank);
          // this pair's K / V tiles are free once the previous item's last MMA is done (multicast commit)
          if (kitem > 0) mbar_wait(bar(DKV_F), (kitem - 1) & 1, 404);
          if (rank == 0) mbar_arrive_expect_tx(bar(F_KV), 2 * 2 * 32768);
#pragma unroll
          for (int pn = 0; pn < 2; ++pn) {
            tma_load_4d_pair(sK + pn * 16384, &p.tm_k, fKV, pn * 64, w.k_off + key0, w.hk, bk);
            tma_load_4d_pair(sV + pn * 16384, &p.tm_v, fKV, pn * 64, w.k_off + key0, w.hk, bk);
          }
          for (int it = 0; it < w.n_iter; ++it, ++git) {
            if (sched.dyn && rank == 0 && !drawn && it + 2 >= w.n_iter) {   // the next item is needed two tiles from now
              nxt = sched.draw(p.sched_counter);
              drawn = true;
            }
            // (first tile: the K^T panel is loaded behind the Q rows below.  It doubles as the dK staging of the previous
            // item's epilogue, so it waits for those TMA stores in both CTAs; dQ, its first consumer, is a tile's work away)
            const int hq = w.hk * p.g + it / w.n_m;
            const int m = w.m_min + it % w.n_m;
            const int st = git & 1;
            const int qrow = w.q_off + m * 128;
            const int64_t srow = ws_rows(p.cu_seqlens_q, w.bidb, p.h, hq, p.sq_r, p.ws_head_pitch).stat_base + m * 128;
            const uint32_t free_par = (git & 1) ^ 1;
            mbar_wait(bar(E_QB), free_par, 400);     // S^T of the previous tile has read sQb
            if (rank == 0) mbar_arrive_expect_tx(bar(F_QB), 2 * 16384);
            tma_load_4d_pair(sQb, &p.tm_q64, fQB, 0, qrow + 64 * int(rank), hq, bq);
            tma_load_4d_pair(sQb + 8192, &p.tm_q64, fQB, 64, qrow + 64 * int(rank), hq, bq);
            if (it == 0) {
              if (kitem > 0) mbar_wait(bar(STAGE_FREE), (kitem - 1) & 1, 405);
              if (rank == 0) mbar_arrive_expect_tx(bar(F_KT), 2 * 32768);
              const uint32_t fKT = mapa_rank(bar(F_KT), 0);
#pragma unroll
              for (int pn = 0; pn < 2; ++pn)
                tma_load_4d_pair(sKt + pn * 16384, &p.tm_k, fKT, dcol, w.k_off + w.pkey0 + 128 * pn, w.hk, bk);
              ++kitem;
            }
            // row statistics of this query block (all 128 rows) for this CTA's compute warps
            mbar_arrive_expect_tx(bar(F_ST0 + st), 1024);
            bulk_load_1d(sStat + st * 512, p.lse_log2 + srow, 512, bar(F_ST0 + st));
            bulk_load_1d(sStat + 1024 + st * 512, p.dpsum + srow, 512, bar(F_ST0 + st));
            mbar_wait(bar(E_DOB), free_par, 401);
            if (rank == 0) mbar_arrive_expect_tx(bar(F_DOB), 2 * 16384);
            tma_load_4d_pair(sDOb, &p.tm_do64, fDOB, 0, qrow + 64 * int(rank), hq, bq);
            tma_load_4d_pair(sDOb + 8192, &p.tm_do64, fDOB, 64, qrow + 64 * int(rank), hq, bq);
            mbar_wait(bar(E_DOG), free_par, 402);
            if (rank == 0) mbar_arrive_expect_tx(bar(F_DOG), 2 * 16384);
            tma_load_4d_pair(sDOg, &p.tm_do, fDOG, dcol, qrow, hq, bq);
            mbar_wait(bar(E_QG), free_par, 403);
            if (rank == 0) mbar_arrive_expect_tx(bar(F_QG), 2 * 16384);
            tma_load_4d_pair(sQg, &p.tm_q, fQG, dcol, qrow, hq, bq);
          }
        }
      }
      __syncwarp();
    }
#if FA_BWD2_TS
    else if (warp == 14) {
      if (rank == 0 && lane == 0 && blockIdx.x < 2) {   // completion times of the MMA groups (leader clock), first item only
        Bwd2Work w;
        const bool ts_on = true;
        if (bwd2_decode_work(p, sched.static_item(0), w)) {
          for (int it = 0; it < w.n_iter; ++it) {
            mbar_wait(bar(T_DV), it & 1, 800); TS(21);
            if (it + 1 < w.n_iter) { mbar_wait(bar(T_S), it & 1, 801); TS(22); }
            mbar_wait(bar(T_DK), it & 1, 802); TS(23);
            mbar_wait(bar(T_DQ), it & 1, 803); TS(24);
            if (it + 1 < w.n_iter) { mbar_wait(bar(T_DP), it & 1, 804); TS(25); }
          }
        }
      }
      __syncwarp();
    }
#endif
    // ================================================================ MMA issuer: one thread of the LEADER CTA
    else if (warp == 12) {
      if (rank == 0 && elect_one()) {
        constexpr uint32_t fmt = Elt<T>::kFmt;
        constexpr uint32_t id_s = make_idesc(fmt, fmt, 0, 0, 256, 128);    // S^T, dP^T : K-major A and B
        constexpr uint32_t id_dkv = make_idesc(fmt, fmt, 0, 1, 256, kD);   // dV, dK   : A in TMEM, B MN-major
        constexpr uint32_t id_dq = make_idesc(fmt, fmt, 1, 1, 128, kD);    // dQ       : A and B MN-major
        constexpr uint64_t dk_major = smem_desc_sw128(16, 1024);
        constexpr uint64_t dmn_major = smem_desc_sw128(16384, 1024);
        constexpr uint32_t hi_k = uint32_t(dk_major >> 32), hi_mn = uint32_t(dmn_major >> 32);
        auto lo_k = [](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | uint32_t(dk_major); };
        auto lo_mn = [](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | uint32_t(dmn_major); };
        auto mk = [](uint32_t hi, uint32_t lo) { return (uint64_t(hi) << 32) | lo; };
        // S^T / dP^T: A = 128 x 128 K-major tile (two 16 KB d panels), B = 64 x 128 K-major (two 8 KB d panels)
        auto mma_kmajor = [&](uint32_t d_t, uint32_t a0, uint32_t b0) {
          const uint32_t la = lo_k(a0), lb = lo_k(b0);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t offa = ((kk / 4) * 16384 + (kk % 4) * 32) >> 4;
            const uint32_t offb = ((kk / 4) * 8192 + (kk % 4) * 32) >> 4;
            mma_ss_pair(d_t, mk(hi_k, la + offa), mk(hi_k, lb + offb), id_s, kk > 0);
          }
        };
        auto mma_ts_mn = [&](uint32_t d_t, uint32_t a_t, uint32_t b0, bool acc) {
          const uint32_t lb = lo_mn(b0);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            mma_ts_pair(d_t, a_t + kk * 8, mk(hi_mn, lb + kk * (2048 >> 4)), id_dkv, (acc || kk > 0) ? 1u : 0u);
        };
        const uint32_t lo_ds = lo_mn(sDS), lo_kt = lo_mn(sKt);
        // TMEM regions and their tenants over one tile n (accumulators dV [256,384) and dK [384,512) aside):
        //   R1 = [0,128)   : S^T(n) fp32 -> once read: P^T(n) packed in [0,64), dQ(n-1 .. see below) in [64,128)
        //   R2 = [128,256) : dP^T(n) fp32 -> once read: dS^T(n) packed in [128,192)
        // Issue order dV(n), S^T(n+1), dK(n), dP^T(n+1), dQ(n): every MMA only waits for its own inputs; S^T(n+1) runs
        // while the compute warps still form dS^T(n), dP^T(n+1) right behind dK(n) while they form P^T(n+1), and dQ(n)
        // -- which has to wait for the exported half of dS^T -- lands in R1[64,128) after S^T(n+1) has been read, off the
        // critical path (it only has to be drained before S^T(n+2) is issued).  Tiles are numbered `git` across ALL items of
        // this pair (barrier phases continue); an item's first S^T / dP^T are issued while its predecessor's epilogue runs.
        int git = 0, kitem = 0;
        int r = -1;
        for (int L = sched.first(); L >= 0; L = sched.next(false, 0)) {
          ++r;
          Bwd2Work w;
          if (!bwd2_decode_work(p, L, w) || w.n_iter == 0) continue;
          const bool ts_on = FA_BWD2_TS && r == 0 && blockIdx.x < 2;
          (void)ts_on;
          mbar_wait(bar(F_KV), kitem & 1, 500);
          if (r == 0) TSX(3);
          // first tile of the item: S^T and dP^T (runs under the previous item's epilogue)
          if (git > 0) mbar_wait(bar(DQ_E), (git & 1) ^ 1, 523);   // dQ(git-1) drained out of R1[64,128) in both CTAs
          mbar_wait(bar(F_QB), git & 1, 501);
          if (r == 0) TSX(4);
          tc_fence_after();
          mma_kmajor(tS, sK, sQb);
          mma_commit_pair(bar(S_F));
          mma_commit_pair(bar(E_QB));
          mbar_wait(bar(F_DOB), git & 1, 502);
          tc_fence_after();
          mma_kmajor(tDP, sV, sDOb);
          mma_commit_pair(bar(DP_F));
          mma_commit_pair(bar(E_DOB));
          // dV / dK accumulators: the previous item's epilogue (both CTAs) must have read them out of TMEM
          if (kitem > 0) mbar_wait(bar(ACC_FREE), (kitem - 1) & 1, 524);
          for (int it = 0; it < w.n_iter; ++it, ++git) {
            const uint32_t ph = git & 1;
            const bool more = it + 1 < w.n_iter;
            mbar_wait(bar(P_F), ph, 510);
            TS(0);
            mbar_wait(bar(S_RD), ph, 521);   // (already complete: every warp reads S^T before it delivers P^T; keeps the phases in step)
            mbar_wait(bar(F_DOG), ph, 515);
            tc_fence_after();
            mma_ts_mn(tDV, tS, sDOg, it > 0);
            mma_commit_pair(bar(E_DOG));
#if FA_BWD2_TS
            mma_commit_pair(bar(T_DV));
#endif
            TS(1);
            if (more) {
              if (git > 0) mbar_wait(bar(DQ_E), ph ^ 1, 513);   // dQ(git-1) drained out of R1[64,128) in both CTAs
              TS(6);
              mbar_wait(bar(F_QB), (git + 1) & 1, 512);
              TS(2);
              tc_fence_after();
              mma_kmajor(tS, sK, sQb);
              mma_commit_pair(bar(S_F));
              mma_commit_pair(bar(E_QB));
#if FA_BWD2_TS
              mma_commit_pair(bar(T_S));
#endif
            }
            mbar_wait(bar(DS_F), ph, 511);
            TS(3);
            mbar_wait(bar(F_QG), ph, 516);
            tc_fence_after();
            mma_ts_mn(tDK, tDP, sQg, it > 0);
            mma_commit_pair(bar(E_QG));
#if FA_BWD2_TS
            mma_commit_pair(bar(T_DK));
#endif
            if (more) {
              mbar_wait(bar(F_DOB), (git + 1) & 1, 514);
              tc_fence_after();
              mma_kmajor(tDP, sV, sDOb);
              mma_commit_pair(bar(DP_F));
              mma_commit_pair(bar(E_DOB));
#if FA_BWD2_TS
              mma_commit_pair(bar(T_DP));
#endif
              TS(7);
            }
            if (it == 0) mbar_wait(bar(F_KT), kitem & 1, 525);   // this item's K^T panels (both CTAs)
            mbar_wait(bar(DS_IN), ph, 517);     // dS^T rows of both key blocks have landed in my shared memory ...
            TS(4);
            mbar_wait(bar(DS_PEER), ph, 518);   // ... and in the peer's
            TS(5);
            if (more) mbar_wait(bar(S_RD), (git + 1) & 1, 519);   // S^T(git+1) is out of R1: its upper half is free for dQ
            else if (git > 0) mbar_wait(bar(DQ_E), ph ^ 1, 522);  // last tile of the item: no S^T issue waited for the previous dQ's drain
            fence_proxy_async();
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 16; ++kk)
              mma_ss_pair(tS + 64, mk(hi_mn, lo_ds + kk * (2048 >> 4)), mk(hi_mn, lo_kt + kk * (2048 >> 4)), id_dq, kk > 0);
            mma_commit_pair(bar(DQ_F));
#if FA_BWD2_TS
            mma_commit_pair(bar(T_DQ));
#endif
          }
          mma_commit_pair(bar(DKV_F));
          ++kitem;
          if (r == 0) TSX(5);
        }
      }
      if (rank == 1 && elect_one()) {
        // relay: tell the leader when this CTA's dS^T tile (local stores + the peer's bulk copy) is complete
        const uint32_t peer_bar = mapa_rank(bar(DS_PEER), 0);
        int git = 0;
        for (int L = sched.first(); L >= 0; L = sched.next(false, 0)) {
          Bwd2Work w;
          if (!bwd2_decode_work(p, L, w) || w.n_iter == 0) continue;
          for (int it = 0; it < w.n_iter; ++it, ++git) {
            mbar_wait(bar(DS_IN), git & 1, 520);
            fence_proxy_async();
            mbar_arrive_cluster(peer_bar);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp >= 8) {
    // ================================================================ dQ drain: 64 query rows x 128 d per CTA.
    // TMEM layout of the M=128 pair accumulator: lane = q + 64 * (d / 64), column = d % 64.
    reg_dealloc<Cfg::kRegsDrain>();
    const int wq = warp & 3;
    const int q = (wq & 1) * 32 + lane;   // query row within this CTA's half
    const int dh = wq >> 1;               // d half held by this warp's lanes
    const uint32_t t_dq = tS + 64 + (uint32_t(wq * 32) << 16);
    const bool leader = (warp == 8 && lane == 0);
    const uint32_t dqe = mapa_rank(bar(DQ_E), 0);
    int git = 0;
    int r = -1;
    for (int L = sched.first(); L >= 0; L = sched.next(true, lane)) {
    ++r;
    Bwd2Work w;
    if (!bwd2_decode_work(p, L, w) || w.n_iter == 0) continue;
    const bool ts_on = FA_BWD2_TS && r == 0 && blockIdx.x < 2;
    (void)ts_on;
    const int bidb = w.bidb, pair = w.pair, shift = w.shift;
    const WsRows ws = ws_rows(p.cu_seqlens_q, bidb, p.h, 0, p.sq_r, p.ws_head_pitch);   // tensor-map coordinates only
    for (int it = 0; it < w.n_iter; ++it, ++git) {
      const int hq = w.hk * p.g + it / w.n_m;
      const int m = w.m_min + it % w.n_m;
      mbar_wait(bar(DQ_F), git & 1, 600);
      if (warp == 8 && lane == 0) TS(12);
      tc_fence_after();
      // deterministic mode: the pairs that touch (query block m, query half rank) add in ascending pair order
      int* sem = nullptr;
      int my_turn = 0;
      if (p.dq_semaphore) {
        sem = p.dq_semaphore + ((int64_t(bidb) * p.h + hq) * (p.sq_r / 128) + m) * 2 + int(rank);
        int n_lo = 0;
        if (p.has_left) {
          const int x = 128 * m - 127 + shift - p.wl;
          n_lo = x <= 0 ? 0 : (x + 127) / 128;
        }
        my_turn = pair - (n_lo >> 1);
        if (leader) {
          int cur;
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(cur) : "l"(sem) : "memory");
          } while (cur != my_turn);
        }
      }
      uint32_t v[64];
      tmem_ld_x32(t_dq, v);
      tmem_ld_x32(t_dq + 32, v + 32);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(dqe);   // TMEM [64,128) may be overwritten by the S^T after next
      if (warp == 8 && lane == 0) TS(13);
#pragma unroll
      for (int c = 0; c < 2; ++c) {   // two rounds: 32-column chunk c of each d half
        if (leader && (git > 0 || c > 0)) tma_store_wait_read<0>();   // the previous reduce-adds have read the staging
        named_bar_sync(1, 128);
        uint8_t* dst = smem_gen + Cfg::kSDQ + dh * 8192 + q * 128;
#pragma unroll
        for (int x = 0; x < 8; ++x)
          *reinterpret_cast<uint4*>(dst + ((x ^ (q & 7)) * 16)) =
              make_uint4(v[c * 32 + 4 * x], v[c * 32 + 4 * x + 1], v[c * 32 + 4 * x + 2], v[c * 32 + 4 * x + 3]);
        fence_proxy_async();
        named_bar_sync(2, 128);
        if (leader) {
          tma_reduce_add_4d(&p.tm_dq64, sDQ, c * 32, ws.tma_row0 + m * 128 + 64 * int(rank), hq, ws.tma_b);
          tma_reduce_add_4d(&p.tm_dq64, sDQ + 8192, 64 + c * 32, ws.tma_row0 + m * 128 + 64 * int(rank), hq, ws.tma_b);
          tma_store_commit();
        }
      }
      if (sem && leader) {
        tma_store_wait<0>();
        asm volatile("fence.proxy.async;" ::: "memory");
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(sem), "r"(my_turn + 1) : "memory");
      }
    }
    }
    if (leader) tma_store_wait_read<0>();   // the staging buffers must outlive the reads; completion is the grid's business
  } else {
    // ================================================================ compute warps (as in fa_bwd_kernel)
    reg_alloc<Cfg::kRegsCompute>();
    const int wq = warp & 3, hf = warp >> 2;          // key-row quadrant, query half
    const int r = wq * 32 + lane;                       // key row within this CTA's block
    const uint32_t lane_sel = uint32_t(wq * 32) << 16;
    const uint32_t t_s = tS + lane_sel + hf * 64;
    const uint32_t t_dp = tDP + lane_sel + hf * 64;
    const uint32_t t_p = tS + lane_sel + hf * 32;
    const uint32_t t_ds = tDP + lane_sel + hf * 32;   // dS^T packed: over the dP^T columns [0,64) once both halves are read
    const float scale_log2 = p.scale_log2;
    const uint32_t pf = mapa_rank(bar(P_F), 0), dsf = mapa_rank(bar(DS_F), 0), srd = mapa_rank(bar(S_RD), 0);
    // my dS^T row (key 128*rank + r of the pair, 64 queries of half hf) lives in the CTA that owns query half hf
    // exporter warps stage their 32 rows (4 KB) locally and ship them with one bulk copy into the peer's sDS
    const uint32_t ds_warp_remote = mapa_rank(sDS + uint32_t(128 * int(rank) + wq * 32) * 128, uint32_t(hf));
    const uint32_t ds_in_remote = mapa_rank(bar(DS_IN), uint32_t(hf));
    const bool ds_is_local = FA_BWD2_LOCAL_DS ? true : uint32_t(hf) == rank;
    DropoutParams drop = p.drop;
    if constexpr (kDropout) {
      if (p.rng_state) {
        const uint64_t sd = p.rng_state[0], of = p.rng_state[1];
        drop.seed_lo = uint32_t(sd); drop.seed_hi = uint32_t(sd >> 32);
        drop.off_lo = uint32_t(of); drop.off_hi = uint32_t(of >> 32);
      }
    }
    const uint32_t acc_free = mapa_rank(bar(ACC_FREE), 0);
    int git = 0, kitem = 0;
    int rr = -1;
    for (int L = sched.first(); L >= 0; L = sched.next(true, lane)) {
    ++rr;
    Bwd2Work w;
    if (!bwd2_decode_work(p, L, w)) continue;
    const bool ts_on = FA_BWD2_TS && rr == 0 && blockIdx.x < 2;
    (void)ts_on;
    const int n_iter = w.n_iter, n_m = w.n_m, m_min = w.m_min, hk = w.hk, bidb = w.bidb, shift = w.shift, sk_b = w.sk_b;
    const int k_off = w.k_off;
    const int key0 = w.pkey0 + 128 * int(rank);
    const int j = key0 + r;                             // key index in the sequence
    const int i_lo = p.has_right ? j - shift - p.wr : -(1 << 30);
    const int i_hi = p.has_left ? j - shift + p.wl : (1 << 30);
    const bool key_valid = j < sk_b;
    for (int it = 0; it < n_iter; ++it, ++git) {
      const int hq = hk * p.g + it / n_m;
      const int m = m_min + it % n_m;
      const int st = git & 1;
      const int q0 = m * 128 + hf * 64;
      float slope = 0.f;
      if (p.alibi_slopes) slope = p.alibi_slopes[bidb * p.alibi_batch_stride + hq] / p.scale_eff;
      mbar_wait(bar(F_ST0 + st), (git >> 1) & 1, 700);
      mbar_wait(bar(S_F), git & 1, 701);
      if (warp == 0 && lane == 0) TS(8);
      tc_fence_after();
      uint32_t su[64];
      tmem_ld_x32(t_s, su);
      tmem_ld_x32(t_s + 32, su + 32);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(srd);   // my S^T columns are in registers
      float* s = reinterpret_cast<float*>(su);
      const float* s_lse = reinterpret_cast<const float*>(smem_gen + Cfg::kSStat + st * 512) + hf * 64;
      const float* s_dps = reinterpret_cast<const float*>(smem_gen + Cfg::kSStat + 1024 + st * 512) + hf * 64;
      const float rel0 = float(q0 + shift - j);
      if constexpr (kSoftcap) {
#pragma unroll
        for (int c = 0; c < 64; ++c) s[c] = tanh_approx(s[c] * p.softcap_pre);
      }
      if (p.alibi_slopes) {
#pragma unroll
        for (int c = 0; c < 64; ++c) s[c] = fmaf(-slope, fabsf(rel0 + float(c)), s[c]);
      }
      bool need_mask = key0 + 127 >= sk_b;
      if (p.has_right) need_mask |= key0 + 127 > m * 128 + shift + p.wr;
      if (p.has_left) need_mask |= key0 < m * 128 + 127 + shift - p.wl;
      float pr[64];
#pragma unroll
      for (int c4 = 0; c4 < 64; c4 += 4) {
        const float4 l4 = *reinterpret_cast<const float4*>(s_lse + c4);
        pr[c4 + 0] = ex2(fmaf(s[c4 + 0], scale_log2, -l4.x));
        pr[c4 + 1] = ex2(fmaf(s[c4 + 1], scale_log2, -l4.y));
        if (FA_BWD2_POLY && ((c4 >> 2) % FA_BWD2_POLY) == 0) {  // FMA-pipe exp2 for a share of the elements (MUFU relief)
          ex2_poly2(fmaf(s[c4 + 2], scale_log2, -l4.z), fmaf(s[c4 + 3], scale_log2, -l4.w), pr[c4 + 2], pr[c4 + 3]);
        } else {
          pr[c4 + 2] = ex2(fmaf(s[c4 + 2], scale_log2, -l4.z));
          pr[c4 + 3] = ex2(fmaf(s[c4 + 3], scale_log2, -l4.w));
        }
      }
      if (need_mask) {
        const int lo = i_lo - q0, hi = i_hi - q0;
#pragma unroll
        for (int c = 0; c < 64; ++c) pr[c] = (!key_valid || c < lo || c > hi) ? 0.f : pr[c];
      }
      uint32_t keep_lo = 0xffffffffu, keep_hi = 0xffffffffu;
      if constexpr (kDropout) {
        keep_lo = keep_hi = 0u;
        const uint32_t bh = uint32_t(bidb * p.h + hq);
#pragma unroll
        for (int c16 = 0; c16 < 64; c16 += 16) {   // the reference's stream: 8 draws per 16 queries (fa_dropout.cuh)
          const uint32_t bits = dropout_keep_col16(drop, uint32_t(q0 + c16) >> 4, uint32_t(j), bh);
          if (c16 < 32) keep_lo |= bits << c16;
          else keep_hi |= bits << (c16 - 32);
        }
      }
      auto kept = [&](int c) { return ((c < 32 ? keep_lo >> c : keep_hi >> (c - 32)) & 1u) != 0u; };
      {
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 64; c += 2) {
          if constexpr (kDropout) pk[c / 2] = Elt<T>::pack(kept(c) ? pr[c] : 0.f, kept(c + 1) ? pr[c + 1] : 0.f);
          else pk[c / 2] = Elt<T>::pack(pr[c], pr[c + 1]);
        }
        tmem_st_x32(t_p, pk);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(pf);
        if (warp == 0 && lane == 0) TS(9);
      }
      mbar_wait(bar(DP_F), git & 1, 702);
      if (warp == 0 && lane == 0) TS(10);
      tc_fence_after();
      uint32_t dk_[32];
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        uint32_t du[32];
        tmem_ld_x32(t_dp + h2 * 32, du);
        tmem_wait_ld();
#pragma unroll
        for (int c4 = 0; c4 < 32; c4 += 4) {
          const int cc = h2 * 32 + c4;
          const float4 d4 = *reinterpret_cast<const float4*>(s_dps + cc);
          float g0 = __uint_as_float(du[c4 + 0]), g1 = __uint_as_float(du[c4 + 1]);
          float g2 = __uint_as_float(du[c4 + 2]), g3 = __uint_as_float(du[c4 + 3]);
          if constexpr (kDropout) {
            g0 = kept(cc + 0) ? g0 * drop.rp : 0.f;
            g1 = kept(cc + 1) ? g1 * drop.rp : 0.f;
            g2 = kept(cc + 2) ? g2 * drop.rp : 0.f;
            g3 = kept(cc + 3) ? g3 * drop.rp : 0.f;
          }
          float e0 = pr[cc + 0] * (g0 - d4.x);
          float e1 = pr[cc + 1] * (g1 - d4.y);
          float e2 = pr[cc + 2] * (g2 - d4.z);
          float e3 = pr[cc + 3] * (g3 - d4.w);
          if constexpr (kSoftcap) {
            float t0 = s[cc + 0], t1 = s[cc + 1], t2 = s[cc + 2], t3 = s[cc + 3];
            if (p.alibi_slopes) {
              t0 = fmaf(slope, fabsf(rel0 + float(cc + 0)), t0);
              t1 = fmaf(slope, fabsf(rel0 + float(cc + 1)), t1);
              t2 = fmaf(slope, fabsf(rel0 + float(cc + 2)), t2);
              t3 = fmaf(slope, fabsf(rel0 + float(cc + 3)), t3);
            }
            e0 *= 1.f - t0 * t0;
            e1 *= 1.f - t1 * t1;
            e2 *= 1.f - t2 * t2;
            e3 *= 1.f - t3 * t3;
          }
          dk_[cc / 2] = Elt<T>::pack(e0, e1);
          dk_[cc / 2 + 1] = Elt<T>::pack(e2, e3);
        }
      }
      if (warp == 0 && lane == 0) TS(14);
      if (warp == 4 && lane == 0) TS(18);
      named_bar_sync(3 + wq, 64);   // the other query-half warp of this lane quadrant has read its dP^T columns too
      if (warp == 0 && lane == 0) TS(16);
      if (warp == 4 && lane == 0) TS(19);
      tmem_st_x32(t_ds, dk_);
      if (warp == 0 && lane == 0) TS(26);
      // shared copy for dQ = dS K: 128-byte row (64 queries of half hf) of key 128*rank + r, 16-byte chunks XOR-swizzled,
      // written into the CTA that owns query half hf (the peer through distributed shared memory)
      // dQ of the previous tile (issued BEHIND this tile's dP^T) must have finished reading sDS in both CTAs before
      // anyone overwrites it (locally or with a bulk copy into the peer)
      if (git > 0) mbar_wait(bar(DQ_F), (git - 1) & 1, 703);
      // first tile of an item: sDS also staged the previous item's dV for its TMA stores (in both CTAs)
      if (it == 0 && kitem > 0) mbar_wait(bar(STAGE_FREE), (kitem - 1) & 1, 704);
      {
        uint8_t* dst = ds_is_local ? smem_gen + Cfg::kSDS + (128 * int(rank) + r) * 128 : smem_gen + Cfg::kSDX + r * 128;
#pragma unroll
        for (int x = 0; x < 8; ++x)
          *reinterpret_cast<uint4*>(dst + ((x ^ (r & 7)) * 16)) =
              make_uint4(dk_[4 * x], dk_[4 * x + 1], dk_[4 * x + 2], dk_[4 * x + 3]);
      }
      if (warp == 0 && lane == 0) TS(17);
      // shared-memory half of the hand-off first (the export has the longest way to go), then the TMEM half
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (ds_is_local) {   // 4 local writer warps + 16 KB of bulk copies from the peer complete DS_IN
          if (wq == 0 && !FA_BWD2_LOCAL_DS) mbar_arrive_expect_tx(bar(DS_IN), 16384);
          else mbar_arrive(bar(DS_IN));
        } else {
          bulk_copy_to_cluster(ds_warp_remote, sDX + uint32_t(wq * 32) * 128, 4096, ds_in_remote);
        }
      }
      tmem_wait_st();
      if (warp == 0 && lane == 0) TS(15);
      if (warp == 4 && lane == 0) TS(20);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(dsf);   // TMEM half (dS^T for dK): all 16 compute warps of the pair
      if (warp == 0 && lane == 0) TS(11);
    }
    // ---------------------------------------------------------------- epilogue: dK (x scale), dV (x 1/p_keep)
    if (n_iter > 0) {
      mbar_wait(bar(DKV_F), kitem & 1, 710);
      if (threadIdx.x == 0 && rr == 0) TSX(6);
      tc_fence_after();
    }
    const int bk = p.cu_seqlens_k ? 0 : bidb;
    // Fast path: the 128 x 128 dV / dK tiles go through the (now idle) dS^T / K^T shared-memory tiles and leave with TMA
    // stores (rows / columns beyond the tensor are clipped by the tensor map); K / V themselves are already being
    // reloaded for the next item.  A ragged tail block of a PACKED varlen
    // batch borders the next sequence's rows, so it takes the predicated per-thread path, as does an empty work item.
    const bool tma_epilogue = n_iter > 0 && !(p.cu_seqlens_k && key0 + 128 > sk_b);
    if (tma_epilogue) {
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const uint32_t t_src = (which == 0 ? tDV : tDK) + lane_sel + hf * 64;
        const float mul = which == 0 ? (kDropout ? drop.rp : 1.0f) : p.scale;
        uint8_t* dst = smem_gen + (which == 0 ? Cfg::kSDS : Cfg::kSKt) + hf * 16384 + r * 128;   // panel hf = d columns [64 hf, 64 hf + 64)
        uint32_t o[64];
        tmem_ld_x32(t_src, o);
        tmem_ld_x32(t_src + 32, o + 32);
        tmem_wait_ld();
        if (which == 1) {   // both accumulators are in registers: the next item's dV / dK MMAs may overwrite them
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc_free);
        }
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          uint4 v4;
          v4.x = Elt<T>::pack(__uint_as_float(o[8 * x + 0]) * mul, __uint_as_float(o[8 * x + 1]) * mul);
          v4.y = Elt<T>::pack(__uint_as_float(o[8 * x + 2]) * mul, __uint_as_float(o[8 * x + 3]) * mul);
          v4.z = Elt<T>::pack(__uint_as_float(o[8 * x + 4]) * mul, __uint_as_float(o[8 * x + 5]) * mul);
          v4.w = Elt<T>::pack(__uint_as_float(o[8 * x + 6]) * mul, __uint_as_float(o[8 * x + 7]) * mul);
          *reinterpret_cast<uint4*>(dst + ((x ^ (r & 7)) * 16)) = v4;
        }
      }
      fence_proxy_async();
      named_bar_sync(7, 256);
      if (threadIdx.x == 0) {
#pragma unroll
        for (int pn = 0; pn < 2; ++pn) {
          tma_store_4d(&p.tm_dv, sDS + pn * 16384, pn * 64, k_off + key0, hk, bk);
          tma_store_4d(&p.tm_dk, sKt + pn * 16384, pn * 64, k_off + key0, hk, bk);
        }
        tma_store_commit();
        tma_store_wait_read<0>();   // shared memory must outlive the reads; the writes complete on their own
        mbar_arrive(bar(STAGE_FREE));                                      // my staging is free again: tell myself ...
        mbar_arrive_cluster(mapa_rank(bar(STAGE_FREE), rank ^ 1u));        // ... and the peer (it bulk-copies into my sDS)
      }
    } else {
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const uint32_t t_src = (which == 0 ? tDV : tDK) + lane_sel;
      const float mul = which == 0 ? (kDropout ? drop.rp : 1.0f) : p.scale;
      T* base = reinterpret_cast<T*>(which == 0 ? p.dv : p.dk) +
                int64_t(bk) * (which == 0 ? p.dv_batch_stride : p.dk_batch_stride) +
                int64_t(k_off + j) * (which == 0 ? p.dv_row_stride : p.dk_row_stride) +
                int64_t(hk) * (which == 0 ? p.dv_head_stride : p.dk_head_stride);
      constexpr int kHalf = kD / 2;
#pragma unroll
      for (int cc = 0; cc < kHalf; cc += 32) {
        uint32_t o[32];
        if (n_iter > 0) {
          tmem_ld_x32(t_src + hf * kHalf + cc, o);
          tmem_wait_ld();
          if (which == 1 && cc + 32 >= kHalf) {   // last TMEM read of this warp: the accumulators may be overwritten
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_free);
          }
        }
        if (key_valid) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            const int col = hf * kHalf + cc + c;
            if (col < p.d) {
              uint4 v4 = make_uint4(0, 0, 0, 0);
              if (n_iter > 0) {
                v4.x = Elt<T>::pack(__uint_as_float(o[c]) * mul, __uint_as_float(o[c + 1]) * mul);
                v4.y = Elt<T>::pack(__uint_as_float(o[c + 2]) * mul, __uint_as_float(o[c + 3]) * mul);
                v4.z = Elt<T>::pack(__uint_as_float(o[c + 4]) * mul, __uint_as_float(o[c + 5]) * mul);
                v4.w = Elt<T>::pack(__uint_as_float(o[c + 6]) * mul, __uint_as_float(o[c + 7]) * mul);
              }
              *reinterpret_cast<uint4*>(base + col) = v4;
            }
          }
        }
      }
    }
    if (n_iter > 0 && threadIdx.x == 0) {   // no staging used on this path: free at once (both CTAs always arrive, see above)
      mbar_arrive(bar(STAGE_FREE));
      mbar_arrive_cluster(mapa_rank(bar(STAGE_FREE), rank ^ 1u));
    }
    }
    if (n_iter > 0) ++kitem;
    }   // work items
  }

  if (threadIdx.x == 0) TSX(7);
  tc_fence_before();
  cluster_sync_all();   // nobody leaves while the peer may still signal its barriers or write its shared memory
  if (threadIdx.x == 0) TSX(8);
#if FA_BWD2_TS
  if (threadIdx.x == 0 && span_slot < 8192) g_bwd2_span[3 * span_slot + 2] = global_timer_ns();
#endif
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace fa
