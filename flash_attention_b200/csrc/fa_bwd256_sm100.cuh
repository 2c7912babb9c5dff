// fa_bwd256_sm100.cuh — backward main loop for head dims 129..256 (reference: flash_bwd_kernel.h:80-795 instantiated
// for hdim 192 / 256, flash_bwd_launch_template.h:262-322).
//
// At d = 256 the two resident accumulators dK, dV alone need all 512 TMEM columns, so the head dimension of the
// OUTPUTS is split over two independent CTAs (c = blockIdx.x & 1 owns columns [128c, 128c+128) of dQ, dK, dV) while
// both contract S^T and dP^T over the full head dim.  The redundant S^T / dP^T work (7 GEMM units instead of 5) buys
// the simplest possible structure: no exchange between the halves, and the d = 128 kernel's role split is kept.
//
// One CTA = one 128-key block x one output half; loop over 64-query tiles (x the g query heads of a GQA group):
//   S^T   = K Q^T               SS  M=128 N=64  K=256   -> TMEM [0,64)
//   dP^T  = V dO^T              SS  M=128 N=64  K=256   -> TMEM [64,128)
//   P^T   = exp2(S^T c - lse)   compute warps, 16-bit   -> TMEM [0,16) / [32,48)   (each warp inside its OWN S^T columns)
//   dS^T  = P^T (dP^T - D)      compute warps           -> TMEM [16,32) / [48,64) and shared memory (swizzled)
//   dV_c += P^T dO[:, c]        TS  M=128 N=128 K=64    -> TMEM [256,384)
//   dK_c += dS^T Q[:, c]        TS  M=128 N=128 K=64    -> TMEM [384,512)
//   dQ_c^T = K[:, c]^T dS^T     SS  M=128 N=64  K=128   -> TMEM [128,192)   (lane = head-dim column, column = query)
//   dQ_c^T -> registers -> red.global.add.f32 into dq_accum (a warp adds 32 consecutive floats of one query row)
//
// Shared memory 225 KB: K, V (4 panels of 64 columns x 128 rows each), Q (4 panels x 64 rows; the two panels of this
// CTA's output half double-buffered), dO (single stage), dS^T.
// Warps as in fa_bwd_kernel: 0-7 compute (key-row quadrant w%4, query half w/4), 8-11 dQ drain, 12 MMA issuer, 13 TMA.
#pragma once
#include "fa_bwd_sm100.cuh"

namespace fa {

struct Bwd256Cfg {
  static constexpr int kThreads = 512;
  static constexpr int kRegsCompute = 160, kRegsDrain = 96, kRegsOther = 96;
  static constexpr int kPanelsD = 4;                  // 64-column panels of the (padded) head dim
  static constexpr int kKPanel = 128 * 128;           // K / V panel: 128 key rows x 128 B
  static constexpr int kQPanel = 64 * 128;            // Q / dO panel: 64 query rows x 128 B
  static constexpr int kSK = 0;
  static constexpr int kSV = kSK + kPanelsD * kKPanel;
  // Q: the two panels of the OTHER output half are dead once S^T has been computed, the two panels of this CTA's half
  // are also the B operand of dK -> own-half panels double-buffered, so that the next tile's Q can be loaded as soon
  // as S^T of this tile is done (a single-stage Q exposed the whole load latency every tile: 528 TFLOP/s)
  static constexpr int kSQo = kSV + kPanelsD * kKPanel;    // other half: 2 panels, 1 stage
  static constexpr int kSQc = kSQo + 2 * kQPanel;          // own half:   2 panels x 2 stages
  static constexpr int kSDO = kSQc + 4 * kQPanel;
  static constexpr int kSDS = kSDO + kPanelsD * kQPanel;   // 128 keys x 64 queries, 16-bit
  static constexpr int kSStat = kSDS + kKPanel;            // 2 stages x (lse_log2[64], dpsum[64])
  static constexpr int kSBar = kSStat + 1024;
  static constexpr int kNumBars = 16;
  static constexpr int kSmemBytes = kSBar + 8 * kNumBars + 16;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB shared memory of an SM");
};

template <typename T, bool kSoftcap, bool kDropout>
__global__ void __launch_bounds__(Bwd256Cfg::kThreads, 1)
fa_bwd256_kernel(const __grid_constant__ BwdKernelParams p) {
  using Cfg = Bwd256Cfg;
  const int n_block = blockIdx.x >> 1, chalf = blockIdx.x & 1;
  const int hk = blockIdx.y, bidb = blockIdx.z;
  int q_off = 0, sq_b = p.sq, k_off = 0, sk_b = p.sk;
  if (p.cu_seqlens_q) { q_off = p.cu_seqlens_q[bidb]; sq_b = p.cu_seqlens_q[bidb + 1] - q_off; }
  if (p.cu_seqlens_k) { k_off = p.cu_seqlens_k[bidb]; sk_b = p.cu_seqlens_k[bidb + 1] - k_off; }
  const int key0 = n_block * 128;
  if (key0 >= sk_b) return;
  const int shift = sk_b - sq_b;
  const int m_total = (sq_b + 63) / 64;
  int m_min = 0, m_max = m_total;
  if (p.has_right) {  // query i sees key j iff j <= i + shift + wr
    const int lo = key0 - shift - p.wr;
    m_min = lo <= 0 ? 0 : lo / 64;
  }
  if (p.has_left) {   // ... and j >= i + shift - wl
    const int hi = key0 + 127 - shift + p.wl;
    const int nb = hi < 0 ? 0 : hi / 64 + 1;
    m_max = nb < m_max ? nb : m_max;
  }
  const int n_m = m_max > m_min ? m_max - m_min : 0;
  const int n_iter = n_m * p.g;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  uint8_t* smem_gen = smem_raw;
  const uint32_t sK = smem_base + Cfg::kSK, sV = smem_base + Cfg::kSV;
  const uint32_t sQo = smem_base + Cfg::kSQo, sQc = smem_base + Cfg::kSQc;
  const uint32_t sDO = smem_base + Cfg::kSDO, sDS = smem_base + Cfg::kSDS;
  const uint32_t sStat = smem_base + Cfg::kSStat, sBar = smem_base + Cfg::kSBar;
  enum { B_KV = 0, B_QF0, B_QF1, B_QEO, B_QEC0, B_QEC1, B_DOF, B_DOE, B_SF, B_DPF, B_PF, B_DSF, B_DQF, B_DQE, B_DKV };
  auto bar = [&](int i) { return sBar + 8u * i; };
  const uint32_t sTmemPtr = sBar + 8u * Cfg::kNumBars;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::kSBar + 8 * Cfg::kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 13 && lane == 0) {
    tma_prefetch_desc(&p.tm_q64); tma_prefetch_desc(&p.tm_do64); tma_prefetch_desc(&p.tm_k); tma_prefetch_desc(&p.tm_v);
  }
  if (warp == 12) {
    if (lane == 0) {
      mbar_init(bar(B_KV), 1);
      mbar_init(bar(B_QF0), 1); mbar_init(bar(B_QF1), 1);
      mbar_init(bar(B_QEO), 1); mbar_init(bar(B_QEC0), 1); mbar_init(bar(B_QEC1), 1);
      mbar_init(bar(B_DOF), 1); mbar_init(bar(B_DOE), 1);
      mbar_init(bar(B_SF), 1); mbar_init(bar(B_DPF), 1);
      mbar_init(bar(B_PF), 8); mbar_init(bar(B_DSF), 8);
      mbar_init(bar(B_DQF), 1); mbar_init(bar(B_DQE), 4);
      mbar_init(bar(B_DKV), 1);
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
  const uint32_t tS = tmem_base, tDP = tmem_base + 64, tDQ = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 384;

  if (warp >= 12) {
    reg_dealloc<Cfg::kRegsOther>();
    // ================================================================ TMA producer
    if (warp == 13) {
      if (elect_one()) {
        const int bk = p.cu_seqlens_k ? 0 : bidb, bq = p.cu_seqlens_q ? 0 : bidb;
        mbar_arrive_expect_tx(bar(B_KV), 2 * Cfg::kPanelsD * Cfg::kKPanel);
#pragma unroll
        for (int pn = 0; pn < Cfg::kPanelsD; ++pn) {
          tma_load_4d(sK + pn * Cfg::kKPanel, &p.tm_k, bar(B_KV), pn * 64, k_off + key0, hk, bk);
          tma_load_4d(sV + pn * Cfg::kKPanel, &p.tm_v, bar(B_KV), pn * 64, k_off + key0, hk, bk);
        }
        for (int it = 0; it < n_iter; ++it) {
          const int hq = hk * p.g + it / n_m;
          const int m = m_min + it % n_m;
          const int64_t srow = ws_rows(p.cu_seqlens_q, bidb, p.h, hq, p.sq_r, p.ws_head_pitch).stat_base + m * 64;
          const int st = it & 1;
          mbar_wait(bar(B_QEO), (it & 1) ^ 1, 400);                  // S^T of tile it-1 done
          mbar_wait(bar(B_QEC0 + st), ((it >> 1) & 1) ^ 1, 401);     // dK of tile it-2 done
          mbar_arrive_expect_tx(bar(B_QF0 + st), Cfg::kPanelsD * Cfg::kQPanel + 512);
#pragma unroll
          for (int pn = 0; pn < Cfg::kPanelsD; ++pn) {
            const uint32_t dst = ((pn >> 1) == chalf ? sQc + st * 2 * Cfg::kQPanel : sQo) + (pn & 1) * Cfg::kQPanel;
            tma_load_4d(dst, &p.tm_q64, bar(B_QF0 + st), pn * 64, q_off + m * 64, hq, bq);
          }
          bulk_load_1d(sStat + st * 512, p.lse_log2 + srow, 256, bar(B_QF0 + st));
          bulk_load_1d(sStat + st * 512 + 256, p.dpsum + srow, 256, bar(B_QF0 + st));
          mbar_wait(bar(B_DOE), (it & 1) ^ 1, 410);
          mbar_arrive_expect_tx(bar(B_DOF), Cfg::kPanelsD * Cfg::kQPanel);
#pragma unroll
          for (int pn = 0; pn < Cfg::kPanelsD; ++pn)
            tma_load_4d(sDO + pn * Cfg::kQPanel, &p.tm_do64, bar(B_DOF), pn * 64, q_off + m * 64, hq, bq);
        }
      }
      __syncwarp();
    }
    // ================================================================ MMA issuer (one elected thread)
    else if (warp == 12) {
      if (n_iter > 0 && elect_one()) {
        constexpr uint32_t fmt = Elt<T>::kFmt;
        constexpr uint32_t id_s = make_idesc(fmt, fmt, 0, 0, 128, 64);     // S^T, dP^T : K-major A and B
        constexpr uint32_t id_dkv128 = make_idesc(fmt, fmt, 0, 1, 128, 128);  // dV, dK : A in TMEM, B MN-major
        constexpr uint32_t id_dkv64 = make_idesc(fmt, fmt, 0, 1, 128, 64);
        // head dims <= 192: the fourth 64-column panel is all zero-fill -> skip its 4 k-steps of S^T / dP^T, and the
        // upper output half only has 64 live columns
        const int nk = p.d <= 192 ? 12 : 16;
        const uint32_t id_dkv = (chalf == 1 && p.d <= 192) ? id_dkv64 : id_dkv128;
        constexpr uint32_t id_dq = make_idesc(fmt, fmt, 1, 1, 128, 64);    // dQ^T      : A and B MN-major
        constexpr uint64_t d_kmaj = smem_desc_sw128(16, 1024);
        constexpr uint64_t d_mn_k = smem_desc_sw128(Cfg::kKPanel, 1024);   // MN-major view of 128-row panels (K, dS^T)
        constexpr uint64_t d_mn_q = smem_desc_sw128(Cfg::kQPanel, 1024);   // MN-major view of 64-row panels (Q, dO)
        auto lo = [](uint64_t base, uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | uint32_t(base); };
        auto mk = [](uint64_t base, uint32_t lo32) { return (base & 0xFFFFFFFF00000000ull) | lo32; };
        // D = A B^T over the full head dim: A = 128-row panels (K or V); B = 64-row panels: panels of this CTA's output
        // half start at b_own, the other two at b_oth (dO: one contiguous tile, Q: split buffers)
        auto mma_full_d = [&](uint32_t d_t, uint32_t a0, uint32_t b_own, uint32_t b_oth) {
          const uint32_t la = lo(d_kmaj, a0), lown = lo(d_kmaj, b_own), loth = lo(d_kmaj, b_oth);
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) {
            const int pn = kk / 4;
            const uint32_t offa = (pn * Cfg::kKPanel + (kk % 4) * 32) >> 4;
            const uint32_t offb = ((pn & 1) * Cfg::kQPanel + (kk % 4) * 32) >> 4;
            if (kk < nk)
              mma_ss<false>(d_t, mk(d_kmaj, la + offa), mk(d_kmaj, ((pn >> 1) == chalf ? lown : loth) + offb), id_s, kk > 0);
          }
        };
        const uint32_t sDOown = sDO + 2 * chalf * Cfg::kQPanel, sDOoth = sDO + 2 * (chalf ^ 1) * Cfg::kQPanel;
        // D (+)= A[tmem, 64 queries packed] * B[64 query rows x 128 head-dim columns of this half, MN-major]
        auto mma_out = [&](uint32_t d_t, uint32_t a_t, uint32_t b_own, bool acc) {
          const uint32_t lb = lo(d_mn_q, b_own);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_ts<false>(d_t, a_t + (kk >> 1) * 32 + (kk & 1) * 8, mk(d_mn_q, lb + kk * (2048 >> 4)), id_dkv,
                          (acc || kk > 0) ? 1u : 0u);
        };
        const uint32_t lo_kt = lo(d_mn_k, sK + 2 * chalf * Cfg::kKPanel), lo_ds = lo(d_mn_k, sDS);
        mbar_wait(bar(B_KV), 0, 500);
        mbar_wait(bar(B_QF0), 0, 501);
        tc_fence_after();
        mma_full_d(tS, sK, sQc, sQo);
        mma_commit(bar(B_SF));
        mma_commit(bar(B_QEO));
        mbar_wait(bar(B_DOF), 0, 502);
        tc_fence_after();
        mma_full_d(tDP, sV, sDOown, sDOoth);
        mma_commit(bar(B_DPF));
        for (int it = 0; it < n_iter; ++it) {
          const uint32_t ph = it & 1;
          mbar_wait(bar(B_PF), ph, 510);
          tc_fence_after();
          const int st = it & 1;
          mma_out(tDV, tS, sDOown, it > 0);
          mma_commit(bar(B_DOE));
          mbar_wait(bar(B_DSF), ph, 511);
          tc_fence_after();
          mma_out(tDK, tS + 16, sQc + st * 2 * Cfg::kQPanel, it > 0);
          mma_commit(bar(B_QEC0 + st));
          if (it > 0) {
            mbar_wait(bar(B_DQE), ph ^ 1, 512);   // dQ^T of the previous tile has been read out of TMEM [128,192)
            tc_fence_after();
          }
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            mma_ss<false>(tDQ, mk(d_mn_k, lo_kt + kk * (2048 >> 4)), mk(d_mn_k, lo_ds + kk * (2048 >> 4)), id_dq, kk > 0);
          mma_commit(bar(B_DQF));
          if (it + 1 < n_iter) {
            const int st1 = st ^ 1;
            mbar_wait(bar(B_QF0 + st1), ((it + 1) >> 1) & 1, 513);
            tc_fence_after();
            mma_full_d(tS, sK, sQc + st1 * 2 * Cfg::kQPanel, sQo);
            mma_commit(bar(B_SF));
            mma_commit(bar(B_QEO));
            mbar_wait(bar(B_DOF), ph ^ 1, 514);
            tc_fence_after();
            mma_full_d(tDP, sV, sDOown, sDOoth);
            mma_commit(bar(B_DPF));
          }
        }
        mma_commit(bar(B_DKV));
      }
      __syncwarp();
    }
  } else if (warp >= 8) {
    // ================================================================ dQ drain: TMEM -> registers -> fp32 red.add
    reg_dealloc<Cfg::kRegsDrain>();
    const int wq = warp & 3;
    const int dcol = chalf * 128 + wq * 32 + lane;            // head-dim column of this thread
    const bool col_live = chalf * 128 + wq * 32 < p.d;        // warp-uniform: columns >= d are padding
    const uint32_t t_dq = tDQ + (uint32_t(wq * 32) << 16);
    const bool leader = (warp == 8 && lane == 0);
    for (int it = 0; it < n_iter; ++it) {
      const int hq = hk * p.g + it / n_m;
      const int m = m_min + it % n_m;
      mbar_wait(bar(B_DQF), it & 1, 600);
      tc_fence_after();
      // Deterministic mode: the key blocks that touch query tile m add their partials in ascending key-block order,
      // one chain per (batch, head, output half, 64-query tile)  (same scheme as fa_bwd_kernel)
      int* sem = nullptr;
      int my_turn = 0;
      if (p.dq_semaphore) {
        sem = p.dq_semaphore + ((int64_t(bidb) * p.h + hq) * 2 + chalf) * (p.sq_r / 64) + m;
        int n_lo = 0;
        if (p.has_left) {
          const int x = 64 * m - 127 + shift - p.wl;
          n_lo = x <= 0 ? 0 : (x + 127) / 128;
        }
        my_turn = n_block - n_lo;
        if (leader) {
          int cur;
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(cur) : "l"(sem) : "memory");
          } while (cur != my_turn);
        }
        named_bar_sync(1, 128);
      }
      float* dst = p.dq_accum + (ws_rows(p.cu_seqlens_q, bidb, p.h, hq, p.sq_r, p.ws_head_pitch).stat_base + m * 64) * 256 + dcol;
      const int rows = sq_b - m * 64;   // >= 1; rows beyond the sequence carry zeros and are skipped
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld_x32(t_dq + half * 32, v);
        tmem_wait_ld();
        if (half == 1) {   // dQ^T is out of TMEM: the MMA warp may overwrite it with the next tile's
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(B_DQE));
        }
        if (col_live) {
#pragma unroll
          for (int q = 0; q < 32; ++q)
            if (half * 32 + q < rows) atomicAdd(dst + (half * 32 + q) * 256, __uint_as_float(v[q]));
        }
      }
      if (sem) {
        __threadfence();
        named_bar_sync(2, 128);
        if (leader) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(sem), "r"(my_turn + 1) : "memory");
      }
    }
  } else {
    // ================================================================ compute warps
    reg_alloc<Cfg::kRegsCompute>();
    const int wq = warp & 3, hf = warp >> 2;            // key-row quadrant, query half (32 queries)
    const int r = wq * 32 + lane;
    const int j = key0 + r;
    const uint32_t lane_sel = uint32_t(wq * 32) << 16;
    const uint32_t t_s = tS + lane_sel + hf * 32;       // my 32 S^T columns
    const uint32_t t_dp = tDP + lane_sel + hf * 32;
    const uint32_t t_p = t_s;                           // P^T packed: 16 columns, inside my own S^T columns
    const uint32_t t_ds = t_s + 16;                     // dS^T packed: the other 16
    const float scale_log2 = p.scale_log2;
    const int i_lo = p.has_right ? j - shift - p.wr : -(1 << 30);
    const int i_hi = p.has_left ? j - shift + p.wl : (1 << 30);
    const bool key_valid = j < sk_b;
    DropoutParams drop = p.drop;
    if constexpr (kDropout) {
      if (p.rng_state) {
        const uint64_t sd = p.rng_state[0], of = p.rng_state[1];
        drop.seed_lo = uint32_t(sd); drop.seed_hi = uint32_t(sd >> 32);
        drop.off_lo = uint32_t(of); drop.off_hi = uint32_t(of >> 32);
      }
    }
    for (int it = 0; it < n_iter; ++it) {
      const int hq = hk * p.g + it / n_m;
      const int m = m_min + it % n_m;
      const int q0 = m * 64 + hf * 32;
      float slope = 0.f;
      if (p.alibi_slopes) slope = p.alibi_slopes[bidb * p.alibi_batch_stride + hq] / p.scale_eff;
      const int st = it & 1;
      mbar_wait(bar(B_QF0 + st), (it >> 1) & 1, 700);   // lse_log2 / dpsum of this stage are in shared memory
      mbar_wait(bar(B_SF), it & 1, 701);
      tc_fence_after();
      uint32_t su[32];
      tmem_ld_x32(t_s, su);
      tmem_wait_ld();
      float* s = reinterpret_cast<float*>(su);
      // this stage is refilled by the load of tile it+2, which waits for dK of this tile, which is issued only after
      // every compute warp has arrived on B_DSF: the statistics stay valid for the whole iteration
      const float* s_lse = reinterpret_cast<const float*>(smem_gen + Cfg::kSStat + st * 512) + hf * 32;
      const float* s_dps = reinterpret_cast<const float*>(smem_gen + Cfg::kSStat + st * 512 + 256) + hf * 32;
      const float rel0 = float(q0 + shift - j);
      if constexpr (kSoftcap) {
#pragma unroll
        for (int c = 0; c < 32; ++c) s[c] = tanh_approx(s[c] * p.softcap_pre);
      }
      if (p.alibi_slopes) {
#pragma unroll
        for (int c = 0; c < 32; ++c) s[c] = fmaf(-slope, fabsf(rel0 + float(c)), s[c]);
      }
      bool need_mask = key0 + 127 >= sk_b;
      if (p.has_right) need_mask |= key0 + 127 > m * 64 + shift + p.wr;
      if (p.has_left) need_mask |= key0 < m * 64 + 63 + shift - p.wl;
      float pr[32];
#pragma unroll
      for (int c4 = 0; c4 < 32; c4 += 4) {
        const float4 l4 = *reinterpret_cast<const float4*>(s_lse + c4);
        pr[c4 + 0] = ex2(fmaf(s[c4 + 0], scale_log2, -l4.x));
        pr[c4 + 1] = ex2(fmaf(s[c4 + 1], scale_log2, -l4.y));
        pr[c4 + 2] = ex2(fmaf(s[c4 + 2], scale_log2, -l4.z));
        pr[c4 + 3] = ex2(fmaf(s[c4 + 3], scale_log2, -l4.w));
      }
      if (need_mask) {
        const int lo = i_lo - q0, hi = i_hi - q0;
#pragma unroll
        for (int c = 0; c < 32; ++c) pr[c] = (!key_valid || c < lo || c > hi) ? 0.f : pr[c];
      }
      uint32_t keep = 0xffffffffu;
      if constexpr (kDropout) {
        keep = 0u;
        const uint32_t bh = uint32_t(bidb * p.h + hq);
#pragma unroll
        for (int c16 = 0; c16 < 32; c16 += 16)   // the reference's stream: 8 draws per 16 queries (fa_dropout.cuh)
          keep |= dropout_keep_col16(drop, uint32_t(q0 + c16) >> 4, uint32_t(j), bh) << c16;
      }
      auto kept = [&](int c) { return ((keep >> c) & 1u) != 0u; };
      {
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          if constexpr (kDropout) pk[c / 2] = Elt<T>::pack(kept(c) ? pr[c] : 0.f, kept(c + 1) ? pr[c + 1] : 0.f);
          else pk[c / 2] = Elt<T>::pack(pr[c], pr[c + 1]);
        }
        tmem_st_x16(t_p, pk);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_PF));
      }
      // ---- dS^T = P^T * (dP^T - D[q])  (* (1 - tanh^2) with softcap)
      mbar_wait(bar(B_DPF), it & 1, 702);
      tc_fence_after();
      uint32_t du[32], dk_[16];
      tmem_ld_x32(t_dp, du);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        float g0 = __uint_as_float(du[c]), g1 = __uint_as_float(du[c + 1]);
        if constexpr (kDropout) {
          g0 = kept(c) ? g0 * drop.rp : 0.f;
          g1 = kept(c + 1) ? g1 * drop.rp : 0.f;
        }
        float e0 = pr[c] * (g0 - s_dps[c]);
        float e1 = pr[c + 1] * (g1 - s_dps[c + 1]);
        if constexpr (kSoftcap) {  // d tanh(x) = 1 - tanh^2 (utils.h:403-409); undo the ALiBi bias first
          float t0 = s[c], t1 = s[c + 1];
          if (p.alibi_slopes) {
            t0 = fmaf(slope, fabsf(rel0 + float(c)), t0);
            t1 = fmaf(slope, fabsf(rel0 + float(c + 1)), t1);
          }
          e0 *= 1.f - t0 * t0;
          e1 *= 1.f - t1 * t1;
        }
        dk_[c / 2] = Elt<T>::pack(e0, e1);
      }
      tmem_st_x16(t_ds, dk_);
      {
        // shared copy for dQ^T = K^T dS^T: one 128-byte row per key (64 queries), 16-byte chunks XOR-swizzled
        uint8_t* dst = smem_gen + Cfg::kSDS + r * 128;
#pragma unroll
        for (int x = 0; x < 4; ++x)
          *reinterpret_cast<uint4*>(dst + (((hf * 4 + x) ^ (r & 7)) * 16)) =
              make_uint4(dk_[4 * x], dk_[4 * x + 1], dk_[4 * x + 2], dk_[4 * x + 3]);
      }
      tmem_wait_st();
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_DSF));
    }
    // ---------------------------------------------------------------- epilogue: this half's 128 columns of dK (x scale), dV
    if (n_iter > 0) {
      mbar_wait(bar(B_DKV), 0, 710);
      tc_fence_after();
    }
    const int bk = p.cu_seqlens_k ? 0 : bidb;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const uint32_t t_src = (which == 0 ? tDV : tDK) + lane_sel;
      const float mul = which == 0 ? (kDropout ? drop.rp : 1.0f) : p.scale;
      T* base = reinterpret_cast<T*>(which == 0 ? p.dv : p.dk) +
                int64_t(bk) * (which == 0 ? p.dv_batch_stride : p.dk_batch_stride) +
                int64_t(k_off + j) * (which == 0 ? p.dv_row_stride : p.dk_row_stride) +
                int64_t(hk) * (which == 0 ? p.dv_head_stride : p.dk_head_stride);
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        uint32_t o[32];
        if (n_iter > 0) {
          tmem_ld_x32(t_src + hf * 64 + cc, o);
          tmem_wait_ld();
        }
        if (key_valid) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            const int col = chalf * 128 + hf * 64 + cc + c;
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace fa
