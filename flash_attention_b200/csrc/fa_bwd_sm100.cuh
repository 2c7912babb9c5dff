// fa_bwd_sm100.cuh — backward attention (dQ, dK, dV) for sm_100a (B200).
//
// Three kernels (the reference's split, flash_bwd_launch_template.h:72-126):
//   fa_bwd_preprocess  D = rowsum(dO * O), lse_log2 = LSE * log2(e)          (flash_bwd_preprocess_kernel.h:57-140)
//   fa_bwd_kernel      per 128-key block: loop over the 128-query blocks (and over the g query heads of a
//                      GQA group) recomputing P and accumulating dK, dV in TMEM; dQ partials are
//                      TMA-reduce-added into an fp32 accumulator                 (flash_bwd_kernel.h:80-795)
//   fa_bwd_convert_dq  dq = T(dq_accum * softmax_scale)                          (flash_bwd_preprocess_kernel.h:185-268)
//
// Main loop, everything transposed so that the resident K/V block is the MMA M operand (TMEM lane = key):
//   S^T  = K Q^T            SS   -> TMEM [0,128)
//   dP^T = V dO^T           SS   -> TMEM [128,256)
//   P^T  = exp2(S^T*c - lse_log2[q])      compute warps, fp32 -> 16-bit into TMEM [0,32) + [64,96)
//   dS^T = P^T * (dP^T - D[q])            compute warps -> TMEM [32,64) + [96,128) and shared memory (swizzled)
//                                         (each warp packs into its OWN 64 S^T columns, so it can never overwrite
//                                         columns that the warp of the other query half has not read yet)
//   dV  += P^T dO           TS   -> TMEM [256,384)      (A = P^T from TMEM, B = dO as MN-major)
//   dK  += dS^T Q           TS   -> TMEM [384,512)
//   dQ   = dS K             SS   -> TMEM [128,256)      (A = dS^T tile as MN-major, B = K as MN-major)
//   dQ -> registers -> swizzled smem -> cp.reduce.async.bulk.tensor (fp32 add) into dq_accum
//
// Warps (512 threads, 4 warpgroups): 0-7 compute (warp w: key rows 32*(w%4).., query half w/4),
// 8-11 dQ drain, 12 MMA issuer, 13 TMA producer.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "fa_dropout.cuh"
#include "fa_fwd_sm100.cuh"
#include "fa_ptx.cuh"

namespace fa {

struct BwdKernelParams {
  CUtensorMap tm_q, tm_do;   // 4-D (d, rows, h, b)    box (64, 128, 1, 1)  16-bit
  CUtensorMap tm_k, tm_v;    // 4-D (d, rows, h_k, b)
  CUtensorMap tm_dq;         // 4-D (kD, sq_r, h, b) fp32, box (32, 128, 1, 1), 128B swizzle, reduce-add target
  CUtensorMap tm_q64, tm_do64;  // CTA-pair kernel (fa_bwd2_sm100.cuh): the same tensors with box (64, 64, 1, 1)
  CUtensorMap tm_dq64;          // CTA-pair kernel: dq_accum with box (32, 64, 1, 1)
  CUtensorMap tm_dk, tm_dv;     // CTA-pair kernel epilogue: dK / dV outputs, box (64, 128, 1, 1)
  const float* lse_log2;     // (b, h, sq_r)
  int* dq_semaphore;         // deterministic mode: (b, h, sq_r/128) counters, zero-initialised; else nullptr
  const float* dpsum;        // (b, h, sq_r)
  void* dk;
  void* dv;
  int64_t dk_batch_stride, dk_row_stride, dk_head_stride;
  int64_t dv_batch_stride, dv_row_stride, dv_head_stride;
  const int* cu_seqlens_q;
  const int* cu_seqlens_k;
  const float* alibi_slopes;
  int64_t alibi_batch_stride;
  int b, h, h_k, sq, sk, d, g;
  int sq_r;                  // round128(max seqlen_q): row pitch of lse_log2 / dpsum / dq_accum (dense layout)
  int64_t ws_head_pitch;     // packed varlen layout: rows per head of lse_log2 / dpsum / dq_accum (see ws_rows below)
  int has_right, has_left, wl, wr;
  float scale;               // softmax_scale: dK is multiplied by it in the epilogue (dQ in the convert kernel)
  float scale_eff;           // scale of the logits inside exp: softcap if softcap > 0 else softmax_scale
  float scale_log2;          // scale_eff * log2(e)
  float softcap_pre;
  int sched_lpt;              // (single-CTA kernel) key block index is the slowest grid dimension: heavy (masked) work first
  int num_pairs, sched_group, total_items;   // persistent CTA-pair kernel (fa_bwd2_sm100.cuh): work-item enumeration
  int* sched_counter;                        // its dynamic scheduler's work counter (zeroed by fa_bwd_preprocess), or nullptr
  DropoutParams drop;         // kDropout kernels only: the forward's rng_state and p_dropout
  const uint64_t* rng_state;  // device copy of {seed, offset}; overrides drop.seed/off when non-null
  float* dq_accum;            // fa_bwd256_kernel only (adds its dQ partials with red.global instead of TMA): (b, h, sq_r, 256)
};

// Row addressing of the backward scratch (lse_log2, dpsum: one float per row; dq_accum: kD floats per row).
//   dense  (b, h, sq_r)            row(b, h, r) = (b * H + h) * sq_r + r
//   varlen (h, R)                  row(b, h, r) = h * R + (cu_seqlens_q[b] & ~3) + 132 * b + r
// The packed form follows the reference's (total_q + 128 * batch) padding (flash_api.cpp:1142-1159) so that scratch
// grows with total_q, not batch * max_seqlen_q; the start is rounded down to 4 rows (16-byte aligned bulk copies of
// the statistics) and the per-batch pad is 132 so that a sequence's round128 rows never reach the next sequence's.
struct WsRows {
  int64_t stat_base;   // index of row 0 of (batch, head) in the statistics arrays / dq_accum rows
  int tma_row0;        // row coordinate of row 0 in the dq_accum tensor map
  int tma_b;           // batch coordinate in the dq_accum tensor map
};
__device__ __forceinline__ WsRows ws_rows(const int* cu_q, int bidb, int h, int head, int sq_r, int64_t head_pitch) {
  WsRows w;
  if (cu_q) {
    w.tma_row0 = (cu_q[bidb] & ~3) + 132 * bidb;
    w.tma_b = 0;
    w.stat_base = int64_t(head) * head_pitch + w.tma_row0;
  } else {
    w.tma_row0 = 0;
    w.tma_b = bidb;
    w.stat_base = (int64_t(bidb) * h + head) * sq_r;
  }
  return w;
}

template <int kD>
struct BwdCfg {
  static constexpr int kThreads = 512;
  // setmaxnreg split over the 4 warpgroups (pool = 512 threads x 128 regs): compute, compute, drain, MMA/TMA
  static constexpr int kRegsCompute = 160, kRegsDrain = 96, kRegsOther = 96;
  static_assert(2 * kRegsCompute + kRegsDrain + kRegsOther <= 4 * 128, "setmaxnreg split exceeds the CTA register pool");
  static constexpr int kPanels = kD / 64;
  static constexpr int kPanelBytes = 128 * 128;
  static constexpr int kTileBytes = kPanels * kPanelBytes;   // one 128-row 16-bit tile
  static constexpr int kSK = 0;
  static constexpr int kSV = kSK + kTileBytes;
  static constexpr int kSQ = kSV + kTileBytes;               // 2 stages
  static constexpr int kSDO = kSQ + 2 * kTileBytes;
  static constexpr int kSDS = kSDO + kTileBytes;             // 128 keys x 128 queries, 16-bit, two 64-query panels
  static constexpr int kSDQ = kSDS + 2 * kPanelBytes;        // 2 staging buffers of 128 rows x 32 fp32
  static constexpr int kSStat = kSDQ + 2 * kPanelBytes;      // lse_log2[2][128], dpsum[2][128]
  static constexpr int kSBar = kSStat + 4 * 512;
  static constexpr int kNumBars = 16;
  static constexpr int kSmemBytes = kSBar + 8 * kNumBars + 16;  // dynamic smem is declared 1024-aligned
};

// ---------------------------------------------------------------------------------------------- preprocess
// one warp per 4 query rows (all 8 loads of a warp in flight): D = sum_d dO*O ; lse_log2 = lse*log2e (+inf on padding
// rows so that P == 0 there); also clears the rows of the fp32 dQ accumulator (`dq_accum_rows` = kD floats per row, or
// nullptr when the host clears it), which saves the separate 256 MB memset pass at the benchmark shape
template <typename T>
__global__ void __launch_bounds__(256) fa_bwd_preprocess(const T* __restrict__ o, const T* __restrict__ dout,
                                                         const float* __restrict__ lse, float* __restrict__ dpsum,
                                                         float* __restrict__ lse_log2, float* __restrict__ softmax_d,
                                                         int64_t o_bs, int64_t o_rs, int64_t o_hs, int64_t do_bs,
                                                         int64_t do_rs, int64_t do_hs, const int* __restrict__ cu_q,
                                                         int b, int h, int sq, int sq_r, int d, int total_q,
                                                         float* __restrict__ dq_accum_rows, int kD, int64_t ws_head_pitch,
                                                         int* __restrict__ sched_counter) {
  constexpr int kRows = 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (sched_counter && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) *sched_counter = 0;
  const int row0 = (blockIdx.x * 8 + warp) * kRows;  // in [0, sq_r)
  const int head = blockIdx.y, bidb = blockIdx.z;
  if (row0 >= sq_r) return;
  int q_off = 0, sq_b = sq;
  if (cu_q) { q_off = cu_q[bidb]; sq_b = cu_q[bidb + 1] - q_off; }
  const int rows_b = cu_q ? (sq_b + 127) / 128 * 128 : sq_r;   // scratch rows of this (batch, head)
  if (row0 >= rows_b) return;
  const int64_t sidx0 = ws_rows(cu_q, bidb, h, head, sq_r, ws_head_pitch).stat_base + row0;
  const T* po = o + (cu_q ? 0 : bidb * o_bs) + int64_t(head) * o_hs;
  const T* pd = dout + (cu_q ? 0 : bidb * do_bs) + int64_t(head) * do_hs;
  float acc[kRows], l2[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) { acc[r] = 0.f; l2[r] = INFINITY; }
  for (int c = lane * 4; c < d; c += 128) {
    uint2 a[kRows], g[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      a[r] = g[r] = make_uint2(0u, 0u);
      if (row0 + r < sq_b) {
        a[r] = *reinterpret_cast<const uint2*>(po + int64_t(q_off + row0 + r) * o_rs + c);
        g[r] = *reinterpret_cast<const uint2*>(pd + int64_t(q_off + row0 + r) * do_rs + c);
      }
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const T* av = reinterpret_cast<const T*>(&a[r]);
      const T* gv = reinterpret_cast<const T*>(&g[r]);
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[r] += float(av[e]) * float(gv[e]);
    }
  }
  if (dq_accum_rows) {
#pragma unroll
    for (int r = 0; r < kRows; ++r)
      if (row0 + r < rows_b && lane * 4 < kD)
        *reinterpret_cast<float4*>(dq_accum_rows + (sidx0 + r) * kD + lane * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], s);
  }
  if (lane < kRows) {
    const int row = row0 + lane;
    float my_acc = 0.f;
#pragma unroll
    for (int r = 0; r < kRows; ++r) my_acc = lane == r ? acc[r] : my_acc;
    if (row < rows_b) {
      float my_l2 = INFINITY;
      if (row < sq_b) {
        const float l = cu_q ? lse[int64_t(head) * total_q + q_off + row] : lse[(int64_t(bidb) * h + head) * sq + row];
        my_l2 = l * 1.4426950408889634f;  // +inf stays +inf (fully masked row), -inf cannot occur for non-split forward
      }
      const int64_t sidx = sidx0 + lane;
      dpsum[sidx] = my_acc;
      lse_log2[sidx] = my_l2;
      // API-visible softmax_d in the reference layouts: dense (b,h,sq_r) | varlen (h, total_q + 128*b)
      if (softmax_d) {
        if (cu_q) softmax_d[int64_t(head) * (total_q + 128 * b) + q_off + 128 * bidb + row] = my_acc;
        else softmax_d[sidx] = my_acc;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- dq convert
template <typename T, int kD>
__global__ void __launch_bounds__(256) fa_bwd_convert_dq(const float* __restrict__ dq_accum, T* __restrict__ dq,
                                                         int64_t dq_bs, int64_t dq_rs, int64_t dq_hs,
                                                         const int* __restrict__ cu_q, int h, int sq, int sq_r, int d,
                                                         float scale, int64_t ws_head_pitch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  const int head = blockIdx.y, bidb = blockIdx.z;
  int q_off = 0, sq_b = sq;
  if (cu_q) { q_off = cu_q[bidb]; sq_b = cu_q[bidb + 1] - q_off; }
  if (row >= sq_b) return;
  T* dst = dq + (cu_q ? 0 : bidb * dq_bs) + int64_t(q_off + row) * dq_rs + int64_t(head) * dq_hs;
  const float* src = dq_accum + (ws_rows(cu_q, bidb, h, head, sq_r, ws_head_pitch).stat_base + row) * kD;
  for (int c = lane * 4; c < d; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(src + c);
    uint2 o;
    o.x = Elt<T>::pack(v.x * scale, v.y * scale);
    o.y = Elt<T>::pack(v.z * scale, v.w * scale);
    *reinterpret_cast<uint2*>(dst + c) = o;
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}

// ---------------------------------------------------------------------------------------------- main kernel
template <typename T, int kD, bool kSoftcap, bool kDropout = false>
__global__ void __launch_bounds__(BwdCfg<kD>::kThreads, 1)
fa_bwd_kernel(const __grid_constant__ BwdKernelParams p) {
  using Cfg = BwdCfg<kD>;
  constexpr float kNegInf = -INFINITY;
  const int n_block = p.sched_lpt ? blockIdx.z : blockIdx.x;
  const int hk = p.sched_lpt ? blockIdx.x : blockIdx.y, bidb = p.sched_lpt ? blockIdx.y : blockIdx.z;
  int q_off = 0, sq_b = p.sq, k_off = 0, sk_b = p.sk;
  if (p.cu_seqlens_q) { q_off = p.cu_seqlens_q[bidb]; sq_b = p.cu_seqlens_q[bidb + 1] - q_off; }
  if (p.cu_seqlens_k) { k_off = p.cu_seqlens_k[bidb]; sk_b = p.cu_seqlens_k[bidb + 1] - k_off; }
  const int key0 = n_block * 128;
  if (key0 >= sk_b) return;
  const int shift = sk_b - sq_b;
  const int m_total = (sq_b + 127) / 128;
  int m_min = 0, m_max = m_total;
  if (p.has_right) {  // query i sees key j iff j <= i + shift + wr  ->  i >= j - shift - wr
    int lo = key0 - shift - p.wr;
    m_min = lo <= 0 ? 0 : lo / 128;
  }
  if (p.has_left) {   // j >= i + shift - wl  ->  i <= j - shift + wl
    int hi = key0 + 127 - shift + p.wl;  // last query that can see the last key of the block
    int nb = hi < 0 ? 0 : hi / 128 + 1;
    m_max = nb < m_max ? nb : m_max;
  }
  const int n_m = m_max > m_min ? m_max - m_min : 0;
  const int n_iter = n_m * p.g;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();  // the 128B-swizzle atoms need 1024-byte aligned tiles
  uint8_t* smem_gen = smem_raw;
  const uint32_t sK = smem_base + Cfg::kSK, sV = smem_base + Cfg::kSV, sQ = smem_base + Cfg::kSQ;
  const uint32_t sDO = smem_base + Cfg::kSDO, sDS = smem_base + Cfg::kSDS, sDQ = smem_base + Cfg::kSDQ;
  const uint32_t sStat = smem_base + Cfg::kSStat, sBar = smem_base + Cfg::kSBar;
  enum { B_KV = 0, B_QF0 = 1, B_QF1 = 2, B_QE0 = 3, B_QE1 = 4, B_DOF = 5, B_DOE = 6, B_SF = 7, B_DPF = 8, B_PF = 9,
         B_DSF = 10, B_DQF = 11, B_DQE = 12, B_DKV = 13 };
  auto bar = [&](int i) { return sBar + 8u * i; };
  const uint32_t sTmemPtr = sBar + 8u * Cfg::kNumBars;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::kSBar + 8 * Cfg::kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 13 && lane == 0) {
    tma_prefetch_desc(&p.tm_q); tma_prefetch_desc(&p.tm_do); tma_prefetch_desc(&p.tm_k);
    tma_prefetch_desc(&p.tm_v); tma_prefetch_desc(&p.tm_dq);
  }
  if (warp == 12) {
    if (lane == 0) {
      mbar_init(bar(B_KV), 1);
      mbar_init(bar(B_QF0), 1); mbar_init(bar(B_QF1), 1);
      mbar_init(bar(B_QE0), 1); mbar_init(bar(B_QE1), 1);
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
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 384;

  if (warp >= 12) {
    reg_dealloc<Cfg::kRegsOther>();
    // ================================================================ TMA producer
    if (warp == 13) {
      if (elect_one()) {
        const int bk = p.cu_seqlens_k ? 0 : bidb, bq = p.cu_seqlens_q ? 0 : bidb;
        mbar_arrive_expect_tx(bar(B_KV), 2 * Cfg::kTileBytes);
#pragma unroll
        for (int pn = 0; pn < Cfg::kPanels; ++pn) {
          tma_load_4d(sK + pn * Cfg::kPanelBytes, &p.tm_k, bar(B_KV), pn * 64, k_off + key0, hk, bk);
          tma_load_4d(sV + pn * Cfg::kPanelBytes, &p.tm_v, bar(B_KV), pn * 64, k_off + key0, hk, bk);
        }
        for (int it = 0; it < n_iter; ++it) {
          const int hq = hk * p.g + it / n_m;
          const int m = m_min + it % n_m;
          const int st = it & 1;
          const int64_t srow = ws_rows(p.cu_seqlens_q, bidb, p.h, hq, p.sq_r, p.ws_head_pitch).stat_base + m * 128;
          mbar_wait(bar(B_QE0 + st), ((it >> 1) & 1) ^ 1, 400 + st);
          mbar_arrive_expect_tx(bar(B_QF0 + st), Cfg::kTileBytes + 1024);
#pragma unroll
          for (int pn = 0; pn < Cfg::kPanels; ++pn)
            tma_load_4d(sQ + st * Cfg::kTileBytes + pn * Cfg::kPanelBytes, &p.tm_q, bar(B_QF0 + st), pn * 64,
                        q_off + m * 128, hq, bq);
          bulk_load_1d(sStat + st * 512, p.lse_log2 + srow, 512, bar(B_QF0 + st));
          bulk_load_1d(sStat + 1024 + st * 512, p.dpsum + srow, 512, bar(B_QF0 + st));
          mbar_wait(bar(B_DOE), (it & 1) ^ 1, 410);
          mbar_arrive_expect_tx(bar(B_DOF), Cfg::kTileBytes);
#pragma unroll
          for (int pn = 0; pn < Cfg::kPanels; ++pn)
            tma_load_4d(sDO + pn * Cfg::kPanelBytes, &p.tm_do, bar(B_DOF), pn * 64, q_off + m * 128, hq, bq);
        }
      }
      __syncwarp();
    }
    // ================================================================ MMA issuer (one elected thread)
    else if (warp == 12) {
      if (n_iter > 0 && elect_one()) {
        constexpr uint32_t fmt = Elt<T>::kFmt;
        constexpr uint32_t id_s = make_idesc(fmt, fmt, 0, 0, 128, 128);    // S^T, dP^T : K-major A and B
        constexpr uint32_t id_dkv = make_idesc(fmt, fmt, 0, 1, 128, kD);   // dV, dK   : A in TMEM, B MN-major
        constexpr uint32_t id_dq = make_idesc(fmt, fmt, 1, 1, 128, kD);    // dQ       : A and B MN-major
        constexpr uint64_t dk_major = smem_desc_sw128(16, 1024);
        constexpr uint64_t dmn_major = smem_desc_sw128(Cfg::kPanelBytes, 1024);
        constexpr uint32_t hi_k = uint32_t(dk_major >> 32), hi_mn = uint32_t(dmn_major >> 32);
        auto lo_k = [](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | uint32_t(dk_major); };
        auto lo_mn = [](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | uint32_t(dmn_major); };
        auto mk = [](uint32_t hi, uint32_t lo) { return (uint64_t(hi) << 32) | lo; };
        auto mma_kmajor = [&](uint32_t d_t, uint32_t a0, uint32_t b0, uint32_t idesc) {
          const uint32_t la = lo_k(a0), lb = lo_k(b0);
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk) {
            const uint32_t off = ((kk / 4) * Cfg::kPanelBytes + (kk % 4) * 32) >> 4;
            mma_ss<false>(d_t, mk(hi_k, la + off), mk(hi_k, lb + off), idesc, kk > 0);
          }
        };
        auto mma_ts_mn = [&](uint32_t d_t, uint32_t a_t, uint32_t b0, bool acc) {
          const uint32_t lb = lo_mn(b0);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)   // 16 queries = 8 packed columns; query half h lives at columns 64 h
            mma_ts<false>(d_t, a_t + (kk >> 2) * 64 + (kk & 3) * 8, mk(hi_mn, lb + kk * (2048 >> 4)), id_dkv,
                          (acc || kk > 0) ? 1u : 0u);
        };
        const uint32_t lo_ds = lo_mn(sDS), lo_kmn = lo_mn(sK);
        mbar_wait(bar(B_KV), 0, 500);
        mbar_wait(bar(B_QF0), 0, 501);
        tc_fence_after();
        mma_kmajor(tS, sK, sQ, id_s);
        mma_commit(bar(B_SF));
        mbar_wait(bar(B_DOF), 0, 502);
        tc_fence_after();
        mma_kmajor(tDP, sV, sDO, id_s);
        mma_commit(bar(B_DPF));
        for (int it = 0; it < n_iter; ++it) {
          const int st = it & 1;
          const uint32_t ph = it & 1;
          mbar_wait(bar(B_PF), ph, 510);
          tc_fence_after();
          mma_ts_mn(tDV, tS, sDO, it > 0);
          mma_commit(bar(B_DOE));
          mbar_wait(bar(B_DSF), ph, 511);
          tc_fence_after();
          mma_ts_mn(tDK, tS + 32, sQ + st * Cfg::kTileBytes, it > 0);
          mma_commit(bar(B_QE0 + st));
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            mma_ss<false>(tDP, mk(hi_mn, lo_ds + kk * (2048 >> 4)), mk(hi_mn, lo_kmn + kk * (2048 >> 4)), id_dq, kk > 0);
          mma_commit(bar(B_DQF));
          if (it + 1 < n_iter) {
            const int st1 = st ^ 1;
            mbar_wait(bar(B_QF0 + st1), ((it + 1) >> 1) & 1, 512);
            tc_fence_after();
            mma_kmajor(tS, sK, sQ + st1 * Cfg::kTileBytes, id_s);
            mma_commit(bar(B_SF));
            mbar_wait(bar(B_DQE), ph, 513);           // dQ_it drained out of TMEM [128,256)
            mbar_wait(bar(B_DOF), (it + 1) & 1, 514);
            tc_fence_after();
            mma_kmajor(tDP, sV, sDO, id_s);
            mma_commit(bar(B_DPF));
          }
        }
        mma_commit(bar(B_DKV));
      }
      __syncwarp();
    }
  } else if (warp >= 8) {
    // ================================================================ dQ drain: TMEM -> smem -> TMA reduce-add
    reg_dealloc<Cfg::kRegsDrain>();
    const int wq = warp & 3;
    const int r = wq * 32 + lane;  // query row within the block
    const uint32_t t_dq = tDP + (uint32_t(wq * 32) << 16);
    const bool leader = (warp == 8 && lane == 0);
    const WsRows ws = ws_rows(p.cu_seqlens_q, bidb, p.h, 0, p.sq_r, p.ws_head_pitch);   // tensor-map coordinates only
    int nbuf = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int hq = hk * p.g + it / n_m;
      const int m = m_min + it % n_m;
      mbar_wait(bar(B_DQF), it & 1, 600);
      tc_fence_after();
      // Deterministic mode (reference: flash_api.cpp:928-929 gives every key-block column its own dq_accum slice;
      // FA4 orders the reduce-adds with semaphores, interface.py:2199-2209 -- same idea here): the key blocks that
      // touch query block m add their dQ partials in ascending key-block order.
      int* sem = nullptr;
      int my_turn = 0;
      if (p.dq_semaphore) {
        sem = p.dq_semaphore + (int64_t(bidb) * p.h + hq) * (p.sq_r / 128) + m;
        int n_lo = 0;
        if (p.has_left) {
          const int x = 128 * m - 127 + shift - p.wl;
          n_lo = x <= 0 ? 0 : (x + 127) / 128;
        }
        my_turn = n_block - n_lo;
        if (leader) {
          int cur;
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(cur) : "l"(sem) : "memory");
          } while (cur != my_turn);
        }
      }
#pragma unroll
      for (int half = 0; half < kD / 64; ++half) {
        uint32_t v[64];
        tmem_ld_x32(t_dq + half * 64, v);
        tmem_ld_x32(t_dq + half * 64 + 32, v + 32);
        tmem_wait_ld();
        if (half == kD / 64 - 1) {  // all of dQ is out of TMEM [128,256): hand it back to the MMA warp (next dP^T)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(B_DQE));
        }
#pragma unroll
        for (int c = 0; c < 2; ++c, ++nbuf) {
          const int buf = nbuf & 1;
          // the TMA reduce that last read this staging buffer must have finished reading it
          if (leader && nbuf >= 2) tma_store_wait_read<1>();
          named_bar_sync(1, 128);
          uint8_t* dst = smem_gen + Cfg::kSDQ + buf * Cfg::kPanelBytes + r * 128;
#pragma unroll
          for (int x = 0; x < 8; ++x)
            *reinterpret_cast<uint4*>(dst + ((x ^ (r & 7)) * 16)) =
                make_uint4(v[c * 32 + 4 * x], v[c * 32 + 4 * x + 1], v[c * 32 + 4 * x + 2], v[c * 32 + 4 * x + 3]);
          fence_proxy_async();
          named_bar_sync(2, 128);
          if (leader) {
            tma_reduce_add_4d(&p.tm_dq, sDQ + buf * Cfg::kPanelBytes, half * 64 + c * 32, ws.tma_row0 + m * 128, hq, ws.tma_b);
            tma_store_commit();
          }
        }
      }
      if (sem && leader) {  // all reduce-adds of this iteration have been performed: pass the turn on
        tma_store_wait<0>();
        asm volatile("fence.proxy.async;" ::: "memory");
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(sem), "r"(my_turn + 1) : "memory");
      }
    }
    if (leader) tma_store_wait<0>();
  } else {
    // ================================================================ compute warps
    reg_alloc<Cfg::kRegsCompute>();
    const int wq = warp & 3, hf = warp >> 2;          // key-row quadrant, query half
    const int r = wq * 32 + lane;                       // key row within the block
    const int j = key0 + r;                             // key index in the sequence
    const uint32_t lane_sel = uint32_t(wq * 32) << 16;
    const uint32_t t_s = tS + lane_sel + hf * 64;       // my 64 S^T columns
    const uint32_t t_dp = tDP + lane_sel + hf * 64;
    const uint32_t t_p = t_s;                           // P^T  packed: 32 columns inside my own S^T columns
    const uint32_t t_ds = t_s + 32;                     // dS^T packed: the other 32
    const float scale_log2 = p.scale_log2;
    // visible queries for this key: i in [i_lo, i_hi]
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
      const int st = it & 1;
      const int q0 = m * 128 + hf * 64;  // first query of my half
      float slope = 0.f;
      if (p.alibi_slopes) slope = p.alibi_slopes[bidb * p.alibi_batch_stride + hq] / p.scale_eff;
      mbar_wait(bar(B_QF0 + st), (it >> 1) & 1, 700);   // lse/dpsum of this stage are in smem
      mbar_wait(bar(B_SF), it & 1, 701);
      tc_fence_after();
      uint32_t su[64];
      tmem_ld_x32(t_s, su);
      tmem_ld_x32(t_s + 32, su + 32);
      tmem_wait_ld();
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
      // ---- phase 1: P^T = exp2(S^T * c - lse_log2[q])
      float pr[64];
#pragma unroll
      for (int c4 = 0; c4 < 64; c4 += 4) {
        const float4 l4 = *reinterpret_cast<const float4*>(s_lse + c4);
        pr[c4 + 0] = ex2(fmaf(s[c4 + 0], scale_log2, -l4.x));
        pr[c4 + 1] = ex2(fmaf(s[c4 + 1], scale_log2, -l4.y));
        pr[c4 + 2] = ex2(fmaf(s[c4 + 2], scale_log2, -l4.z));
        pr[c4 + 3] = ex2(fmaf(s[c4 + 3], scale_log2, -l4.w));
      }
      if (need_mask) {
        const int lo = i_lo - q0, hi = i_hi - q0;
#pragma unroll
        for (int c = 0; c < 64; ++c) pr[c] = (!key_valid || c < lo || c > hi) ? 0.f : pr[c];
      }
      // dropout (dropout.h:apply_dropout): dV sees the thinned P, dS keeps the full P and thins dP instead
      // (flash_bwd_kernel.h:560-596).  keep_lo/hi bit c = query q0 + c is kept for this key.
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
        if (lane == 0) mbar_arrive(bar(B_PF));
      }
      // ---- phase 2: dS^T = P^T * (dP^T - D[q])  (* (1 - tanh^2) with softcap)
      mbar_wait(bar(B_DPF), it & 1, 702);
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
          if constexpr (kSoftcap) {  // d tanh(x) = 1 - tanh^2 (utils.h:403-409); undo the ALiBi bias first
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
      tmem_st_x32(t_ds, dk_);
      {
        // shared copy for dQ = dS K: panel hf (64 queries = 128 B per key row), 16-byte chunks XOR-swizzled
        uint8_t* dst = smem_gen + Cfg::kSDS + hf * Cfg::kPanelBytes + r * 128;
#pragma unroll
        for (int x = 0; x < 8; ++x)
          *reinterpret_cast<uint4*>(dst + ((x ^ (r & 7)) * 16)) =
              make_uint4(dk_[4 * x], dk_[4 * x + 1], dk_[4 * x + 2], dk_[4 * x + 3]);
      }
      tmem_wait_st();
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_DSF));
    }
    // ---------------------------------------------------------------- epilogue: dK (x scale), dV
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
      constexpr int kHalf = kD / 2;  // columns handled by each query-half warp in the epilogue
#pragma unroll
      for (int cc = 0; cc < kHalf; cc += 32) {
        uint32_t o[32];
        if (n_iter > 0) {
          tmem_ld_x32(t_src + hf * kHalf + cc, o);
          tmem_wait_ld();
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace fa
