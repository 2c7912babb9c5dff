/*
 * fa_b200.h — C ABI of the B200-native fused attention library (libfa_b200.so).
 *
 * Drop-in boundary for the reference's `flash_attn_2_cuda` operator module
 * (csrc/flash_attn/flash_api.cpp:1535-1542).  Every entry point takes plain
 * device pointers, element strides and sizes — no torch / ATen types — plus the
 * CUDA stream to launch on.  The reference-side binding (pybind / ctypes stub) is
 * shown in INTEGRATION.md; the in-repo mirror is
 * flash_attention_b200/flash_attn_2_cuda.py.
 *
 *   reference host entry (flash_api.cpp)            this ABI
 *   ------------------------------------            ---------------------------
 *   mha_fwd            :368-536                     fa_b200_mha_fwd
 *   mha_varlen_fwd     :538-788                     fa_b200_mha_varlen_fwd
 *   mha_bwd            :800-1008                    fa_b200_mha_bwd
 *   mha_varlen_bwd     :1010-1241                   fa_b200_mha_varlen_bwd
 *   mha_fwd_kvcache    :1243-1532                   fa_b200_mha_fwd_kvcache
 *   num_splits_heuristic :281-315                   fa_b200_num_splits_heuristic
 *   combine_attn_seqk_parallel (flash_fwd_kernel.h:1117) fa_b200_fwd_combine
 *   Flash_fwd_params / Flash_bwd_params (flash.h:20-189) fa_b200_fwd_args / fa_b200_bwd_args
 *
 * Conventions
 *   - all strides are in ELEMENTS of the tensor's dtype (like flash.h `index_t`), last-dim stride is 1
 *   - return value: 0 = ok, non-zero = error; the message is available from
 *     fa_b200_last_error() (thread-local).  Messages mirror the reference's TORCH_CHECK texts.
 *   - every call is asynchronous on `stream`; no host sync is performed.
 *   - scratch memory is caller-owned: query the size with the *_workspace_bytes
 *     function, allocate (any 256-byte aligned device buffer), pass it in `workspace`.
 */
#ifndef FA_B200_H_
#define FA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fa_b200_stream_t; /* cudaStream_t */

enum fa_b200_dtype {
  FA_B200_FP16 = 0,
  FA_B200_BF16 = 1,
  FA_B200_FP8_E4M3 = 2 /* forward only; output is bf16 (spec: hopper/flash_api.cpp mha_fwd, flash_attn/cute/interface.py:637-643) */
};

enum fa_b200_entry { FA_B200_ENTRY_INFER = 0, FA_B200_ENTRY_DENSE = 1, FA_B200_ENTRY_VARLEN = 2, FA_B200_ENTRY_KVCACHE = 3 };

/* Forward arguments (dense, varlen, paged-KV decode share one struct; unused pointers are NULL). */
typedef struct fa_b200_fwd_args {
  /* tensors */
  const void* q;   /* dense (b, sq, h, d) | varlen (total_q, h, d) */
  const void* k;   /* dense (b, sk, h_k, d) | varlen (total_k, h_k, d) | paged (num_blocks, page, h_k, d) */
  const void* v;
  void* out;       /* same layout as q; dtype = q's dtype (bf16 for fp8 inputs) */
  float* softmax_lse; /* dense (b, h, sq) | varlen (h, total_q), fp32 */
  int64_t q_batch_stride, q_row_stride, q_head_stride;
  int64_t k_batch_stride, k_row_stride, k_head_stride; /* paged: batch stride = stride between pages */
  int64_t v_batch_stride, v_row_stride, v_head_stride;
  int64_t o_batch_stride, o_row_stride, o_head_stride;
  /* sizes */
  int32_t b, seqlen_q, seqlen_k, h, h_k, d;  /* varlen: seqlen_* are the max_seqlen_* arguments */
  int32_t total_q, total_k;                  /* varlen only (rows of the packed tensors), else 0 */
  int32_t dtype;                             /* enum fa_b200_dtype */
  int32_t b_kvcache;                         /* kvcache with cache_batch_idx: batch extent of k/v (0 = same as b) */
  /* varlen / kvcache metadata (device pointers, int32) */
  const int32_t* cu_seqlens_q; /* (b+1) or NULL */
  const int32_t* cu_seqlens_k; /* (b+1) or NULL */
  const int32_t* seqused_k;    /* (b) or NULL: number of valid keys per batch (cache_seqlens) */
  const int32_t* seqused_q;    /* (b) or NULL: number of valid queries per batch (FA3 `seqused_q`); rows beyond are not written */
  const int32_t* cache_batch_idx; /* (b) or NULL: batch b reads cache row cache_batch_idx[b] (kvcache, not paged) */
  const int32_t* leftpad_k;    /* (b) or NULL: keys of batch b start leftpad_k[b] rows into its rows and seqused_k counts from
                                * the start of those rows (kvcache and varlen, not paged; block_info.h:12-45) */
  const int32_t* block_table;  /* (b, max_blocks_per_seq) or NULL */
  int64_t block_table_batch_stride;
  int32_t page_block_size;     /* multiple of 128 (reference requires 256: flash_api.cpp:1305) */
  int32_t num_pages;
  /* masks / score mods */
  const float* alibi_slopes;   /* (h) or (b,h) fp32, or NULL */
  int64_t alibi_slopes_batch_stride; /* 0 for (h) */
  float softmax_scale;
  float softcap;               /* 0 = off */
  int32_t is_causal;
  int32_t window_size_left, window_size_right; /* -1 = infinite */
  /* fp8 de-scales, (b, h_k) fp32 each or NULL (= 1.0) */
  const float* q_descale; const float* k_descale; const float* v_descale;
  int64_t q_descale_batch_stride, k_descale_batch_stride, v_descale_batch_stride;
  /* split-KV */
  int32_t num_splits;          /* 0 = heuristic, 1 = never split, >1 = forced */
  int32_t pack_gqa;            /* 0 = heuristic, 1 = pack the query heads of a KV head into M, 2 = never (FA3 `pack_gqa`) */
  void* workspace; size_t workspace_bytes;
  /* debugging: optional fp32 buffer (128x128) receiving the raw S tile of CTA (0,0,0) */
  float* debug_s;
  /* attention dropout (mha_fwd's p_dropout / rng_state, flash_api.cpp:370,505-517): Philox4x32-7 keyed by
   * (philox_seed, philox_offset) -- the two int64 the reference returns as `rng_state`.  0 = off.
   * Not available for kvcache, fp8, split-KV or together with softcap (flash_api.cpp:391). */
  float p_dropout;
  uint64_t philox_seed, philox_offset;
  /* return_softmax: optional (b, h, round128(seqlen_q), round128(seqlen_k)) buffer in q's dtype receiving the
   * normalised probabilities, negated where dropped ("for testing only", flash_attn_interface.py:1196) */
  void* s_dmask;
  /* which entry point these args are for (enum fa_b200_entry).  The launch functions ignore it; the planner queries
   * (fa_b200_fwd_workspace_bytes / _num_splits / _plan) use it so that their answer matches the launch -- the kvcache
   * entry plans its splits differently from the dense one even when neither cache_seqlens nor block_table is given.
   * 0 = infer from the pointers (cu_seqlens_q -> varlen, block_table | seqused_k -> kvcache, else dense). */
  int32_t entry;
  /* FA3 `sm_margin`: leave this many SMs free (persistent grids are sized num_SMs - sm_margin) */
  int32_t sm_margin;
} fa_b200_fwd_args;

typedef struct fa_b200_bwd_args {
  const void* dout; const void* q; const void* k; const void* v; const void* out;
  const float* softmax_lse;
  void* dq; void* dk; void* dv;           /* written in place, same layouts as q/k/v */
  float* softmax_d;                       /* dense (b, h, round128(sq)) | varlen (h, total_q + 128*b) */
  int64_t q_batch_stride, q_row_stride, q_head_stride;
  int64_t k_batch_stride, k_row_stride, k_head_stride;
  int64_t v_batch_stride, v_row_stride, v_head_stride;
  int64_t o_batch_stride, o_row_stride, o_head_stride;
  int64_t do_batch_stride, do_row_stride, do_head_stride;
  int64_t dq_batch_stride, dq_row_stride, dq_head_stride;
  int64_t dk_batch_stride, dk_row_stride, dk_head_stride;
  int64_t dv_batch_stride, dv_row_stride, dv_head_stride;
  int32_t b, seqlen_q, seqlen_k, h, h_k, d;
  int32_t total_q, total_k;
  int32_t dtype;
  const int32_t* cu_seqlens_q; const int32_t* cu_seqlens_k;
  const float* alibi_slopes; int64_t alibi_slopes_batch_stride;
  float softmax_scale; float softcap;
  int32_t is_causal; int32_t window_size_left, window_size_right;
  int32_t deterministic;
  void* workspace; size_t workspace_bytes;
  /* dropout of the forward being differentiated: p_dropout and its rng_state (flash_api.cpp:810,990-1003).
   * `rng_state` (device pointer to {seed, offset}, the tensor the forward returned) wins over the two host values */
  float p_dropout;
  uint64_t philox_seed, philox_offset;
  const uint64_t* rng_state;
  int32_t sm_margin;                      /* FA3 `sm_margin` (see fa_b200_fwd_args) */
} fa_b200_bwd_args;

/* ---- forward ---- */
int fa_b200_mha_fwd(const fa_b200_fwd_args* args, fa_b200_stream_t stream);
int fa_b200_mha_varlen_fwd(const fa_b200_fwd_args* args, fa_b200_stream_t stream);
int fa_b200_mha_fwd_kvcache(const fa_b200_fwd_args* args, fa_b200_stream_t stream);
/* bytes of scratch the call will need with these args (0 if it will not split) */
size_t fa_b200_fwd_workspace_bytes(const fa_b200_fwd_args* args);
/* the split count the call will use (after heuristic / clamping) */
int fa_b200_fwd_num_splits(const fa_b200_fwd_args* args);
/* the full plan: out[0] = num_splits, out[1] = GQA pack factor, out[2] = number of 256-row m-blocks (FA3 get_scheduler_metadata) */
int fa_b200_fwd_plan(const fa_b200_fwd_args* args, int32_t out[3]);
/* which forward kernel the call will launch: 1 = one query tile per CTA (two CTAs per SM: decode, short items),
 * 2 = two query tiles per CTA (the throughput configuration), 3 = the wide-S kernel (S = Q K^T 256 keys per
 * instruction; dense fp16 / bf16, head dim <= 128, no dropout / softcap / ALiBi / split-KV; opt-in: FA_B200_FWD_WIDE=1);
 * negative on an argument error.  (The reference picks its kernel inside run_mha_fwd, flash_api.cpp:243-279.) */
int fa_b200_fwd_kernel_kind(const fa_b200_fwd_args* args);

/* ---- backward ---- */
int fa_b200_mha_bwd(const fa_b200_bwd_args* args, fa_b200_stream_t stream);
int fa_b200_mha_varlen_bwd(const fa_b200_bwd_args* args, fa_b200_stream_t stream);
size_t fa_b200_bwd_workspace_bytes(const fa_b200_bwd_args* args);

/* ---- split-KV combine as a standalone op (hopper/flash_api.cpp fwd_combine, flash_attn/cute/interface.py:3687) ----
 * o_partial (splits, rows, d) fp32, lse_partial (splits, rows) fp32 -> out (rows, d) `dtype`, lse (rows) fp32 */
int fa_b200_fwd_combine(const float* o_partial, const float* lse_partial, void* out, float* lse,
                        int32_t num_splits, int64_t rows, int32_t d, int32_t dtype,
                        fa_b200_stream_t stream);

/* ---- decode prologue: append k/v into the cache (+ rotary on the new keys and on q), in place
 *      (mha_fwd_kvcache's knew/vnew/rotary_cos/rotary_sin arguments, flash_api.cpp:1243-1264, flash_fwd_kernel.h:691-828) */
typedef struct fa_b200_append_args {
  void* kcache; void* vcache;           /* (b_c, sk, h_k, d) | paged (num_pages, page, h_k, d) */
  const void* knew; const void* vnew;   /* (b, s_new, h_k, d) */
  const void* q; void* q_out;           /* (b, sq, h, d): q_out receives the rotated q (NULL without rotary) */
  int64_t kc_batch_stride, kc_row_stride, kc_head_stride, vc_batch_stride, vc_row_stride, vc_head_stride;
  int64_t kn_batch_stride, kn_row_stride, kn_head_stride, vn_batch_stride, vn_row_stride, vn_head_stride;
  int64_t q_batch_stride, q_row_stride, q_head_stride, qo_batch_stride, qo_row_stride, qo_head_stride;
  const int32_t* cache_seqlens;         /* (b) lengths BEFORE the append */
  const int32_t* cache_batch_idx; const int32_t* leftpad_k;
  const int32_t* block_table; int64_t block_table_batch_stride; int32_t page_block_size;
  const void* rotary_cos; const void* rotary_sin; /* (seqlen_ro, rotary_dim/2), dtype of q, or NULL */
  int32_t rotary_dim; int32_t is_rotary_interleaved;
  int32_t q_pos_per_row;                /* 1: query i is rotated at cache_seqlens+i (causal/local); 0: all at cache_seqlens */
  int32_t b, seqlen_new, seqlen_q, h, h_k, d, dtype;
} fa_b200_append_args;
int fa_b200_kvcache_append(const fa_b200_append_args* args, fa_b200_stream_t stream);
size_t fa_b200_sizeof_append_args(void);

/* ---- host-buffer plumbing: pitched copy between a (pinned) host tensor slice and a dense device buffer, used by the
 *      host-buffer pipeline to move one (batch chunk, head group) at a time.  kind: 1 = host -> device, 2 = device -> host.
 *      `height` rows of `width_bytes`; pitches in bytes.  Asynchronous on `stream`. */
int fa_b200_memcpy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                           size_t height, int32_t kind, fa_b200_stream_t stream);

/* writes {seed, offset} into a device (2,) uint64 buffer on `stream`: the `rng_state` tensor mha_fwd returns
 * (flash_api.cpp:501-517), filled without a host->device copy */
int fa_b200_write_rng_state(uint64_t* dst, uint64_t seed, uint64_t offset, fa_b200_stream_t stream);

/* ---- helpers ---- */
int fa_b200_num_splits_heuristic(int batch_nheads_mblocks, int num_sms, int num_n_blocks, int max_splits);
const char* fa_b200_last_error(void);
/* ABI guards for FFI bindings: sizeof the argument structs as compiled into the library */
size_t fa_b200_sizeof_fwd_args(void);
size_t fa_b200_sizeof_bwd_args(void);
const char* fa_b200_version(void);
/* number of kernel launches issued by this library since load (for bench.py's gpu_launches) */
uint64_t fa_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FA_B200_H_ */
