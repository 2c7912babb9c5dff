// flash_attn_3_ops.cpp — TORCH_LIBRARY(flash_attn_3): the FA3 operator library on the fa_b200 kernels.
//
// Same schemas as the reference's hopper/flash_api.cpp:1673-1768 (fwd / bwd / fwd_combine / get_scheduler_metadata), so
// that hopper/flash_attn_interface.py (`import flash_attn_3._C; torch.ops.flash_attn_3.fwd(...)`) runs unchanged once
// flash_attention_b200.install_fa3() has registered this shared object under that module name.  One `fwd` op covers
// dense, packed-varlen and KV-cache calls (append + rotary, page_table, kv_batch_idx, leftpad_k, seqused_q/k), fp8 e4m3
// inputs with per-(batch, kv-head) de-scales on every path, num_splits and pack_gqa; arguments this library has no
// kernel for (qv, attention_chunk, cu_seqlens_k_new, seqlens_rotary, V head dim != QK head dim) raise.
//
// Host logic restated from hopper/flash_api.cpp mha_fwd :672-1197, mha_bwd :1267-1567, mha_combine :1570-1670.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include <cmath>
#include <tuple>

#include "fa_torch_common.h"

namespace fa_ext {

static void unsupported(bool present, const char* name) {
  TORCH_CHECK(!present, "fa_b200: `", name, "` is not supported by this build");
}

std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor> fa3_fwd(
    at::Tensor q, at::Tensor k, at::Tensor v, std::optional<at::Tensor> k_new_, std::optional<at::Tensor> v_new_,
    std::optional<at::Tensor> q_v_, std::optional<at::Tensor> out_, std::optional<at::Tensor> cu_seqlens_q_,
    std::optional<at::Tensor> cu_seqlens_k_, std::optional<at::Tensor> cu_seqlens_k_new_, std::optional<at::Tensor> seqused_q_,
    std::optional<at::Tensor> seqused_k_, std::optional<int64_t> max_seqlen_q_, std::optional<int64_t> max_seqlen_k_,
    std::optional<at::Tensor> page_table_, std::optional<at::Tensor> kv_batch_idx_, std::optional<at::Tensor> leftpad_k_,
    std::optional<at::Tensor> rotary_cos_, std::optional<at::Tensor> rotary_sin_, std::optional<at::Tensor> seqlens_rotary_,
    std::optional<at::Tensor> q_descale_, std::optional<at::Tensor> k_descale_, std::optional<at::Tensor> v_descale_,
    std::optional<double> softmax_scale_, bool is_causal, int64_t window_size_left, int64_t window_size_right,
    int64_t attention_chunk, double softcap, bool is_rotary_interleaved, std::optional<at::Tensor> scheduler_metadata_,
    int64_t num_splits, std::optional<bool> pack_gqa_, int64_t sm_margin) {
  unsupported(q_v_.has_value(), "qv");
  unsupported(attention_chunk != 0, "attention_chunk");
  unsupported(cu_seqlens_k_new_.has_value(), "cu_seqlens_k_new");
  unsupported(seqlens_rotary_.has_value(), "seqlens_rotary");
  const auto q_type = q.scalar_type();
  TORCH_CHECK(q_type == at::kHalf || q_type == at::kBFloat16 || q_type == at::kFloat8_e4m3fn,
              "FlashAttention only supports fp16, bf16, and fp8_e4m3 data type");
  TORCH_CHECK(k.scalar_type() == q_type, "query and key must have the same dtype");
  TORCH_CHECK(v.scalar_type() == q_type, "query and value must have the same dtype");
  TORCH_CHECK(q.is_cuda() && k.is_cuda() && v.is_cuda(), "Input tensor must be on CUDA device");
  TORCH_CHECK(q.stride(-1) == 1, "Input tensor must have contiguous last dimension");
  TORCH_CHECK(k.stride(-1) == 1, "Input tensor must have contiguous last dimension");
  TORCH_CHECK(v.stride(-1) == 1, "Input tensor must have contiguous last dimension");
  c10::cuda::CUDAGuard guard(q.device());
  const bool fp8 = q_type == at::kFloat8_e4m3fn;
  const bool paged = page_table_.has_value();
  const bool varlen_q = cu_seqlens_q_.has_value(), varlen_k = cu_seqlens_k_.has_value();
  if (paged) {
    TORCH_CHECK(page_table_.value().is_cuda(), "Input tensor must be on CUDA device");
    TORCH_CHECK(page_table_.value().dtype() == at::kInt, "page_table must have dtype torch.int32");
    TORCH_CHECK(page_table_.value().stride(-1) == 1, "page_table must have contiguous last dimension");
  }
  if (varlen_q) {
    TORCH_CHECK(cu_seqlens_q_.value().dtype() == at::kInt, "cu_seqlens_q must have dtype torch.int32");
    TORCH_CHECK(cu_seqlens_q_.value().is_cuda() && cu_seqlens_q_.value().is_contiguous(), "cu_seqlens_q must be a contiguous CUDA tensor");
    TORCH_CHECK(max_seqlen_q_.has_value(), "max_seqlen_q must be provided if cu_seqlens_q is provided");
  }
  if (varlen_k) {
    TORCH_CHECK(cu_seqlens_k_.value().dtype() == at::kInt, "cu_seqlens_k must have dtype torch.int32");
    TORCH_CHECK(cu_seqlens_k_.value().is_cuda() && cu_seqlens_k_.value().is_contiguous(), "cu_seqlens_k must be a contiguous CUDA tensor");
    TORCH_CHECK(max_seqlen_k_.has_value(), "max_seqlen_k must be provided if cu_seqlens_k is provided");
    TORCH_CHECK(!paged, "If cu_seqlens_k is passed in, then page table is not supported");
    TORCH_CHECK(!kv_batch_idx_.has_value(), "If cu_seqlens_k is passed in, then page table is not supported");
  }
  TORCH_CHECK(varlen_q || !varlen_k, "fa_b200: packed keys with batched (non-packed) queries are not supported");
  const int64_t b = !varlen_q ? q.size(0) : cu_seqlens_q_.value().size(0) - 1;
  const int64_t sq = !varlen_q ? q.size(1) : max_seqlen_q_.value();
  const int64_t total_q = !varlen_q ? b * q.size(1) : q.size(0);
  const int64_t h = q.size(-2), d = q.size(-1), dv = v.size(-1);
  const int64_t max_pages = paged ? page_table_.value().size(1) : 0;
  const int64_t num_pages = paged ? k.size(0) : 0, page = paged ? k.size(1) : 1;
  const int64_t sk = !varlen_k ? (!paged ? k.size(1) : max_pages * page) : max_seqlen_k_.value();
  const int64_t total_k = !varlen_k ? 0 : k.size(0);
  const int64_t hk = k.size(-2);
  const int64_t b_k = paged ? page_table_.value().size(0) : (!varlen_k ? k.size(0) : cu_seqlens_k_.value().size(0) - 1);
  const double softmax_scale = softmax_scale_.has_value() ? softmax_scale_.value() : 1.0 / std::sqrt(double(d));
  if (!kv_batch_idx_.has_value()) TORCH_CHECK(b == b_k, "batch_size must be equal to batch_size_k");
  TORCH_CHECK(d <= 256, "FlashAttention forward only supports head dimension at most 256");
  TORCH_CHECK(h % hk == 0, "Number of heads in key/value must divide number of heads in query");
  TORCH_CHECK(dv == d, "fa_b200: V head dimension must equal the Q/K head dimension");
  const int alignment = fp8 ? 16 : 8;
  TORCH_CHECK(d % alignment == 0, "head_size should be a multiple of ", alignment);
  if (!varlen_q) check_shape(q, {b, sq, h, d}, "q"); else check_shape(q, {total_q, h, d}, "q");
  if (paged) {
    TORCH_CHECK(page % 128 == 0, "fa_b200: paged KV block size must be a multiple of 128");
    check_shape(k, {num_pages, page, hk, d}, "k");
    check_shape(v, {num_pages, page, hk, d}, "v");
    check_shape(page_table_.value(), {b_k, max_pages}, "page_table");
  } else if (!varlen_k) {
    check_shape(k, {b_k, sk, hk, d}, "k");
    check_shape(v, {b_k, sk, hk, d}, "v");
  } else {
    check_shape(k, {total_k, hk, d}, "k");
    check_shape(v, {total_k, hk, d}, "v");
    check_shape(cu_seqlens_k_.value(), {b + 1}, "cu_seqlens_k");
  }
  const auto out_type = fp8 ? at::kBFloat16 : q_type;
  at::Tensor out;
  if (out_.has_value()) {
    out = out_.value();
    TORCH_CHECK(out.scalar_type() == out_type, "For FP16/BF16 input, output must have the same dtype as inputs. For FP8 input, output must have dtype BF16");
    TORCH_CHECK(out.is_cuda(), "Input tensor must be on CUDA device");
    TORCH_CHECK(out.stride(-1) == 1, "Output tensor must have contiguous last dimension");
    if (!varlen_q) check_shape(out, {b, sq, h, d}, "out"); else check_shape(out, {total_q, h, d}, "out");
  } else {
    out = !varlen_q ? at::empty({b, sq, h, d}, q.options().dtype(out_type)) : at::empty({total_q, h, d}, q.options().dtype(out_type));
  }
  at::Tensor lse = !varlen_q ? at::empty({b, h, sq}, q.options().dtype(at::kFloat)) : at::empty({h, total_q}, q.options().dtype(at::kFloat));
  // out_accum / softmax_lse_accum are internal scratch here: two distinct empty placeholders (returns may not alias)
  at::Tensor none = at::empty({0}, q.options().dtype(at::kFloat)), none2 = at::empty({0}, q.options().dtype(at::kFloat));
  if (total_q == 0) return {out, lse, none, none2};
  if (sk == 0) {
    out.zero_();
    lse.fill_(std::numeric_limits<float>::infinity());
    return {out, lse, none, none2};
  }
  if (rotary_cos_.has_value()) {
    TORCH_CHECK(k_new_.has_value(), "If rotary cos/sin are provided, new key / value to be appended to KV cache must also be provided");
    TORCH_CHECK(rotary_sin_.has_value(), "If rotary cos is provided, rotary sin must also be provided");
  }
  // ---- append (k_new, v_new) into the cache first (kv-cache calls only)
  at::Tensor seqlens_eff;
  at::Tensor q_used = q;
  if (k_new_.has_value()) {
    TORCH_CHECK(!fp8, "fa_b200: appending new keys to an fp8 cache is not supported");
    TORCH_CHECK(!varlen_q && !varlen_k, "fa_b200: k_new with packed sequences is not supported");
    TORCH_CHECK(v_new_.has_value(), "If k_new is supplied, v_new must also be passed in");
    TORCH_CHECK(seqused_k_.has_value(), "If k_new is supplied, seqlens_k must also be passed in");
    const auto& kn = k_new_.value();
    const auto& vn = v_new_.value();
    const int64_t s_new = kn.size(1);
    TORCH_CHECK(kn.scalar_type() == q_type && vn.scalar_type() == q_type, "k_new / v_new must have the same dtype as query");
    TORCH_CHECK(kn.stride(-1) == 1 && vn.stride(-1) == 1, "k_new / v_new must have contiguous last dimension");
    check_shape(kn, {b, s_new, hk, d}, "k_new");
    check_shape(vn, {b, s_new, hk, d}, "v_new");
    const bool local = window_size_left >= 0 || window_size_right >= 0;
    fa_b200_append_args ap = {};
    ap.kcache = k.data_ptr(); ap.vcache = v.data_ptr();
    ap.kc_batch_stride = k.stride(0); ap.kc_row_stride = k.stride(1); ap.kc_head_stride = k.stride(2);
    ap.vc_batch_stride = v.stride(0); ap.vc_row_stride = v.stride(1); ap.vc_head_stride = v.stride(2);
    ap.knew = kn.data_ptr(); ap.vnew = vn.data_ptr();
    ap.kn_batch_stride = kn.stride(0); ap.kn_row_stride = kn.stride(1); ap.kn_head_stride = kn.stride(2);
    ap.vn_batch_stride = vn.stride(0); ap.vn_row_stride = vn.stride(1); ap.vn_head_stride = vn.stride(2);
    ap.seqlen_new = int(s_new);
    if (rotary_cos_.has_value()) {
      const auto& rc = rotary_cos_.value();
      const auto& rs = rotary_sin_.value();
      TORCH_CHECK(rc.is_cuda() && rc.is_contiguous() && rs.is_cuda() && rs.is_contiguous(), "rotary_cos / rotary_sin must be contiguous CUDA tensors");
      TORCH_CHECK(rc.scalar_type() == q_type && rs.scalar_type() == q_type, "rotary_cos must have the same dtype as query");
      const int64_t rotary_dim = rc.size(1) * 2;
      TORCH_CHECK(rotary_dim <= d, "rotary_dim must be <= headdim");
      TORCH_CHECK(rotary_dim % 16 == 0, "Only rotary dimensions divisible by 16 are currently supported");
      TORCH_CHECK(rc.size(0) >= sk, "cos/sin seqlen must be at least the seqlen of KV cache");
      check_shape(rs, {rc.size(0), rotary_dim / 2}, "rotary_sin");
      auto q_rot = at::empty_like(q);
      ap.q = q.data_ptr(); ap.q_out = q_rot.data_ptr();
      ap.q_batch_stride = q.stride(0); ap.q_row_stride = q.stride(1); ap.q_head_stride = q.stride(2);
      ap.qo_batch_stride = q_rot.stride(0); ap.qo_row_stride = q_rot.stride(1); ap.qo_head_stride = q_rot.stride(2);
      ap.rotary_cos = rc.data_ptr(); ap.rotary_sin = rs.data_ptr();
      ap.rotary_dim = int(rotary_dim);
      ap.is_rotary_interleaved = is_rotary_interleaved;
      ap.q_pos_per_row = (is_causal || local) ? 1 : 0;
      q_used = q_rot;
    }
    ap.cache_seqlens = int_vec(seqused_k_, b, "seqused_k");
    ap.cache_batch_idx = int_vec(kv_batch_idx_, b, "kv_batch_idx");
    ap.leftpad_k = int_vec(leftpad_k_, b, "leftpad_k");
    if (paged) {
      ap.block_table = page_table_.value().data_ptr<int32_t>();
      ap.block_table_batch_stride = page_table_.value().stride(0);
      ap.page_block_size = int(page);
    }
    ap.b = int(b); ap.seqlen_q = int(sq); ap.h = int(h); ap.h_k = int(hk); ap.d = int(d);
    ap.dtype = dtype_code(q);
    check_rc(fa_b200_kvcache_append(&ap, cur_stream()));
    seqlens_eff = seqused_k_.value() + s_new;
  } else if (seqused_k_.has_value()) {
    seqlens_eff = seqused_k_.value();
  }
  // ---- the attention call
  fa_b200_fwd_args a = {};
  a.q = q_used.data_ptr(); a.k = k.data_ptr(); a.v = v.data_ptr(); a.out = out.data_ptr();
  a.softmax_lse = lse.data_ptr<float>();
  if (!varlen_q) {
    a.q_batch_stride = q_used.stride(0); a.q_row_stride = q_used.stride(1); a.q_head_stride = q_used.stride(2);
    a.o_batch_stride = out.stride(0); a.o_row_stride = out.stride(1); a.o_head_stride = out.stride(2);
  } else {
    a.q_row_stride = q_used.stride(0); a.q_head_stride = q_used.stride(1);
    a.o_row_stride = out.stride(0); a.o_head_stride = out.stride(1);
    a.cu_seqlens_q = cu_seqlens_q_.value().data_ptr<int32_t>();
    a.total_q = int(total_q);
  }
  if (varlen_k) {
    a.k_row_stride = k.stride(0); a.k_head_stride = k.stride(1);
    a.v_row_stride = v.stride(0); a.v_head_stride = v.stride(1);
    a.cu_seqlens_k = cu_seqlens_k_.value().data_ptr<int32_t>();
    a.total_k = int(total_k);
  } else {
    a.k_batch_stride = k.stride(0); a.k_row_stride = k.stride(1); a.k_head_stride = k.stride(2);
    a.v_batch_stride = v.stride(0); a.v_row_stride = v.stride(1); a.v_head_stride = v.stride(2);
  }
  if (paged) {
    a.block_table = page_table_.value().data_ptr<int32_t>();
    a.block_table_batch_stride = page_table_.value().stride(0);
    a.page_block_size = int(page); a.num_pages = int(num_pages);
  }
  a.b = int(b); a.seqlen_q = int(sq); a.seqlen_k = int(sk); a.h = int(h); a.h_k = int(hk); a.d = int(d);
  a.dtype = dtype_code(q);
  if (seqlens_eff.defined()) {
    TORCH_CHECK(seqlens_eff.dtype() == at::kInt && seqlens_eff.is_cuda() && seqlens_eff.is_contiguous() && seqlens_eff.numel() == b,
                "seqused_k must be a contiguous int32 CUDA tensor of shape (batch_size,)");
    a.seqused_k = seqlens_eff.data_ptr<int32_t>();
  }
  a.seqused_q = int_vec(seqused_q_, b, "seqused_q");
  a.cache_batch_idx = int_vec(kv_batch_idx_, b, "kv_batch_idx");
  a.leftpad_k = int_vec(leftpad_k_, b, "leftpad_k");
  if (!varlen_k && !paged) a.b_kvcache = int(k.size(0));
  std::vector<at::Tensor> keep;
  if (fp8) set_descales(a, q_descale_, k_descale_, v_descale_, b, hk, keep);
  a.softmax_scale = float(softmax_scale); a.softcap = float(softcap);
  a.is_causal = is_causal; a.window_size_left = int(window_size_left); a.window_size_right = int(window_size_right);
  a.num_splits = int(num_splits);
  if (scheduler_metadata_.has_value() && num_splits == 0 && !scheduler_metadata_.value().is_cuda())
    a.num_splits = scheduler_metadata_.value().data_ptr<int32_t>()[0];   // a host copy of get_scheduler_metadata's plan
  a.pack_gqa = pack_code(pack_gqa_);
  a.sm_margin = int(sm_margin);
  const bool cache_call = !varlen_q && (seqlens_eff.defined() || paged || kv_batch_idx_.has_value() || leftpad_k_.has_value() || k_new_.has_value());
  at::Tensor ws;
  if (varlen_q) {
    a.entry = FA_B200_ENTRY_VARLEN;
    a.num_splits = 1;
    ws = fwd_workspace(a, q);
    check_rc(fa_b200_mha_varlen_fwd(&a, cur_stream()));
  } else if (cache_call) {
    a.entry = FA_B200_ENTRY_KVCACHE;
    ws = fwd_workspace(a, q);
    check_rc(fa_b200_mha_fwd_kvcache(&a, cur_stream()));
  } else {
    a.entry = FA_B200_ENTRY_DENSE;
    ws = fwd_workspace(a, q);
    check_rc(fa_b200_mha_fwd(&a, cur_stream()));
  }
  return {out, lse, none, none2};
}

// mha_bwd (hopper/flash_api.cpp:1267-1567): dq / dk / dv are written in place (allocated when absent); returns
// (softmax_d, softmax_lse_log2, dq_accum, dk_accum, dv_accum) -- the last four are this library's internal scratch and
// come back as empty placeholders (the reference's interface only reads softmax_d: hopper/flash_attn_interface.py:262)
std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor> fa3_bwd(
    at::Tensor dout, at::Tensor q, at::Tensor k, at::Tensor v, at::Tensor out, at::Tensor softmax_lse,
    std::optional<at::Tensor> dq_, std::optional<at::Tensor> dk_, std::optional<at::Tensor> dv_,
    std::optional<at::Tensor> cu_seqlens_q_, std::optional<at::Tensor> cu_seqlens_k_, std::optional<at::Tensor> seqused_q_,
    std::optional<at::Tensor> seqused_k_, std::optional<int64_t> max_seqlen_q_, std::optional<int64_t> max_seqlen_k_,
    std::optional<double> softmax_scale_, bool is_causal, int64_t window_size_left, int64_t window_size_right, double softcap,
    bool deterministic, int64_t sm_margin) {
  unsupported(seqused_q_.has_value(), "seqused_q (backward)");
  unsupported(seqused_k_.has_value(), "seqused_k (backward)");
  const auto q_type = q.scalar_type();
  TORCH_CHECK(q_type == at::kHalf || q_type == at::kBFloat16, "FlashAttention only support fp16 and bf16 data type");
  TORCH_CHECK(k.scalar_type() == q_type, "query and key must have the same dtype");
  TORCH_CHECK(v.scalar_type() == q_type, "query and value must have the same dtype");
  TORCH_CHECK(out.scalar_type() == q_type, "query and out must have the same dtype");
  TORCH_CHECK(dout.scalar_type() == q_type, "query and dout must have the same dtype");
  TORCH_CHECK(q.is_cuda() && k.is_cuda() && v.is_cuda() && out.is_cuda() && dout.is_cuda() && softmax_lse.is_cuda(),
              "Input tensor must be on CUDA device");
  for (const at::Tensor* t : {&q, &k, &v, &out, &dout})
    TORCH_CHECK(t->stride(-1) == 1, "Input tensor must have contiguous last dimension");
  c10::cuda::CUDAGuard guard(q.device());
  const bool varlen = cu_seqlens_q_.has_value();
  TORCH_CHECK(varlen == cu_seqlens_k_.has_value(), "fa_b200: cu_seqlens_q and cu_seqlens_k must be given together (backward)");
  if (varlen) {
    TORCH_CHECK(max_seqlen_q_.has_value(), "max_seqlen_q must be provided if cu_seqlens_q is provided");
    TORCH_CHECK(max_seqlen_k_.has_value(), "max_seqlen_k must be provided if cu_seqlens_k is provided");
    for (const auto& c : {cu_seqlens_q_.value(), cu_seqlens_k_.value()})
      TORCH_CHECK(c.dtype() == at::kInt && c.is_cuda() && c.is_contiguous(), "cu_seqlens must be contiguous int32 CUDA tensors");
  }
  const int64_t b = !varlen ? q.size(0) : cu_seqlens_q_.value().size(0) - 1;
  const int64_t sq = !varlen ? q.size(1) : max_seqlen_q_.value();
  const int64_t sk = !varlen ? k.size(1) : max_seqlen_k_.value();
  const int64_t total_q = !varlen ? b * q.size(1) : q.size(0), total_k = !varlen ? b * k.size(1) : k.size(0);
  const int64_t h = q.size(-2), hk = k.size(-2), d = q.size(-1);
  TORCH_CHECK(v.size(-1) == d, "fa_b200: V head dimension must equal the Q/K head dimension");
  TORCH_CHECK(d % 8 == 0, "head_size should be a multiple of 8");
  TORCH_CHECK(d <= 256, "FlashAttention backward only supports head dimension at most 256");
  TORCH_CHECK(h % hk == 0, "Number of heads in key/value must divide number of heads in query");
  const double softmax_scale = softmax_scale_.has_value() ? softmax_scale_.value() : 1.0 / std::sqrt(double(d));
  TORCH_CHECK(softmax_lse.dtype() == at::kFloat && softmax_lse.is_contiguous(), "softmax_lse must be a contiguous fp32 tensor");
  auto alloc = [&](std::optional<at::Tensor>& t_, const at::Tensor& like, const char* name) {
    if (!t_.has_value()) return at::empty_like(like);
    TORCH_CHECK(t_.value().scalar_type() == q_type && t_.value().is_cuda() && t_.value().stride(-1) == 1 &&
                t_.value().sizes() == like.sizes(), name, " must have the dtype and shape of its input and a contiguous last dimension");
    return t_.value();
  };
  at::Tensor dq = alloc(dq_, q, "dq"), dk = alloc(dk_, k, "dk"), dv = alloc(dv_, v, "dv");
  at::Tensor softmax_d = !varlen ? at::empty({b, h, round128(sq)}, q.options().dtype(at::kFloat))
                                 : at::empty({h, total_q + 128 * b}, q.options().dtype(at::kFloat));
  auto empty_f = [&] { return at::empty({0}, q.options().dtype(at::kFloat)); };
  if (total_q == 0) return {softmax_d, empty_f(), empty_f(), empty_f(), empty_f()};
  if (sk == 0 || total_k == 0) {
    dq.zero_(); dk.zero_(); dv.zero_(); softmax_d.zero_();
    return {softmax_d, empty_f(), empty_f(), empty_f(), empty_f()};
  }
  fa_b200_bwd_args a = {};
  a.dout = dout.data_ptr(); a.q = q.data_ptr(); a.k = k.data_ptr(); a.v = v.data_ptr(); a.out = out.data_ptr();
  a.softmax_lse = softmax_lse.data_ptr<float>();
  a.dq = dq.data_ptr(); a.dk = dk.data_ptr(); a.dv = dv.data_ptr();
  a.softmax_d = softmax_d.data_ptr<float>();
  const int o = varlen ? 0 : 1;   // index of the row dimension
#define FA_SET3(name, t) a.name##_batch_stride = varlen ? 0 : (t).stride(0); a.name##_row_stride = (t).stride(o); a.name##_head_stride = (t).stride(o + 1)
  FA_SET3(q, q); FA_SET3(k, k); FA_SET3(v, v); FA_SET3(o, out); FA_SET3(do, dout); FA_SET3(dq, dq); FA_SET3(dk, dk); FA_SET3(dv, dv);
#undef FA_SET3
  a.b = int(b); a.seqlen_q = int(sq); a.seqlen_k = int(sk); a.h = int(h); a.h_k = int(hk); a.d = int(d);
  if (varlen) {
    a.total_q = int(total_q); a.total_k = int(total_k);
    a.cu_seqlens_q = cu_seqlens_q_.value().data_ptr<int32_t>();
    a.cu_seqlens_k = cu_seqlens_k_.value().data_ptr<int32_t>();
  }
  a.dtype = dtype_code(q);
  a.softmax_scale = float(softmax_scale); a.softcap = float(softcap);
  a.is_causal = is_causal; a.window_size_left = int(window_size_left); a.window_size_right = int(window_size_right);
  a.deterministic = deterministic;
  a.sm_margin = int(sm_margin);
  const size_t nbytes = fa_b200_bwd_workspace_bytes(&a);
  TORCH_CHECK(nbytes > 0, fa_b200_last_error());
  auto ws = at::empty({int64_t(nbytes)}, q.options().dtype(at::kByte));
  a.workspace = ws.data_ptr();
  a.workspace_bytes = nbytes;
  check_rc(varlen ? fa_b200_mha_varlen_bwd(&a, cur_stream()) : fa_b200_mha_bwd(&a, cur_stream()));
  return {softmax_d, empty_f(), empty_f(), empty_f(), empty_f()};
}

// mha_combine (hopper/flash_api.cpp:1570-1670): out_partial (splits, b, s, h, d) fp32, lse_partial (splits, b, s, h) fp32
// -> (out (b, s, h, d), lse (b, h, s)).  Any leading layout works for this library's kernel: rows are (…) flattened.
std::tuple<at::Tensor, at::Tensor> fa3_combine(at::Tensor out_partial, at::Tensor lse_partial, std::optional<at::Tensor> out_,
                                               std::optional<at::ScalarType> out_dtype_) {
  TORCH_CHECK(out_partial.is_cuda() && lse_partial.is_cuda(), "Input tensor must be on CUDA device");
  TORCH_CHECK(out_partial.dtype() == at::kFloat, "FlashAttention combine only support fp32 data type");
  TORCH_CHECK(lse_partial.dtype() == at::kFloat, "lse_partial must have dtype float32");
  TORCH_CHECK(out_partial.dim() >= 2 && lse_partial.dim() == out_partial.dim() - 1, "out_partial (splits, ..., d), lse_partial (splits, ...)");
  c10::cuda::CUDAGuard guard(out_partial.device());
  auto op = out_partial.contiguous();
  auto lp = lse_partial.contiguous();
  const int64_t splits = op.size(0), d = op.size(-1);
  TORCH_CHECK(splits <= 256, "FlashAttention combine only supports num_splits at most 256");
  TORCH_CHECK(lp.size(0) == splits, "lse_partial has the wrong number of splits");
  const int64_t rows = lp.numel() / std::max<int64_t>(splits, 1);
  const at::ScalarType dt = out_.has_value() ? out_.value().scalar_type() : out_dtype_.value_or(at::kBFloat16);
  TORCH_CHECK(dt == at::kBFloat16 || dt == at::kHalf, "fa_b200: combine output dtype must be fp16 or bf16");
  auto shape = op.sizes().slice(1).vec();
  at::Tensor out = out_.has_value() ? out_.value() : at::empty(shape, op.options().dtype(dt));
  TORCH_CHECK(out.is_contiguous() && out.sizes() == at::IntArrayRef(shape), "out has the wrong shape / is not contiguous");
  at::Tensor lse = at::empty(lp.sizes().slice(1), lp.options());
  check_rc(fa_b200_fwd_combine(op.data_ptr<float>(), lp.data_ptr<float>(), out.data_ptr(), lse.data_ptr<float>(), int32_t(splits),
                               rows, int32_t(d), dt == at::kBFloat16 ? FA_B200_BF16 : FA_B200_FP16, cur_stream()));
  if (lse.dim() == 3) lse = lse.transpose(1, 2);   // (b, s, h) -> (b, h, s), the reference's returned layout
  return {out, lse};
}

// mha_fwd_get_scheduler_metadata (hopper/flash_api.cpp:521-670).  The plan this library will use for the shape, as an
// int32 tensor on seqused_k's device: [num_splits, gqa_pack_factor, num_m_blocks, tiles_per_cta, num_sms_used].
at::Tensor fa3_get_scheduler_metadata(int64_t batch_size, int64_t max_seqlen_q, int64_t max_seqlen_k, int64_t num_heads,
                                      int64_t num_heads_k, int64_t headdim, int64_t headdim_v, at::ScalarType qkv_dtype,
                                      at::Tensor seqused_k, std::optional<at::Tensor> cu_seqlens_q_,
                                      std::optional<at::Tensor> cu_seqlens_k_, std::optional<at::Tensor> cu_seqlens_k_new_,
                                      std::optional<at::Tensor> seqused_q_, std::optional<at::Tensor> leftpad_k_,
                                      std::optional<int64_t> page_size, int64_t max_seqlen_k_new, bool is_causal,
                                      int64_t window_size_left, int64_t window_size_right, int64_t attention_chunk,
                                      bool has_softcap, int64_t num_splits, std::optional<bool> pack_gqa_, int64_t sm_margin) {
  TORCH_CHECK(qkv_dtype == at::kHalf || qkv_dtype == at::kBFloat16 || qkv_dtype == at::kFloat8_e4m3fn,
              "FlashAttention only supports fp16, bf16, and fp8_e4m3 data type");
  TORCH_CHECK(num_heads % num_heads_k == 0, "Number of heads in key/value must divide number of heads in query");
  TORCH_CHECK(headdim == headdim_v, "fa_b200: V head dimension must equal the Q/K head dimension");
  unsupported(attention_chunk != 0, "attention_chunk");
  fa_b200_fwd_args a = {};
  a.b = int(batch_size); a.seqlen_q = int(max_seqlen_q); a.seqlen_k = int(max_seqlen_k + max_seqlen_k_new);
  a.h = int(num_heads); a.h_k = int(num_heads_k); a.d = int(headdim);
  a.dtype = qkv_dtype == at::kHalf ? FA_B200_FP16 : (qkv_dtype == at::kBFloat16 ? FA_B200_BF16 : FA_B200_FP8_E4M3);
  a.entry = cu_seqlens_q_.has_value() ? FA_B200_ENTRY_VARLEN : FA_B200_ENTRY_KVCACHE;
  if (cu_seqlens_q_.has_value()) { a.cu_seqlens_q = reinterpret_cast<const int32_t*>(16); a.cu_seqlens_k = reinterpret_cast<const int32_t*>(16); }
  a.is_causal = is_causal; a.window_size_left = int(window_size_left); a.window_size_right = int(window_size_right);
  a.softcap = has_softcap ? 1.f : 0.f;
  a.num_splits = int(num_splits);
  a.pack_gqa = pack_code(pack_gqa_);
  a.sm_margin = int(sm_margin);
  if (page_size.has_value()) {
    a.block_table = reinterpret_cast<const int32_t*>(16);   // marks the paged mode; never dereferenced by the planner
    a.page_block_size = int(page_size.value());
    a.num_pages = 1;
  }
  int32_t plan[3];
  check_rc(fa_b200_fwd_plan(&a, plan));
  auto host = at::empty({3}, at::TensorOptions().dtype(at::kInt));
  for (int i = 0; i < 3; ++i) host.data_ptr<int32_t>()[i] = plan[i];
  return host.to(seqused_k.device());
}

}  // namespace fa_ext

TORCH_LIBRARY(flash_attn_3, m) {
  m.def("fwd("
        "Tensor q, Tensor k, Tensor v, Tensor(k_new!)? k_new = None, Tensor(v_new!)? v_new = None, Tensor? q_v = None,"
        "Tensor(out!)? out = None, Tensor? cu_seqlens_q = None, Tensor? cu_seqlens_k = None, Tensor? cu_seqlens_k_new = None,"
        "Tensor? seqused_q = None, Tensor? seqused_k = None, int? max_seqlen_q = None, int? max_seqlen_k = None,"
        "Tensor? page_table = None, Tensor? kv_batch_idx = None, Tensor? leftpad_k = None, Tensor? rotary_cos = None,"
        "Tensor? rotary_sin = None, Tensor? seqlens_rotary = None, Tensor? q_descale = None, Tensor? k_descale = None,"
        "Tensor? v_descale = None, float? softmax_scale = None, bool is_causal = False, int window_size_left = -1,"
        "int window_size_right = -1, int attention_chunk = 0, float softcap = 0.0, bool is_rotary_interleaved = False,"
        "Tensor? scheduler_metadata = None, int num_splits = 0, bool? pack_gqa = None, int sm_margin = 0)"
        " -> (Tensor(out!), Tensor, Tensor, Tensor)");
  m.def("bwd("
        "Tensor dout, Tensor q, Tensor k, Tensor v, Tensor out, Tensor softmax_lse, Tensor(dq!)? dq = None,"
        "Tensor(dk!)? dk = None, Tensor(dv!)? dv = None, Tensor? cu_seqlens_q = None, Tensor? cu_seqlens_k = None,"
        "Tensor? seqused_q = None, Tensor? seqused_k = None, int? max_seqlen_q = None, int? max_seqlen_k = None,"
        "float? softmax_scale = None, bool is_causal = False, int window_size_left = -1, int window_size_right = -1,"
        "float softcap = 0.0, bool deterministic = False, int sm_margin = 0) -> (Tensor, Tensor, Tensor, Tensor, Tensor)");
  m.def("fwd_combine(Tensor out_partial, Tensor lse_partial, Tensor(out!)? out = None, ScalarType? out_dtype = None)"
        " -> (Tensor(out!), Tensor)");
  m.def("get_scheduler_metadata("
        "int batch_size, int max_seqlen_q, int max_seqlen_k, int num_heads, int num_heads_k, int headdim, int headdim_v,"
        "ScalarType qkv_dtype, Tensor seqused_k, Tensor? cu_seqlens_q = None, Tensor? cu_seqlens_k = None,"
        "Tensor? cu_seqlens_k_new = None, Tensor? seqused_q = None, Tensor? leftpad_k = None, int? page_size = None,"
        "int max_seqlen_k_new = 0, bool is_causal = False, int window_size_left = -1, int window_size_right = -1,"
        "int attention_chunk = 0, bool has_softcap = False, int num_splits = 0, bool? pack_gqa = None, int sm_margin = 0)"
        " -> Tensor");
}

TORCH_LIBRARY_IMPL(flash_attn_3, CUDA, m) {
  m.impl("fwd", &fa_ext::fa3_fwd);
  m.impl("bwd", &fa_ext::fa3_bwd);
  m.impl("fwd_combine", &fa_ext::fa3_combine);
  m.impl("get_scheduler_metadata", &fa_ext::fa3_get_scheduler_metadata);
}
