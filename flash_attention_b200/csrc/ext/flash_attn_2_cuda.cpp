// flash_attn_2_cuda.cpp — the compiled operator module `flash_attn_2_cuda` (pybind11) on top of libfa_b200.so.
//
// Exact surface of the reference's extension (csrc/flash_attn/flash_api.cpp:1535-1542): fwd, varlen_fwd, bwd,
// varlen_bwd, fwd_kvcache -- positional arguments, list-of-tensors results, TORCH_CHECK -> RuntimeError with the
// reference's message texts -- so that flash_attn/flash_attn_interface.py (2.8.3 and 2.8.4) imports it unchanged.
// Everything the reference's host entries do before the launch lives here in C++ (argument checks, output / scratch
// allocation from the caching allocator, shape tricks, dropout Philox state: flash_api.cpp:368-536, 538-788, 800-1008,
// 1010-1241, 1243-1532); the compute is one call into the C ABI (include/fa_b200.h) on the current CUDA stream.
// Extras beyond FA2's five (same module, used by this package's own API and the FA3-style library in
// flash_attn_3_ops.cpp): fwd_fp8, fwd_ex, fwd_combine, plan_kvcache.
//
// There is no fallback: a failing C-ABI call raises with the library's error text.
#include <ATen/cuda/CUDAContext.h>
#include <ATen/cuda/CUDAGeneratorImpl.h>
#include <ATen/cuda/CUDAGraphsUtils.cuh>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <array>
#include <mutex>
#include <vector>

#include "fa_torch_common.h"

namespace fa_ext {

// ------------------------------------------------------------------------------------------------ helpers
// Philox (seed, offset) of this call from the default CUDA generator, advanced like the reference does
// (flash_api.cpp:505-517: counter_offset = b * h * 32); returns the `rng_state` tensor the reference returns.
at::Tensor dropout_state(float p_dropout, int64_t b, int64_t h, const at::Device& dev, uint64_t* seed, uint64_t* offset) {
  TORCH_CHECK(p_dropout >= 0.f && p_dropout < 1.f, "p_dropout must be in [0, 1)");
  *seed = *offset = 0;
  // p_dropout == 0: a fresh, uninitialised (2,) int64 tensor, exactly as the reference returns (flash_api.cpp:501-503); it
  // must be a new tensor per call: the reference's custom ops (and ours) may not return aliases of earlier outputs
  if (p_dropout == 0.f) return at::empty({2}, at::TensorOptions().dtype(at::kLong).device(dev));
  auto gen = at::get_generator_or_default<at::CUDAGeneratorImpl>(std::nullopt, at::cuda::detail::getDefaultCUDAGenerator());
  at::PhiloxCudaState st;
  {
    std::lock_guard<std::mutex> lock(gen->mutex_);
    st = gen->philox_cuda_state(uint64_t(b * h * 32));
  }
  TORCH_CHECK(!st.captured_, "fa_b200: attention dropout under CUDA graph capture is not supported");
  *seed = st.seed_.val;
  *offset = st.offset_.val;
  auto rng = at::empty({2}, at::TensorOptions().dtype(at::kLong).device(dev));
  check_rc(fa_b200_write_rng_state(reinterpret_cast<uint64_t*>(rng.data_ptr()), *seed, *offset, cur_stream()));
  return rng;
}

void set_alibi(fa_b200_fwd_args& a, const std::optional<at::Tensor>& alibi, int64_t b, int64_t h) {
  if (!alibi.has_value()) return;
  const auto& t = alibi.value();
  TORCH_CHECK(t.dtype() == at::kFloat, "ALiBi slopes must have dtype fp32");
  TORCH_CHECK(t.is_cuda(), "Input tensor must be on CUDA device");
  TORCH_CHECK(t.stride(-1) == 1, "ALiBi slopes tensor must have contiguous last dimension");
  TORCH_CHECK(t.sizes() == at::IntArrayRef({h}) || t.sizes() == at::IntArrayRef({b, h}),
              "ALiBi slopes must have shape (nheads,) or (batch, nheads)");
  a.alibi_slopes = t.data_ptr<float>();
  a.alibi_slopes_batch_stride = t.dim() == 2 ? t.stride(0) : 0;
}

void common_checks(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v) {   // order as in flash_api.cpp:393-407
  TORCH_CHECK(q.dtype() == at::kHalf || q.dtype() == at::kBFloat16, "FlashAttention only support fp16 and bf16 data type");
  TORCH_CHECK(k.dtype() == q.dtype(), "query and key must have the same dtype");
  TORCH_CHECK(v.dtype() == q.dtype(), "query and value must have the same dtype");
  TORCH_CHECK(q.is_cuda() && k.is_cuda() && v.is_cuda(), "Input tensor must be on CUDA device");
  TORCH_CHECK(q.stride(-1) == 1, "Input tensor must have contiguous last dimension");
  TORCH_CHECK(k.stride(-1) == 1, "Input tensor must have contiguous last dimension");
  TORCH_CHECK(v.stride(-1) == 1, "Input tensor must have contiguous last dimension");
}

void check_generator_slot(const py::object& gen) {
  TORCH_CHECK(gen.is_none(), "flash-attn: the RNG `generator` argument is no longer supported and must be None; "
                             "dropout (when enabled) uses the default CUDA generator.");   // flash_api.cpp:384-386
}

inline void set_qkvo_dense(fa_b200_fwd_args& a, const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, const at::Tensor& out) {
  a.q = q.data_ptr(); a.k = k.data_ptr(); a.v = v.data_ptr(); a.out = out.data_ptr();
  a.q_batch_stride = q.stride(0); a.q_row_stride = q.stride(1); a.q_head_stride = q.stride(2);
  a.k_batch_stride = k.stride(0); a.k_row_stride = k.stride(1); a.k_head_stride = k.stride(2);
  a.v_batch_stride = v.stride(0); a.v_row_stride = v.stride(1); a.v_head_stride = v.stride(2);
  a.o_batch_stride = out.stride(0); a.o_row_stride = out.stride(1); a.o_head_stride = out.stride(2);
}

// ------------------------------------------------------------------------------------------------ fwd
// mha_fwd (flash_api.cpp:368-536): q (b,sq,h,d), k/v (b,sk,h_k,d) -> [out, softmax_lse (b,h,sq), p, rng_state]
std::vector<at::Tensor> fwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, std::optional<at::Tensor> out_,
                            std::optional<at::Tensor> alibi_slopes_, double p_dropout, double softmax_scale, bool is_causal,
                            int64_t window_size_left, int64_t window_size_right, double softcap, bool return_softmax,
                            const py::object& unused_generator_compat) {
  check_generator_slot(unused_generator_compat);
  common_checks(q, k, v);
  c10::cuda::CUDAGuard guard(q.device());
  if (return_softmax) TORCH_CHECK(p_dropout > 0.0, "return_softmax is only supported when p_dropout > 0.0");
  TORCH_CHECK(q.dim() == 4 && k.dim() == 4 && v.dim() == 4, "q, k, v must be 4-D (batch, seqlen, nheads, headdim)");
  const int64_t b = q.size(0), sq = q.size(1), h = q.size(2), d = q.size(3), sk = k.size(1), hk = k.size(2);
  TORCH_CHECK(b > 0, "batch size must be positive");
  TORCH_CHECK(d <= 256, "FlashAttention forward only supports head dimension at most 256");
  TORCH_CHECK(d % 8 == 0, "query, key, value, and out_ must have a head_size that is a multiple of 8");
  TORCH_CHECK(h % hk == 0, "Number of heads in key/value must divide number of heads in query");
  check_shape(k, {b, sk, hk, d}, "k");
  check_shape(v, {b, sk, hk, d}, "v");
  if (softcap > 0.0) TORCH_CHECK(p_dropout == 0.0, "Softcapping does not support dropout for now");
  at::Tensor out;
  if (out_.has_value()) {
    out = out_.value();
    TORCH_CHECK(out.dtype() == q.dtype(), "Output must have the same dtype as inputs");
    TORCH_CHECK(out.is_cuda(), "Input tensor must be on CUDA device");
    TORCH_CHECK(out.stride(-1) == 1, "Output tensor must have contiguous last dimension");
    check_shape(out, {b, sq, h, d}, "out");
  } else {
    out = at::empty_like(q);
  }
  auto lse = at::empty({b, h, sq}, q.options().dtype(at::kFloat));
  at::Tensor p = empty_like_dtype(q);
  fa_b200_fwd_args a = {};
  auto rng_state = dropout_state(float(p_dropout), b, h, q.device(), &a.philox_seed, &a.philox_offset);
  a.p_dropout = float(p_dropout);
  if (sq == 0) return {out, lse, p, rng_state};
  if (sk == 0) {  // flash_api.cpp:521-528
    out.zero_();
    lse.fill_(std::numeric_limits<float>::infinity());
    return {out, lse, p, rng_state};
  }
  if (return_softmax) {  // (b, h, seqlen_q_rounded, seqlen_k_rounded), flash_api.cpp:466-470
    p = at::empty({b, h, round128(sq), round128(sk)}, q.options());
    a.s_dmask = p.data_ptr();
  }
  set_qkvo_dense(a, q, k, v, out);
  a.softmax_lse = lse.data_ptr<float>();
  a.b = int(b); a.seqlen_q = int(sq); a.seqlen_k = int(sk); a.h = int(h); a.h_k = int(hk); a.d = int(d);
  a.dtype = dtype_code(q);
  a.softmax_scale = float(softmax_scale); a.softcap = float(softcap);
  a.is_causal = is_causal; a.window_size_left = int(window_size_left); a.window_size_right = int(window_size_right);
  a.num_splits = 0;
  a.entry = FA_B200_ENTRY_DENSE;
  set_alibi(a, alibi_slopes_, b, h);
  auto ws = fwd_workspace(a, q);
  check_rc(fa_b200_mha_fwd(&a, cur_stream()));
  return {out, lse, p, rng_state};
}

// ------------------------------------------------------------------------------------------------ fwd_fp8 / fwd_ex
// FP8 e4m3 forward (no FA2 equivalent: flash_api.cpp:396 rejects fp8; spec = FA3 mha_fwd descales, hopper/flash_api.cpp,
// flash_attn/cute/interface.py:637-643): q/k/v float8_e4m3fn, descales (b, h_k) fp32 or None; bf16 out -> [out, lse]
std::vector<at::Tensor> fwd_fp8(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, std::optional<at::Tensor> out_,
                                std::optional<at::Tensor> q_descale, std::optional<at::Tensor> k_descale,
                                std::optional<at::Tensor> v_descale, double softmax_scale, bool is_causal,
                                int64_t window_size_left, int64_t window_size_right, double softcap,
                                std::optional<at::Tensor> alibi_slopes_, int64_t num_splits, std::optional<bool> pack_gqa) {
  TORCH_CHECK(q.is_cuda() && k.is_cuda() && v.is_cuda(), "Input tensor must be on CUDA device");
  TORCH_CHECK(q.dtype() == at::kFloat8_e4m3fn && k.dtype() == at::kFloat8_e4m3fn && v.dtype() == at::kFloat8_e4m3fn,
              "fwd_fp8 expects float8_e4m3fn q, k, v");
  TORCH_CHECK(q.stride(-1) == 1 && k.stride(-1) == 1 && v.stride(-1) == 1, "Input tensor must have contiguous last dimension");
  c10::cuda::CUDAGuard guard(q.device());
  const int64_t b = q.size(0), sq = q.size(1), h = q.size(2), d = q.size(3), sk = k.size(1), hk = k.size(2);
  TORCH_CHECK(d % 16 == 0 && d <= 128, "fp8 forward supports head dimensions that are multiples of 16 up to 128");
  TORCH_CHECK(h % hk == 0, "Number of heads in key/value must divide number of heads in query");
  at::Tensor out = out_.has_value() ? out_.value() : at::empty({b, sq, h, d}, q.options().dtype(at::kBFloat16));
  TORCH_CHECK(out.dtype() == at::kBFloat16 && out.stride(-1) == 1, "out must be bf16 (b,sq,h,d)");
  check_shape(out, {b, sq, h, d}, "out");
  auto lse = at::empty({b, h, sq}, q.options().dtype(at::kFloat));
  if (sq == 0) return {out, lse};
  if (sk == 0) {
    out.zero_();
    lse.fill_(std::numeric_limits<float>::infinity());
    return {out, lse};
  }
  fa_b200_fwd_args a = {};
  set_qkvo_dense(a, q, k, v, out);
  a.softmax_lse = lse.data_ptr<float>();
  a.b = int(b); a.seqlen_q = int(sq); a.seqlen_k = int(sk); a.h = int(h); a.h_k = int(hk); a.d = int(d);
  a.dtype = FA_B200_FP8_E4M3;
  std::vector<at::Tensor> keep;
  set_descales(a, q_descale, k_descale, v_descale, b, hk, keep);
  a.softmax_scale = float(softmax_scale); a.softcap = float(softcap);
  a.is_causal = is_causal; a.window_size_left = int(window_size_left); a.window_size_right = int(window_size_right);
  a.num_splits = int(num_splits);
  a.pack_gqa = pack_code(pack_gqa);
  a.entry = FA_B200_ENTRY_DENSE;
  set_alibi(a, alibi_slopes_, b, h);
  auto ws = fwd_workspace(a, q);
  check_rc(fa_b200_mha_fwd(&a, cur_stream()));
  return {out, lse};
}

// dense forward with the FA3 scheduling knobs (`num_splits`, `pack_gqa`; hopper/flash_api.cpp mha_fwd) -> [out, lse]
std::vector<at::Tensor> fwd_ex(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, double softmax_scale, bool is_causal,
                               int64_t window_size_left, int64_t window_size_right, double softcap, int64_t num_splits,
                               std::optional<bool> pack_gqa, std::optional<at::Tensor> alibi_slopes_) {
  common_checks(q, k, v);
  c10::cuda::CUDAGuard guard(q.device());
  const int64_t b = q.size(0), sq = q.size(1), h = q.size(2), d = q.size(3), sk = k.size(1), hk = k.size(2);
  TORCH_CHECK(d % 8 == 0 && h % hk == 0, "bad head_size / number of heads");
  auto out = at::empty_like(q);
  auto lse = at::empty({b, h, sq}, q.options().dtype(at::kFloat));
  fa_b200_fwd_args a = {};
  set_qkvo_dense(a, q, k, v, out);
  a.softmax_lse = lse.data_ptr<float>();
  a.b = int(b); a.seqlen_q = int(sq); a.seqlen_k = int(sk); a.h = int(h); a.h_k = int(hk); a.d = int(d);
  a.dtype = dtype_code(q);
  a.softmax_scale = float(softmax_scale); a.softcap = float(softcap);
  a.is_causal = is_causal; a.window_size_left = int(window_size_left); a.window_size_right = int(window_size_right);
  a.num_splits = int(num_splits);
  a.pack_gqa = pack_code(pack_gqa);
  a.entry = FA_B200_ENTRY_DENSE;
  set_alibi(a, alibi_slopes_, b, h);
  auto ws = fwd_workspace(a, q);
  check_rc(fa_b200_mha_fwd(&a, cur_stream()));
  return {out, lse};
}

// host-only: (num_splits, gqa_pack_factor, num_m_blocks) the library will use for this decode shape
std::tuple<int64_t, int64_t, int64_t> plan_kvcache(int64_t batch, int64_t seqlen_q, int64_t seqlen_k, int64_t h, int64_t hk,
                                                   int64_t d, int64_t page_size, bool causal, int64_t wl, int64_t wr,
                                                   int64_t num_splits, std::optional<bool> pack_gqa) {
  fa_b200_fwd_args a = {};
  a.b = int(batch); a.seqlen_q = int(seqlen_q); a.seqlen_k = int(seqlen_k); a.h = int(h); a.h_k = int(hk); a.d = int(d);
  a.dtype = FA_B200_BF16;
  a.entry = FA_B200_ENTRY_KVCACHE;
  a.is_causal = causal; a.window_size_left = int(wl); a.window_size_right = int(wr);
  a.num_splits = int(num_splits);
  a.pack_gqa = pack_code(pack_gqa);
  if (page_size) {
    a.block_table = reinterpret_cast<const int32_t*>(16);  // non-null marks the paged mode; never dereferenced by the planner
    a.page_block_size = int(page_size);
    a.num_pages = 1;
  }
  int32_t out[3];
  check_rc(fa_b200_fwd_plan(&a, out));
  return {out[0], out[1], out[2]};
}

// ------------------------------------------------------------------------------------------------ varlen_fwd
// mha_varlen_fwd (flash_api.cpp:538-788): q (total_q,h,d); k/v (total_k,h_k,d) or paged (num_blocks,page,h_k,d) with
// block_table (b,max_blocks) -> [out, softmax_lse (h,total_q), p, rng_state]
std::vector<at::Tensor> varlen_fwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, std::optional<at::Tensor> out_,
                                   const at::Tensor& cu_seqlens_q, const at::Tensor& cu_seqlens_k,
                                   std::optional<at::Tensor> seqused_k, std::optional<at::Tensor> leftpad_k_,
                                   std::optional<at::Tensor> block_table_, std::optional<at::Tensor> alibi_slopes_,
                                   int64_t max_seqlen_q, int64_t max_seqlen_k, double p_dropout, double softmax_scale,
                                   bool zero_tensors, bool is_causal, int64_t window_size_left, int64_t window_size_right,
                                   double softcap, bool return_softmax, const py::object& unused_generator_compat,
                                   int64_t num_splits, std::optional<at::Tensor> seqused_q) {
  check_generator_slot(unused_generator_compat);
  common_checks(q, k, v);
  c10::cuda::CUDAGuard guard(q.device());
  if (return_softmax) TORCH_CHECK(p_dropout > 0.0, "return_softmax is only supported when p_dropout > 0.0");
  TORCH_CHECK(cu_seqlens_q.dtype() == at::kInt, "cu_seqlens_q must have dtype int32");
  TORCH_CHECK(cu_seqlens_k.dtype() == at::kInt, "cu_seqlens_k must have dtype int32");
  TORCH_CHECK(cu_seqlens_q.is_cuda() && cu_seqlens_k.is_cuda(), "cu_seqlens must be on CUDA device");
  TORCH_CHECK(cu_seqlens_q.is_contiguous() && cu_seqlens_k.is_contiguous(), "cu_seqlens must be contiguous");
  const bool paged = block_table_.has_value();
  TORCH_CHECK(q.dim() == 3, "q must be (total_q, nheads, headdim)");
  const int64_t total_q = q.size(0), h = q.size(1), d = q.size(2);
  const int64_t b = cu_seqlens_q.numel() - 1;
  TORCH_CHECK(b > 0, "batch size must be positive");
  TORCH_CHECK(d <= 256, "FlashAttention forward only supports head dimension at most 256");
  TORCH_CHECK(d % 8 == 0, "query, key, value, and out_ must have a head_size that is a multiple of 8");
  check_shape(cu_seqlens_k, {b + 1}, "cu_seqlens_k");
  int64_t total_k = 0, hk = 0, page = 0, num_pages = 0;
  if (paged) {
    const auto& bt = block_table_.value();
    TORCH_CHECK(bt.is_cuda(), "Input tensor must be on CUDA device");
    TORCH_CHECK(bt.dtype() == at::kInt, "block_table must have dtype torch.int32");
    TORCH_CHECK(bt.stride(-1) == 1, "block_table must have contiguous last dimension");
    TORCH_CHECK(k.dim() == 4, "paged k must be (num_blocks, page_block_size, nheads_k, headdim)");
    num_pages = k.size(0); page = k.size(1); hk = k.size(2);
    TORCH_CHECK(page % 256 == 0, "Paged KV cache block size must be divisible by 256");
    TORCH_CHECK(num_splits <= 1, "num_splits > 1 is not supported for varlen paged KV");   // flash_api.cpp:736
    check_shape(bt, {b, bt.size(1)}, "block_table");
    check_shape(v, {num_pages, page, hk, d}, "v");
  } else {
    TORCH_CHECK(k.dim() == 3, "k must be (total_k, nheads_k, headdim)");
    total_k = k.size(0); hk = k.size(1);
    check_shape(v, {total_k, hk, d}, "v");
  }
  TORCH_CHECK(h % hk == 0, "Number of heads in key/value must divide number of heads in query");
  if (softcap > 0.0) TORCH_CHECK(p_dropout == 0.0, "Softcapping does not support dropout for now");
  at::Tensor out;
  if (out_.has_value()) {
    out = out_.value();
    TORCH_CHECK(out.dtype() == q.dtype(), "Output must have the same dtype as inputs");
    TORCH_CHECK(out.is_cuda() && out.stride(-1) == 1, "Output tensor must have contiguous last dimension");
    check_shape(out, {total_q, h, d}, "out");
  } else {
    out = at::empty_like(q);
  }
  auto lse = at::empty({h, total_q}, q.options().dtype(at::kFloat));
  at::Tensor p = empty_like_dtype(q);
  fa_b200_fwd_args a = {};
  if (p_dropout > 0.0)
    TORCH_CHECK(!paged && !seqused_k.has_value() && !seqused_q.has_value(), "dropout is not supported together with paged KV / seqused");
  auto rng_state = dropout_state(float(p_dropout), b, h, q.device(), &a.philox_seed, &a.philox_offset);
  a.p_dropout = float(p_dropout);
  if (zero_tensors) {
    out.zero_();
    lse.fill_(-std::numeric_limits<float>::infinity());
  }
  if (total_q == 0) return {out, lse, p, rng_state};
  if (max_seqlen_k == 0) {
    out.zero_();
    lse.fill_(std::numeric_limits<float>::infinity());
    return {out, lse, p, rng_state};
  }
  if (return_softmax) {  // flash_api.cpp:690-694
    p = at::zeros({b, h, round128(max_seqlen_q), round128(max_seqlen_k)}, q.options());
    a.s_dmask = p.data_ptr();
  }
  a.q = q.data_ptr(); a.k = k.data_ptr(); a.v = v.data_ptr(); a.out = out.data_ptr();
  a.softmax_lse = lse.data_ptr<float>();
  a.q_batch_stride = 0; a.q_row_stride = q.stride(0); a.q_head_stride = q.stride(1);
  a.o_batch_stride = 0; a.o_row_stride = out.stride(0); a.o_head_stride = out.stride(1);
  if (paged) {
    const auto& bt = block_table_.value();
    a.k_batch_stride = k.stride(0); a.k_row_stride = k.stride(1); a.k_head_stride = k.stride(2);
    a.v_batch_stride = v.stride(0); a.v_row_stride = v.stride(1); a.v_head_stride = v.stride(2);
    a.block_table = bt.data_ptr<int32_t>();
    a.block_table_batch_stride = bt.stride(0);
    a.page_block_size = int(page); a.num_pages = int(num_pages);
  } else {
    a.k_batch_stride = 0; a.k_row_stride = k.stride(0); a.k_head_stride = k.stride(1);
    a.v_batch_stride = 0; a.v_row_stride = v.stride(0); a.v_head_stride = v.stride(1);
  }
  a.b = int(b); a.seqlen_q = int(max_seqlen_q); a.seqlen_k = int(max_seqlen_k); a.h = int(h); a.h_k = int(hk); a.d = int(d);
  a.total_q = int(total_q); a.total_k = int(total_k);
  a.dtype = dtype_code(q);
  a.cu_seqlens_q = cu_seqlens_q.data_ptr<int32_t>();
  a.cu_seqlens_k = cu_seqlens_k.data_ptr<int32_t>();
  a.seqused_k = int_vec(seqused_k, b, "seqused_k");
  a.seqused_q = int_vec(seqused_q, b, "seqused_q");
  if (leftpad_k_.has_value()) {  // flash_api.cpp:740-747
    TORCH_CHECK(!paged, "We don't support Paged KV and leftpad_k running at the same time yet");
    TORCH_CHECK(leftpad_k_.value().dtype() == at::kInt, "leftpad_k must have dtype int32");
    a.leftpad_k = int_vec(leftpad_k_, b, "leftpad_k");
  }
  a.softmax_scale = float(softmax_scale); a.softcap = float(softcap);
  a.is_causal = is_causal; a.window_size_left = int(window_size_left); a.window_size_right = int(window_size_right);
  a.num_splits = 1;
  a.entry = FA_B200_ENTRY_VARLEN;
  set_alibi(a, alibi_slopes_, b, h);
  auto ws = fwd_workspace(a, q);   // the (b + 1)-entry m-block prefix sum of the persistent kernel's work list
  check_rc(fa_b200_mha_varlen_fwd(&a, cur_stream()));
  return {out, lse, p, rng_state};
}

// ------------------------------------------------------------------------------------------------ fwd_kvcache
// mha_fwd_kvcache (flash_api.cpp:1243-1532): q (b,sq,h,d); kcache/vcache (b_c,sk,h_k,d) or paged
// (num_blocks,page,h_k,d) -> [out, softmax_lse (b,h,sq)]; appends k_/v_ into the cache in place (+ rotary)
std::vector<at::Tensor> fwd_kvcache(const at::Tensor& q_in, const at::Tensor& kcache, const at::Tensor& vcache,
                                    std::optional<at::Tensor> k_, std::optional<at::Tensor> v_,
                                    std::optional<at::Tensor> seqlens_k_, std::optional<at::Tensor> rotary_cos_,
                                    std::optional<at::Tensor> rotary_sin_, std::optional<at::Tensor> cache_batch_idx_,
                                    std::optional<at::Tensor> leftpad_k_, std::optional<at::Tensor> block_table_,
                                    std::optional<at::Tensor> alibi_slopes_, std::optional<at::Tensor> out_,
                                    double softmax_scale, bool is_causal, int64_t window_size_left, int64_t window_size_right,
                                    double softcap, bool is_rotary_interleaved, int64_t num_splits) {
  common_checks(q_in, kcache, vcache);
  c10::cuda::CUDAGuard guard(q_in.device());
  at::Tensor q = q_in;
  TORCH_CHECK(q.dim() == 4, "q must be (batch, seqlen_q, nheads, headdim)");
  const int64_t b = q.size(0), sq = q.size(1), h = q.size(2), d = q.size(3);
  const bool paged = block_table_.has_value();
  int64_t sk, hk, page = 0, num_pages = 0, max_blocks = 0;
  if (paged) {
    const auto& bt = block_table_.value();
    TORCH_CHECK(!cache_batch_idx_.has_value(), "Paged KVcache does not support cache_batch_idx");
    TORCH_CHECK(bt.is_cuda(), "Input tensor must be on CUDA device");
    TORCH_CHECK(bt.dtype() == at::kInt, "block_table must have dtype torch.int32");
    TORCH_CHECK(bt.stride(-1) == 1, "block_table must have contiguous last dimension");
    num_pages = kcache.size(0); page = kcache.size(1); hk = kcache.size(2);
    TORCH_CHECK(page % 256 == 0, "Paged KV cache block size must be divisible by 256");
    max_blocks = bt.size(1);
    sk = max_blocks * page;
    check_shape(bt, {b, max_blocks}, "block_table");
    check_shape(vcache, {num_pages, page, hk, d}, "vcache");
  } else {
    TORCH_CHECK(cache_batch_idx_.has_value() || kcache.size(0) == b, "kcache batch must equal q batch unless cache_batch_idx is given");
    sk = kcache.size(1); hk = kcache.size(2);
    check_shape(vcache, {kcache.size(0), sk, hk, d}, "vcache");
  }
  TORCH_CHECK(b > 0, "batch size must be positive");
  TORCH_CHECK(d <= 256, "FlashAttention forward only supports head dimension at most 256");
  TORCH_CHECK(d % 8 == 0, "fa_b200: fwd_kvcache needs a head_size that is a multiple of 8 (the cache is used in place)");
  TORCH_CHECK(h % hk == 0, "Number of heads in key/value must divide number of heads in query");
  const int64_t s_new = k_.has_value() ? k_.value().size(1) : 0;
  if (seqlens_k_.has_value()) {
    const auto& sl = seqlens_k_.value();
    TORCH_CHECK(sl.dtype() == at::kInt, "seqlens_k must have dtype int32");
    TORCH_CHECK(sl.is_cuda(), "Input tensor must be on CUDA device");
    TORCH_CHECK(sl.is_contiguous(), "seqlens_k must be contiguous");
    check_shape(sl, {b}, "seqlens_k");
    if (paged) {
      // The kernel indexes block_table up to seqlens_k[b] (+ seqlen_knew): validate the caller contract like the
      // reference (flash_api.cpp:1433-1448; one device->host sync, paid on the paged path only)
      const int64_t max_sk = sl.max().item<int>() + s_new;
      TORCH_CHECK(max_sk <= max_blocks * page, "Paged KV cache: max(seqlens_k)", s_new > 0 ? " + seqlen_knew" : "", " (= ", max_sk,
                  ") exceeds the capacity addressable by block_table (max_num_blocks_per_seq * page_block_size = ",
                  max_blocks * page, "). Allocate more columns in block_table, otherwise the kernel would index block_table out of bounds.");
    }
  }
  if (cache_batch_idx_.has_value()) {
    TORCH_CHECK(cache_batch_idx_.value().scalar_type() == at::kInt, "cache_batch_idx must have dtype int32");
    TORCH_CHECK(!paged, "Paged KVcache does not support cache_batch_idx");
  }
  if (leftpad_k_.has_value()) {
    TORCH_CHECK(!paged, "We don't support Paged KV and leftpad_k running at the same time yet");
    TORCH_CHECK(leftpad_k_.value().dtype() == at::kInt, "leftpad_k must have dtype int32");
    TORCH_CHECK(seqlens_k_.has_value(), "leftpad_k requires seqlens_k");
  }
  if (rotary_cos_.has_value()) {  // flash_api.cpp:1454-1478
    TORCH_CHECK(k_.has_value(), "If rotary cos/sin are provided, new key / value to be appended to KV cache must also be provided");
    const auto& rc = rotary_cos_.value();
    TORCH_CHECK(rc.is_cuda() && rc.dim() == 2, "rotary_cos must be a 2-D CUDA tensor (seqlen_ro, rotary_dim / 2)");
    const int64_t rotary_dim = rc.size(1) * 2, seqlen_ro = rc.size(0);
    TORCH_CHECK(rotary_dim <= d, "rotary_dim must be <= headdim");
    TORCH_CHECK(rotary_dim % 16 == 0, "Only rotary dimensions divisible by 16 are currently supported");
    TORCH_CHECK(seqlen_ro >= sk, "cos/sin seqlen must be at least the seqlen of KV cache");
    TORCH_CHECK(rc.is_contiguous(), "rotary_cos must be contiguous");
    TORCH_CHECK(rc.scalar_type() == q.scalar_type(), "rotary_cos must have the same dtype as query");
    TORCH_CHECK(rotary_sin_.has_value(), "If rotary cos is provided, rotary sin must also be provided");
    const auto& rs = rotary_sin_.value();
    TORCH_CHECK(rs.is_cuda() && rs.is_contiguous(), "rotary_sin must be contiguous");
    check_shape(rs, {seqlen_ro, rotary_dim / 2}, "rotary_sin");
    TORCH_CHECK(rs.scalar_type() == q.scalar_type(), "rotary_cos must have the same dtype as query");
  }
  at::Tensor seqlens_eff;  // lengths the attention runs with (after the append)
  if (k_.has_value()) {
    // append new keys / values into the cache at position cache_seqlens (flash_fwd_kernel.h:691-790)
    TORCH_CHECK(v_.has_value(), "If key is supplied, value must also be passed in");
    TORCH_CHECK(seqlens_k_.has_value(), "If key is supplied, seqlens_k must also be passed in");
    const auto& kn = k_.value();
    const auto& vn = v_.value();
    TORCH_CHECK(kn.dtype() == q.dtype() && vn.dtype() == q.dtype(), "Key and value must have the same dtype as query");
    TORCH_CHECK(kn.is_cuda() && vn.is_cuda(), "Input tensor must be on CUDA device");
    TORCH_CHECK(kn.stride(-1) == 1 && vn.stride(-1) == 1, "Key / value tensor must have contiguous last dimension");
    check_shape(kn, {b, s_new, hk, d}, "k");
    check_shape(vn, {b, s_new, hk, d}, "v");
    TORCH_CHECK(s_new <= sk, "fa_b200: more new keys than the cache can hold");
    // query i is rotated at cache_seqlens + i when causal or local, else all queries at cache_seqlens
    const bool local = window_size_left >= 0 || window_size_right >= 0;
    fa_b200_append_args ap = {};
    ap.kcache = kcache.data_ptr(); ap.vcache = vcache.data_ptr();
    ap.kc_batch_stride = kcache.stride(0); ap.kc_row_stride = kcache.stride(1); ap.kc_head_stride = kcache.stride(2);
    ap.vc_batch_stride = vcache.stride(0); ap.vc_row_stride = vcache.stride(1); ap.vc_head_stride = vcache.stride(2);
    ap.knew = kn.data_ptr(); ap.vnew = vn.data_ptr();
    ap.kn_batch_stride = kn.stride(0); ap.kn_row_stride = kn.stride(1); ap.kn_head_stride = kn.stride(2);
    ap.vn_batch_stride = vn.stride(0); ap.vn_row_stride = vn.stride(1); ap.vn_head_stride = vn.stride(2);
    ap.seqlen_new = int(s_new);
    if (rotary_cos_.has_value()) {
      auto q_rot = at::empty_like(q);
      ap.q = q.data_ptr(); ap.q_out = q_rot.data_ptr();
      ap.q_batch_stride = q.stride(0); ap.q_row_stride = q.stride(1); ap.q_head_stride = q.stride(2);
      ap.qo_batch_stride = q_rot.stride(0); ap.qo_row_stride = q_rot.stride(1); ap.qo_head_stride = q_rot.stride(2);
      ap.rotary_cos = rotary_cos_.value().data_ptr(); ap.rotary_sin = rotary_sin_.value().data_ptr();
      ap.rotary_dim = int(rotary_cos_.value().size(1) * 2);
      ap.is_rotary_interleaved = is_rotary_interleaved;
      ap.q_pos_per_row = (is_causal || local) ? 1 : 0;
      q = q_rot;
    }
    ap.cache_seqlens = seqlens_k_.value().data_ptr<int32_t>();
    ap.cache_batch_idx = int_vec(cache_batch_idx_, b, "cache_batch_idx");
    ap.leftpad_k = int_vec(leftpad_k_, b, "leftpad_k");
    if (paged) {
      ap.block_table = block_table_.value().data_ptr<int32_t>();
      ap.block_table_batch_stride = block_table_.value().stride(0);
      ap.page_block_size = int(page);
    }
    ap.b = int(b); ap.seqlen_q = int(sq); ap.h = int(h); ap.h_k = int(hk); ap.d = int(d);
    ap.dtype = dtype_code(q_in);
    check_rc(fa_b200_kvcache_append(&ap, cur_stream()));
    seqlens_eff = seqlens_k_.value() + s_new;
  } else if (seqlens_k_.has_value()) {
    seqlens_eff = seqlens_k_.value();
  }
  at::Tensor out;
  if (out_.has_value()) {
    out = out_.value();
    TORCH_CHECK(out.dtype() == q.dtype(), "Output must have the same dtype as inputs");
    TORCH_CHECK(out.is_cuda() && out.stride(-1) == 1, "Output tensor must have contiguous last dimension");
    check_shape(out, {b, sq, h, d}, "out");
  } else {
    out = at::empty_like(q_in);
  }
  auto lse = at::empty({b, h, sq}, q.options().dtype(at::kFloat));
  fa_b200_fwd_args a = {};
  set_qkvo_dense(a, q, kcache, vcache, out);
  a.softmax_lse = lse.data_ptr<float>();
  a.b = int(b); a.seqlen_q = int(sq); a.seqlen_k = int(sk); a.h = int(h); a.h_k = int(hk); a.d = int(d);
  a.dtype = dtype_code(q);
  if (paged) {
    a.block_table = block_table_.value().data_ptr<int32_t>();
    a.block_table_batch_stride = block_table_.value().stride(0);
    a.page_block_size = int(page); a.num_pages = int(num_pages);
  }
  if (seqlens_eff.defined()) a.seqused_k = seqlens_eff.data_ptr<int32_t>();
  a.cache_batch_idx = int_vec(cache_batch_idx_, b, "cache_batch_idx");
  a.leftpad_k = int_vec(leftpad_k_, b, "leftpad_k");
  a.b_kvcache = paged ? 0 : int(kcache.size(0));
  a.softmax_scale = float(softmax_scale); a.softcap = float(softcap);
  a.is_causal = is_causal; a.window_size_left = int(window_size_left); a.window_size_right = int(window_size_right);
  a.num_splits = int(num_splits);
  a.entry = FA_B200_ENTRY_KVCACHE;
  set_alibi(a, alibi_slopes_, b, h);
  auto ws = fwd_workspace(a, q);
  check_rc(fa_b200_mha_fwd_kvcache(&a, cur_stream()));
  return {out, lse};
}

// ------------------------------------------------------------------------------------------------ bwd
static at::Tensor alloc_or_check(std::optional<at::Tensor>& t_, const at::Tensor& like, const char* name) {
  if (!t_.has_value()) return at::empty_like(like);
  auto& t = t_.value();
  TORCH_CHECK(t.dtype() == like.dtype(), name, " must have the same dtype as q");
  TORCH_CHECK(t.is_cuda(), name, " must be on CUDA device");
  TORCH_CHECK(t.stride(-1) == 1, name, " must have contiguous last dimension");
  TORCH_CHECK(t.sizes() == like.sizes(), name, " has the wrong shape");
  return t;
}

static void bwd_dropout(fa_b200_bwd_args& a, double p_dropout, const std::optional<at::Tensor>& rng_state, double softcap) {
  // p_dropout + the forward's rng_state (flash_api.cpp:990-1003); the kernel reads (seed, offset) from the device tensor
  TORCH_CHECK(p_dropout >= 0.0 && p_dropout < 1.0, "p_dropout must be in [0, 1)");
  if (p_dropout == 0.0) return;
  TORCH_CHECK(softcap == 0.0, "Softcapping does not support dropout for now");
  TORCH_CHECK(rng_state.has_value(), "fa_b200: the backward of a dropout forward needs its rng_state");
  const auto& r = rng_state.value();
  TORCH_CHECK(r.is_cuda() && r.dtype() == at::kLong && r.numel() == 2 && r.is_contiguous(),
              "rng_state must be a contiguous CUDA int64 tensor of 2 elements");
  a.p_dropout = float(p_dropout);
  a.rng_state = reinterpret_cast<const uint64_t*>(r.data_ptr());
}

static at::Tensor bwd_finish_args(fa_b200_bwd_args& a, const at::Tensor& dout, const at::Tensor& q, const at::Tensor& k,
                                  const at::Tensor& v, const at::Tensor& out, const at::Tensor& lse, const at::Tensor& dq,
                                  const at::Tensor& dk, const at::Tensor& dv, const std::optional<at::Tensor>& alibi,
                                  int64_t b, int64_t h, double softmax_scale, bool is_causal, int64_t wl, int64_t wr,
                                  double softcap, bool deterministic) {
  a.dout = dout.data_ptr(); a.q = q.data_ptr(); a.k = k.data_ptr(); a.v = v.data_ptr(); a.out = out.data_ptr();
  a.softmax_lse = lse.data_ptr<float>();
  a.dq = dq.data_ptr(); a.dk = dk.data_ptr(); a.dv = dv.data_ptr();
  a.dtype = dtype_code(q);
  a.softmax_scale = float(softmax_scale); a.softcap = float(softcap);
  a.is_causal = is_causal; a.window_size_left = int(wl); a.window_size_right = int(wr);
  a.deterministic = deterministic;
  if (alibi.has_value()) {
    const auto& t = alibi.value();
    TORCH_CHECK(t.dtype() == at::kFloat && t.is_cuda() && t.stride(-1) == 1, "ALiBi slopes must be fp32 cuda contiguous");
    TORCH_CHECK(t.sizes() == at::IntArrayRef({h}) || t.sizes() == at::IntArrayRef({b, h}),
                "ALiBi slopes must have shape (nheads,) or (batch, nheads)");
    a.alibi_slopes = t.data_ptr<float>();
    a.alibi_slopes_batch_stride = t.dim() == 2 ? t.stride(0) : 0;
  }
  const size_t nbytes = fa_b200_bwd_workspace_bytes(&a);
  TORCH_CHECK(nbytes > 0, fa_b200_last_error());
  auto ws = at::empty({int64_t(nbytes)}, q.options().dtype(at::kByte));
  a.workspace = ws.data_ptr();
  a.workspace_bytes = nbytes;
  return ws;
}

// mha_bwd (flash_api.cpp:800-1008) -> [dq, dk, dv, softmax_d (b,h,round128(sq))]
std::vector<at::Tensor> bwd(const at::Tensor& dout, const at::Tensor& q, const at::Tensor& k, const at::Tensor& v,
                            const at::Tensor& out, const at::Tensor& softmax_lse, std::optional<at::Tensor> dq_,
                            std::optional<at::Tensor> dk_, std::optional<at::Tensor> dv_, std::optional<at::Tensor> alibi_slopes_,
                            double p_dropout, double softmax_scale, bool is_causal, int64_t window_size_left,
                            int64_t window_size_right, double softcap, bool deterministic,
                            const py::object& unused_generator_compat, std::optional<at::Tensor> rng_state) {
  check_generator_slot(unused_generator_compat);
  common_checks(q, k, v);
  c10::cuda::CUDAGuard guard(q.device());
  TORCH_CHECK(dout.dtype() == q.dtype(), "query and dout must have the same dtype");
  TORCH_CHECK(out.dtype() == q.dtype(), "query and out must have the same dtype");
  TORCH_CHECK(dout.is_cuda() && out.is_cuda() && softmax_lse.is_cuda(), "Input tensor must be on CUDA device");
  TORCH_CHECK(out.stride(-1) == 1, "out tensor must have contiguous last dimension");
  TORCH_CHECK(dout.stride(-1) == 1, "dout tensor must have contiguous last dimension");
  const int64_t b = q.size(0), sq = q.size(1), h = q.size(2), d = q.size(3), sk = k.size(1), hk = k.size(2);
  TORCH_CHECK(b > 0, "batch size must be positive");
  TORCH_CHECK(d % 8 == 0, "head_size should be a multiple of 8");
  TORCH_CHECK(d <= 256, "FlashAttention backward only supports head dimension at most 256");
  TORCH_CHECK(h % hk == 0, "Number of heads in key/value must divide number of heads in query");
  check_shape(k, {b, sk, hk, d}, "k");
  check_shape(v, {b, sk, hk, d}, "v");
  check_shape(out, {b, sq, h, d}, "out");
  check_shape(dout, {b, sq, h, d}, "dout");
  TORCH_CHECK(softmax_lse.dtype() == at::kFloat && softmax_lse.is_contiguous(), "softmax_lse must be a contiguous fp32 tensor");
  check_shape(softmax_lse, {b, h, sq}, "softmax_lse");
  auto dq = alloc_or_check(dq_, q, "dq");
  auto dk = alloc_or_check(dk_, k, "dk");
  auto dv = alloc_or_check(dv_, v, "dv");
  auto softmax_d = at::empty({b, h, round128(sq)}, q.options().dtype(at::kFloat));
  if (sq == 0) return {dq, dk, dv, softmax_d};
  if (sk == 0) {  // flash_api.cpp:995-1000
    dk.zero_(); dv.zero_(); dq.zero_(); softmax_d.zero_();
    return {dq, dk, dv, softmax_d};
  }
  fa_b200_bwd_args a = {};
  a.softmax_d = softmax_d.data_ptr<float>();
  bwd_dropout(a, p_dropout, rng_state, softcap);
#define FA_SET3(name, t) a.name##_batch_stride = (t).stride(0); a.name##_row_stride = (t).stride(1); a.name##_head_stride = (t).stride(2)
  FA_SET3(q, q); FA_SET3(k, k); FA_SET3(v, v); FA_SET3(o, out); FA_SET3(do, dout); FA_SET3(dq, dq); FA_SET3(dk, dk); FA_SET3(dv, dv);
#undef FA_SET3
  a.b = int(b); a.seqlen_q = int(sq); a.seqlen_k = int(sk); a.h = int(h); a.h_k = int(hk); a.d = int(d);
  auto ws = bwd_finish_args(a, dout, q, k, v, out, softmax_lse, dq, dk, dv, alibi_slopes_, b, h, softmax_scale, is_causal,
                            window_size_left, window_size_right, softcap, deterministic);
  check_rc(fa_b200_mha_bwd(&a, cur_stream()));
  return {dq, dk, dv, softmax_d};
}

// mha_varlen_bwd (flash_api.cpp:1010-1241) -> [dq, dk, dv, softmax_d (h, total_q + 128 * b)]
std::vector<at::Tensor> varlen_bwd(const at::Tensor& dout, const at::Tensor& q, const at::Tensor& k, const at::Tensor& v,
                                   const at::Tensor& out, const at::Tensor& softmax_lse, std::optional<at::Tensor> dq_,
                                   std::optional<at::Tensor> dk_, std::optional<at::Tensor> dv_, const at::Tensor& cu_seqlens_q,
                                   const at::Tensor& cu_seqlens_k, std::optional<at::Tensor> alibi_slopes_, int64_t max_seqlen_q,
                                   int64_t max_seqlen_k, double p_dropout, double softmax_scale, bool zero_tensors, bool is_causal,
                                   int64_t window_size_left, int64_t window_size_right, double softcap, bool deterministic,
                                   const py::object& unused_generator_compat, std::optional<at::Tensor> rng_state) {
  check_generator_slot(unused_generator_compat);
  common_checks(q, k, v);
  c10::cuda::CUDAGuard guard(q.device());
  TORCH_CHECK(dout.dtype() == q.dtype(), "query and dout must have the same dtype");
  TORCH_CHECK(out.dtype() == q.dtype(), "query and out must have the same dtype");
  TORCH_CHECK(cu_seqlens_q.dtype() == at::kInt, "cu_seqlens_q must have dtype int32");
  TORCH_CHECK(cu_seqlens_k.dtype() == at::kInt, "cu_seqlens_k must have dtype int32");
  TORCH_CHECK(cu_seqlens_q.is_cuda() && cu_seqlens_k.is_cuda() && dout.is_cuda() && out.is_cuda() && softmax_lse.is_cuda(),
              "Input tensor must be on CUDA device");
  TORCH_CHECK(cu_seqlens_q.is_contiguous() && cu_seqlens_k.is_contiguous(), "cu_seqlens must be contiguous");
  TORCH_CHECK(out.stride(-1) == 1, "out tensor must have contiguous last dimension");
  TORCH_CHECK(dout.stride(-1) == 1, "dout tensor must have contiguous last dimension");
  const int64_t total_q = q.size(0), h = q.size(1), d = q.size(2), total_k = k.size(0), hk = k.size(1);
  const int64_t b = cu_seqlens_q.numel() - 1;
  TORCH_CHECK(b > 0, "batch size must be positive");
  TORCH_CHECK(d % 8 == 0, "head_size should be a multiple of 8");
  TORCH_CHECK(d <= 256, "FlashAttention backward only supports head dimension at most 256");
  TORCH_CHECK(h % hk == 0, "Number of heads in key/value must divide number of heads in query");
  check_shape(v, {total_k, hk, d}, "v");
  check_shape(out, {total_q, h, d}, "out");
  check_shape(dout, {total_q, h, d}, "dout");
  check_shape(cu_seqlens_k, {b + 1}, "cu_seqlens_k");
  TORCH_CHECK(softmax_lse.dtype() == at::kFloat && softmax_lse.is_contiguous(), "softmax_lse must be a contiguous fp32 tensor");
  check_shape(softmax_lse, {h, total_q}, "softmax_lse");
  auto dq = alloc_or_check(dq_, q, "dq");
  auto dk = alloc_or_check(dk_, k, "dk");
  auto dv = alloc_or_check(dv_, v, "dv");
  auto softmax_d = at::empty({h, total_q + 128 * b}, q.options().dtype(at::kFloat));
  if (zero_tensors) { dq.zero_(); dk.zero_(); dv.zero_(); softmax_d.zero_(); }
  if (total_q == 0) return {dq, dk, dv, softmax_d};
  if (max_seqlen_k == 0 || total_k == 0) {
    dk.zero_(); dv.zero_(); dq.zero_(); softmax_d.zero_();
    return {dq, dk, dv, softmax_d};
  }
  fa_b200_bwd_args a = {};
  a.softmax_d = softmax_d.data_ptr<float>();
  bwd_dropout(a, p_dropout, rng_state, softcap);
#define FA_SET3(name, t) a.name##_batch_stride = 0; a.name##_row_stride = (t).stride(0); a.name##_head_stride = (t).stride(1)
  FA_SET3(q, q); FA_SET3(k, k); FA_SET3(v, v); FA_SET3(o, out); FA_SET3(do, dout); FA_SET3(dq, dq); FA_SET3(dk, dk); FA_SET3(dv, dv);
#undef FA_SET3
  a.b = int(b); a.seqlen_q = int(max_seqlen_q); a.seqlen_k = int(max_seqlen_k); a.h = int(h); a.h_k = int(hk); a.d = int(d);
  a.total_q = int(total_q); a.total_k = int(total_k);
  a.cu_seqlens_q = cu_seqlens_q.data_ptr<int32_t>();
  a.cu_seqlens_k = cu_seqlens_k.data_ptr<int32_t>();
  auto ws = bwd_finish_args(a, dout, q, k, v, out, softmax_lse, dq, dk, dv, alibi_slopes_, b, h, softmax_scale, is_causal,
                            window_size_left, window_size_right, softcap, deterministic);
  check_rc(fa_b200_mha_varlen_bwd(&a, cur_stream()));
  return {dq, dk, dv, softmax_d};
}

// flash_attn_combine (flash_attn/cute/interface.py:3687; hopper/flash_api.cpp mha_combine): o_partial (splits, ..., d) fp32,
// lse_partial (splits, ...) fp32 -> (out (..., d), lse (...))
std::tuple<at::Tensor, at::Tensor> fwd_combine(const at::Tensor& o_partial_, const at::Tensor& lse_partial_,
                                               std::optional<at::ScalarType> out_dtype, std::optional<at::Tensor> out_) {
  TORCH_CHECK(o_partial_.is_cuda() && lse_partial_.is_cuda(), "Input tensor must be on CUDA device");
  TORCH_CHECK(o_partial_.dtype() == at::kFloat && lse_partial_.dtype() == at::kFloat, "partials must be fp32");
  TORCH_CHECK(o_partial_.dim() >= 2 && lse_partial_.dim() == o_partial_.dim() - 1, "out_partial (splits, ..., d) and lse_partial (splits, ...)");
  c10::cuda::CUDAGuard guard(o_partial_.device());
  auto o_partial = o_partial_.contiguous();
  auto lse_partial = lse_partial_.contiguous();
  const int64_t splits = o_partial.size(0), d = o_partial.size(-1);
  TORCH_CHECK(lse_partial.size(0) == splits, "lse_partial has the wrong number of splits");
  const int64_t rows = lse_partial.numel() / std::max<int64_t>(splits, 1);
  const at::ScalarType dt = out_.has_value() ? out_.value().scalar_type() : out_dtype.value_or(at::kBFloat16);
  TORCH_CHECK(dt == at::kBFloat16 || dt == at::kHalf, "combine: output dtype must be fp16 or bf16");
  auto shape = o_partial.sizes().slice(1).vec();
  at::Tensor out = out_.has_value() ? out_.value() : at::empty(shape, o_partial.options().dtype(dt));
  TORCH_CHECK(out.is_contiguous() && out.sizes() == at::IntArrayRef(shape), "out has the wrong shape / is not contiguous");
  auto lse = at::empty(lse_partial.sizes().slice(1), lse_partial.options());
  check_rc(fa_b200_fwd_combine(o_partial.data_ptr<float>(), lse_partial.data_ptr<float>(), out.data_ptr(), lse.data_ptr<float>(),
                               int32_t(splits), rows, int32_t(d), dt == at::kBFloat16 ? FA_B200_BF16 : FA_B200_FP16, cur_stream()));
  return {out, lse};
}

}  // namespace fa_ext

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  using namespace fa_ext;
  m.doc() = "fa_b200: B200-native (sm_100a) FlashAttention operator module, drop-in for the reference's flash_attn_2_cuda";
  m.def("fwd", &fwd, "Forward pass", py::arg("q"), py::arg("k"), py::arg("v"), py::arg("out_"), py::arg("alibi_slopes_"),
        py::arg("p_dropout"), py::arg("softmax_scale"), py::arg("is_causal"), py::arg("window_size_left"),
        py::arg("window_size_right"), py::arg("softcap"), py::arg("return_softmax"),
        py::arg("unused_generator_compat") = py::none());
  m.def("varlen_fwd", &varlen_fwd, "Forward pass (variable length)", py::arg("q"), py::arg("k"), py::arg("v"), py::arg("out_"),
        py::arg("cu_seqlens_q"), py::arg("cu_seqlens_k"), py::arg("seqused_k"), py::arg("leftpad_k_"), py::arg("block_table_"),
        py::arg("alibi_slopes_"), py::arg("max_seqlen_q"), py::arg("max_seqlen_k"), py::arg("p_dropout"),
        py::arg("softmax_scale"), py::arg("zero_tensors"), py::arg("is_causal"), py::arg("window_size_left"),
        py::arg("window_size_right"), py::arg("softcap"), py::arg("return_softmax"),
        py::arg("unused_generator_compat") = py::none(), py::arg("num_splits") = 0, py::arg("seqused_q") = py::none());
  m.def("bwd", &bwd, "Backward pass", py::arg("dout"), py::arg("q"), py::arg("k"), py::arg("v"), py::arg("out"),
        py::arg("softmax_lse"), py::arg("dq_"), py::arg("dk_"), py::arg("dv_"), py::arg("alibi_slopes_"), py::arg("p_dropout"),
        py::arg("softmax_scale"), py::arg("is_causal"), py::arg("window_size_left"), py::arg("window_size_right"),
        py::arg("softcap"), py::arg("deterministic"), py::arg("unused_generator_compat") = py::none(),
        py::arg("rng_state") = py::none());
  m.def("varlen_bwd", &varlen_bwd, "Backward pass (variable length)", py::arg("dout"), py::arg("q"), py::arg("k"), py::arg("v"),
        py::arg("out"), py::arg("softmax_lse"), py::arg("dq_"), py::arg("dk_"), py::arg("dv_"), py::arg("cu_seqlens_q"),
        py::arg("cu_seqlens_k"), py::arg("alibi_slopes_"), py::arg("max_seqlen_q"), py::arg("max_seqlen_k"), py::arg("p_dropout"),
        py::arg("softmax_scale"), py::arg("zero_tensors"), py::arg("is_causal"), py::arg("window_size_left"),
        py::arg("window_size_right"), py::arg("softcap"), py::arg("deterministic"),
        py::arg("unused_generator_compat") = py::none(), py::arg("rng_state") = py::none());
  m.def("fwd_kvcache", &fwd_kvcache, "Forward pass, with KV-cache", py::arg("q"), py::arg("kcache"), py::arg("vcache"), py::arg("k_"),
        py::arg("v_"), py::arg("seqlens_k_"), py::arg("rotary_cos_"), py::arg("rotary_sin_"), py::arg("cache_batch_idx_"),
        py::arg("leftpad_k_"), py::arg("block_table_"), py::arg("alibi_slopes_"), py::arg("out_"), py::arg("softmax_scale"),
        py::arg("is_causal"), py::arg("window_size_left"), py::arg("window_size_right"), py::arg("softcap"),
        py::arg("is_rotary_interleaved"), py::arg("num_splits"));
  m.def("fwd_fp8", &fwd_fp8, "FP8 e4m3 forward pass", py::arg("q"), py::arg("k"), py::arg("v"), py::arg("out_"), py::arg("q_descale"),
        py::arg("k_descale"), py::arg("v_descale"), py::arg("softmax_scale"), py::arg("is_causal"), py::arg("window_size_left"),
        py::arg("window_size_right"), py::arg("softcap") = 0.0, py::arg("alibi_slopes_") = py::none(), py::arg("num_splits") = 0,
        py::arg("pack_gqa") = py::none());
  m.def("fwd_ex", &fwd_ex, "Forward pass with num_splits / pack_gqa", py::arg("q"), py::arg("k"), py::arg("v"), py::arg("softmax_scale"),
        py::arg("is_causal"), py::arg("window_size_left"), py::arg("window_size_right"), py::arg("softcap") = 0.0,
        py::arg("num_splits") = 0, py::arg("pack_gqa") = py::none(), py::arg("alibi_slopes_") = py::none());
  m.def("plan_kvcache", &plan_kvcache, "Host-only: the decode plan", py::arg("batch"), py::arg("seqlen_q"), py::arg("seqlen_k"),
        py::arg("h"), py::arg("hk"), py::arg("d"), py::arg("page_size"), py::arg("causal"), py::arg("wl"), py::arg("wr"),
        py::arg("num_splits") = 0, py::arg("pack_gqa") = py::none());
  m.def("fwd_combine", &fwd_combine, "Combine split attention partials", py::arg("o_partial"), py::arg("lse_partial"),
        py::arg("out_dtype") = py::none(), py::arg("out") = py::none());
  m.def("launch_count", [] { return fa_b200_launch_count(); });
  m.def("library_version", [] { return std::string(fa_b200_version()); });
}
