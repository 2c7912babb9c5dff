// fa_torch_common.h — helpers shared by the two ATen-facing translation units of the compiled operator module
// (flash_attn_2_cuda.cpp: the FA2 pybind surface; flash_attn_3_ops.cpp: TORCH_LIBRARY(flash_attn_3)).
#pragma once
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <initializer_list>
#include <optional>
#include <vector>

#include "../../../include/fa_b200.h"

namespace fa_ext {

inline void check_rc(int rc) { TORCH_CHECK(rc == 0, fa_b200_last_error()); }
inline fa_b200_stream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }
inline int64_t round128(int64_t x) { return (x + 127) / 128 * 128; }

inline int dtype_code(const at::Tensor& t) {
  if (t.scalar_type() == at::kHalf) return FA_B200_FP16;
  if (t.scalar_type() == at::kBFloat16) return FA_B200_BF16;
  if (t.scalar_type() == at::kFloat8_e4m3fn) return FA_B200_FP8_E4M3;
  TORCH_CHECK(false, "FlashAttention only support fp16 and bf16 data type");
  return -1;
}

inline void check_shape(const at::Tensor& t, std::initializer_list<int64_t> shape, const char* name) {
  TORCH_CHECK(t.sizes() == at::IntArrayRef(shape.begin(), shape.size()), name, " must have shape (",
              at::IntArrayRef(shape.begin(), shape.size()), "), got ", t.sizes());
}

// zero-element tensor of q's dtype on q's device ("p" when return_softmax is off)
inline at::Tensor empty_like_dtype(const at::Tensor& q) { return at::empty({0}, q.options()); }

// FA3 `pack_gqa` (None / True / False) -> fa_b200_fwd_args.pack_gqa (0 heuristic / 1 pack / 2 never)
inline int pack_code(const std::optional<bool>& p) { return !p.has_value() ? 0 : (p.value() ? 1 : 2); }

// optional (batch,) int32 contiguous CUDA vector -> device pointer or nullptr
inline const int32_t* int_vec(const std::optional<at::Tensor>& t_, int64_t b, const char* name) {
  if (!t_.has_value()) return nullptr;
  const auto& t = t_.value();
  TORCH_CHECK(t.dtype() == at::kInt, name, " must have dtype int32");
  TORCH_CHECK(t.is_cuda(), name, " must be on CUDA device");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
  TORCH_CHECK(t.numel() == b, name, " must have shape (batch_size,)");
  return t.data_ptr<int32_t>();
}

// split-KV scratch for a forward call (0 bytes -> undefined tensor); kept alive by the caller until the launch is enqueued
inline at::Tensor fwd_workspace(fa_b200_fwd_args& a, const at::Tensor& like) {
  const size_t nbytes = fa_b200_fwd_workspace_bytes(&a);
  if (nbytes == 0) return at::Tensor();
  auto ws = at::empty({int64_t(nbytes)}, like.options().dtype(at::kByte));
  a.workspace = ws.data_ptr();
  a.workspace_bytes = nbytes;
  return ws;
}

// fp8 de-scales: (b, h_k) fp32 each or None (= 1.0); contiguous copies are parked in `keep`
inline void set_descales(fa_b200_fwd_args& a, const std::optional<at::Tensor>& qd, const std::optional<at::Tensor>& kd,
                         const std::optional<at::Tensor>& vd, int64_t b, int64_t hk, std::vector<at::Tensor>& keep) {
  auto one = [&](const std::optional<at::Tensor>& t_, const char* name, const float** ptr, int64_t* bs) {
    if (!t_.has_value()) return;
    TORCH_CHECK(t_.value().dtype() == at::kFloat && t_.value().is_cuda(), name, " must be fp32 (batch, nheads_k) on CUDA");
    TORCH_CHECK(t_.value().dim() == 2 && t_.value().size(0) == b && t_.value().size(1) == hk, name, " must be fp32 (batch, nheads_k)");
    keep.push_back(t_.value().contiguous());
    *ptr = keep.back().data_ptr<float>();
    *bs = keep.back().stride(0);
  };
  one(qd, "q_descale", &a.q_descale, &a.q_descale_batch_stride);
  one(kd, "k_descale", &a.k_descale, &a.k_descale_batch_stride);
  one(vd, "v_descale", &a.v_descale, &a.v_descale_batch_stride);
}

}  // namespace fa_ext
