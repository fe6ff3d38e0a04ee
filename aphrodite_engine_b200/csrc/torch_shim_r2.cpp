// torch_shim_r2.cpp — second part of the `_C` torch-op shim (linked into the same _C.abi3.so as torch_shim.cpp):
// the SURVEY §8(f) rows — fp8 activation quantisation, the W8A8 scaled GEMM, the sampling kernels — under the
// reference's schemas (kernels/torch_bindings.cpp:235-253, 294-350, 374-390), plus this repo's own op for the prefill
// attention that the reference implements in Triton (no torch op there: aphrodite/attention/ops/prefix_prefill.py:696).
// Marshalling only; every op forwards to the C ABI (include/b200_decode.h).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/all.h>
#include <torch/library.h>

#include <string>
#include <vector>

#include "b200_decode.h"

namespace {

inline void check(int rc) { TORCH_CHECK(rc == 0, b200_last_error()); }
inline void* cur_stream() { return (void*)at::cuda::getCurrentCUDAStream().stream(); }
inline int dtype_code(const torch::Tensor& t, const char* what) {
  switch (t.scalar_type()) {
    case at::ScalarType::Float: return B200_F32;
    case at::ScalarType::Half: return B200_F16;
    case at::ScalarType::BFloat16: return B200_BF16;
    default: TORCH_CHECK(false, what, ": unsupported dtype ", t.scalar_type());
  }
  return -1;
}

// ---- fp8 activation quantisation (reference: kernels/quantization/fp8/common.cu:260-321) ----------------------
void static_scaled_fp8_quant(torch::Tensor& out, torch::Tensor const& input, torch::Tensor const& scale) {
  const at::cuda::OptionalCUDAGuard guard(device_of(input));
  TORCH_CHECK(out.scalar_type() == at::ScalarType::Float8_e4m3fn, "out must be float8_e4m3fn");
  TORCH_CHECK(input.is_contiguous() && out.is_contiguous() && scale.scalar_type() == at::kFloat, "static_scaled_fp8_quant: contiguous input / out, fp32 scale");
  check(b200_static_scaled_fp8_quant(out.data_ptr(), input.data_ptr(), scale.data_ptr<float>(), input.numel(),
                                     dtype_code(input, "static_scaled_fp8_quant"), cur_stream()));
}
void dynamic_scaled_fp8_quant(torch::Tensor& out, torch::Tensor const& input, torch::Tensor& scale) {
  const at::cuda::OptionalCUDAGuard guard(device_of(input));
  TORCH_CHECK(out.scalar_type() == at::ScalarType::Float8_e4m3fn, "out must be float8_e4m3fn");
  TORCH_CHECK(input.is_contiguous() && out.is_contiguous() && scale.scalar_type() == at::kFloat, "dynamic_scaled_fp8_quant: contiguous input / out, fp32 scale");
  check(b200_dynamic_scaled_fp8_quant(out.data_ptr(), input.data_ptr(), scale.data_ptr<float>(), input.numel(),
                                      dtype_code(input, "dynamic_scaled_fp8_quant"), cur_stream()));
}
void dynamic_per_token_scaled_fp8_quant(torch::Tensor& out, torch::Tensor const& input, torch::Tensor& scales,
                                        std::optional<at::Tensor> const& scale_ub) {
  TORCH_CHECK(input.is_contiguous());
  TORCH_CHECK(out.is_contiguous());
  const at::cuda::OptionalCUDAGuard guard(device_of(input));
  TORCH_CHECK(out.scalar_type() == at::ScalarType::Float8_e4m3fn, "out must be float8_e4m3fn");
  const int hidden = (int)input.size(-1);
  const int tokens = (int)(input.numel() / hidden);
  check(b200_dynamic_per_token_scaled_fp8_quant(out.data_ptr(), input.data_ptr(), scales.data_ptr<float>(),
                                                scale_ub ? scale_ub->data_ptr<float>() : nullptr, tokens, hidden,
                                                dtype_code(input, "dynamic_per_token_scaled_fp8_quant"), cur_stream()));
}

// ---- W8A8 scaled GEMM (reference: kernels/quantization/cutlass_w8a8/scaled_mm_entry.cu:92-140) ---------------------
bool cutlass_scaled_mm_supports_fp8(int64_t cuda_device_capability) {
  return b200_cutlass_scaled_mm_supports_fp8((int)cuda_device_capability) != 0;
}
void cutlass_scaled_mm(torch::Tensor& c, torch::Tensor const& a, torch::Tensor const& b, torch::Tensor const& a_scales,
                       torch::Tensor const& b_scales, std::optional<torch::Tensor> const& bias) {
  // the reference's conformality / stride checks, verbatim in meaning
  TORCH_CHECK(a.dim() == 2 && b.dim() == 2 && c.dim() == 2);
  TORCH_CHECK(c.size(0) == a.size(0) && a.size(1) == b.size(0) && b.size(1) == c.size(1));
  TORCH_CHECK(a_scales.numel() == 1 || a_scales.numel() == a.size(0));
  TORCH_CHECK(b_scales.numel() == 1 || b_scales.numel() == b.size(1));
  TORCH_CHECK(a.stride(1) == 1 && c.stride(1) == 1);                        // row-major
  TORCH_CHECK(b.stride(0) == 1);                                            // column-major
  TORCH_CHECK(c.stride(0) % 16 == 0 && b.stride(1) % 16 == 0);              // 16-byte alignment
  TORCH_CHECK(a_scales.is_contiguous() && b_scales.is_contiguous());
  TORCH_CHECK(a_scales.scalar_type() == at::kFloat && b_scales.scalar_type() == at::kFloat, "scales must be float32");
  if (bias) {
    TORCH_CHECK(bias->numel() == b.size(1) && bias->is_contiguous() && bias->dim() == 1);
    TORCH_CHECK(bias->scalar_type() == c.scalar_type(), "currently bias dtype must match output dtype ", c.dtype());
  }
  int ab = -1;
  if (a.scalar_type() == at::ScalarType::Float8_e4m3fn && b.scalar_type() == at::ScalarType::Float8_e4m3fn) ab = B200_AB_FP8_E4M3;
  if (a.scalar_type() == at::kChar && b.scalar_type() == at::kChar) ab = B200_AB_INT8;
  TORCH_CHECK(ab >= 0, "cutlass_scaled_mm: a and b must both be float8_e4m3fn or both be int8");
  const at::cuda::OptionalCUDAGuard guard(device_of(a));
  // k-split partial tiles: fp32 scratch from the caching allocator (safe under CUDA-graph capture); the reference op has
  // no workspace argument to carry it
  const int split = b200_scaled_mm_plan((int)a.size(0), (int)b.size(1), (int)a.size(1));
  torch::Tensor ws;
  if (split > 1) ws = torch::empty({(int64_t)split, a.size(0), b.size(1)}, torch::dtype(torch::kFloat32).device(a.device()));
  check(b200_cutlass_scaled_mm(c.data_ptr(), a.data_ptr(), b.data_ptr(), a_scales.data_ptr<float>(),
                               b_scales.data_ptr<float>(), bias ? bias->data_ptr() : nullptr, (int)a.size(0),
                               (int)b.size(1), (int)a.size(1), a.stride(0), b.stride(1), c.stride(0),
                               (int)a_scales.numel(), (int)b_scales.numel(), ab, dtype_code(c, "cutlass_scaled_mm"), split,
                               split > 1 ? ws.data_ptr() : nullptr, cur_stream()));
}
void cutlass_scaled_mm_azp(torch::Tensor& c, torch::Tensor const& a, torch::Tensor const& b, torch::Tensor const& a_scales,
                           torch::Tensor const& b_scales, torch::Tensor const& azp_adj,
                           std::optional<torch::Tensor> const& azp, std::optional<torch::Tensor> const& bias) {
  TORCH_CHECK(false, "cutlass_scaled_mm_azp (asymmetric int8 activations) is not implemented by the B200 library; "
                     "use symmetric quantisation (cutlass_scaled_mm)");
}

// ---- sampling (reference: kernels/sampling/sampling.cu) ---------------------------------------------------------
#define CHECK_CUDA_CONTIG(x) TORCH_CHECK((x).is_cuda() && (x).is_contiguous(), #x " must be a contiguous CUDA tensor")

torch::Tensor sampling_from_probs(torch::Tensor probs, torch::Tensor uniform_samples, bool deterministic) {
  CHECK_CUDA_CONTIG(probs);
  CHECK_CUDA_CONTIG(uniform_samples);
  TORCH_CHECK(probs.dim() == 2, "probs must be a 2D tensor");
  TORCH_CHECK(uniform_samples.dim() == 1, "uniform_samples must be a 1D tensor");
  TORCH_CHECK(probs.size(0) == uniform_samples.size(0), "CHECK_EQ(probs.size(0), uniform_samples.size(0)) failed. ",
              probs.size(0), " vs ", uniform_samples.size(0));
  const at::cuda::OptionalCUDAGuard guard(device_of(probs));
  probs = probs.to(torch::kFloat32);
  uniform_samples = uniform_samples.to(torch::kFloat32);
  auto samples = torch::empty({probs.size(0)}, torch::dtype(torch::kInt32).device(probs.device()));
  check(b200_sampling_from_probs(probs.data_ptr<float>(), uniform_samples.data_ptr<float>(), samples.data_ptr<int>(),
                                 (int)probs.size(0), (int)probs.size(1), deterministic ? 1 : 0, cur_stream()));
  return samples;
}

static std::vector<torch::Tensor> rejection_sample(int mode, torch::Tensor probs, torch::Tensor uniform_samples,
                                                   std::optional<torch::Tensor> k_arr, int64_t k_val,
                                                   std::optional<torch::Tensor> p_arr, double p_val, bool deterministic) {
  CHECK_CUDA_CONTIG(probs);
  CHECK_CUDA_CONTIG(uniform_samples);
  TORCH_CHECK(probs.dim() == 2, "probs must be a 2D tensor");
  TORCH_CHECK(uniform_samples.dim() == 2, "uniform_samples must be a 2D tensor");      // (max_rounds, batch_size)
  TORCH_CHECK(probs.size(0) == uniform_samples.size(1), "CHECK_EQ(probs.size(0), uniform_samples.size(1)) failed. ",
              probs.size(0), " vs ", uniform_samples.size(1));
  const int64_t batch = probs.size(0);
  const at::cuda::OptionalCUDAGuard guard(device_of(probs));
  torch::Tensor k_t, p_t;
  if (k_arr) {
    CHECK_CUDA_CONTIG(*k_arr);
    TORCH_CHECK(k_arr->dim() == 1 && k_arr->size(0) == batch, "top_k_arr must be a 1D tensor of batch_size entries");
    k_t = k_arr->to(torch::kInt32);
  }
  if (p_arr) {
    CHECK_CUDA_CONTIG(*p_arr);
    TORCH_CHECK(p_arr->dim() == 1 && p_arr->size(0) == batch, "the per-row parameter must be a 1D tensor of batch_size entries");
    p_t = p_arr->to(torch::kFloat32);
  }
  probs = probs.to(torch::kFloat32);
  uniform_samples = uniform_samples.to(torch::kFloat32);
  auto samples = torch::empty({batch}, torch::dtype(torch::kInt32).device(probs.device()));
  auto success = torch::empty({batch}, torch::dtype(torch::kBool).device(probs.device()));
  check(b200_rejection_sampling_from_probs(
      mode, probs.data_ptr<float>(), uniform_samples.data_ptr<float>(), samples.data_ptr<int>(),
      reinterpret_cast<uint8_t*>(success.data_ptr<bool>()), k_arr ? k_t.data_ptr<int>() : nullptr, (int)k_val,
      p_arr ? p_t.data_ptr<float>() : nullptr, (float)p_val, (int)batch, (int)probs.size(1), (int)uniform_samples.size(0),
      deterministic ? 1 : 0, cur_stream()));
  return {samples, success};
}
std::vector<torch::Tensor> top_k_sampling_from_probs(torch::Tensor probs, torch::Tensor uniform_samples,
                                                     std::optional<torch::Tensor> maybe_top_k_arr, int64_t top_k_val,
                                                     bool deterministic) {
  return rejection_sample(B200_SAMPLE_TOP_K, probs, uniform_samples, maybe_top_k_arr, top_k_val, std::nullopt, 0.0, deterministic);
}
std::vector<torch::Tensor> top_p_sampling_from_probs(torch::Tensor probs, torch::Tensor uniform_samples,
                                                     std::optional<torch::Tensor> maybe_top_p_arr, double top_p_val,
                                                     bool deterministic) {
  return rejection_sample(B200_SAMPLE_TOP_P, probs, uniform_samples, std::nullopt, 0, maybe_top_p_arr, top_p_val, deterministic);
}
std::vector<torch::Tensor> min_p_sampling_from_probs(torch::Tensor probs, torch::Tensor uniform_samples,
                                                     std::optional<torch::Tensor> maybe_min_p_arr, double min_p_val,
                                                     bool deterministic) {
  return rejection_sample(B200_SAMPLE_MIN_P, probs, uniform_samples, std::nullopt, 0, maybe_min_p_arr, min_p_val, deterministic);
}
std::vector<torch::Tensor> top_k_top_p_sampling_from_probs(torch::Tensor probs, torch::Tensor uniform_samples,
                                                           std::optional<torch::Tensor> maybe_top_k_arr, double top_k_val,
                                                           std::optional<torch::Tensor> maybe_top_p_arr, double top_p_val,
                                                           bool deterministic) {
  return rejection_sample(B200_SAMPLE_TOP_K_TOP_P, probs, uniform_samples, maybe_top_k_arr, (int64_t)top_k_val,
                          maybe_top_p_arr, top_p_val, deterministic);
}

template <int WHICH>   // 0 top_p_renorm_prob, 1 top_k_renorm_prob, 2 top_k_mask_logits
static torch::Tensor renorm(torch::Tensor x, std::optional<torch::Tensor> arr, double val) {
  CHECK_CUDA_CONTIG(x);
  TORCH_CHECK(x.dim() == 2, "probs / logits must be a 2D tensor");
  const int64_t batch = x.size(0);
  const at::cuda::OptionalCUDAGuard guard(device_of(x));
  torch::Tensor a_t;
  if (arr) {
    CHECK_CUDA_CONTIG(*arr);
    TORCH_CHECK(arr->dim() == 1 && arr->size(0) == batch, "the per-row parameter must be a 1D tensor of batch_size entries");
    a_t = arr->to(WHICH == 0 ? torch::kFloat32 : torch::kInt32);
  }
  x = x.to(torch::kFloat32);
  auto out = torch::empty({batch, x.size(1)}, torch::dtype(torch::kFloat32).device(x.device()));
  if (WHICH == 0)
    check(b200_top_p_renorm_prob(x.data_ptr<float>(), out.data_ptr<float>(), arr ? a_t.data_ptr<float>() : nullptr,
                                 (float)val, (int)batch, (int)x.size(1), cur_stream()));
  else if (WHICH == 1)
    check(b200_top_k_renorm_prob(x.data_ptr<float>(), out.data_ptr<float>(), arr ? a_t.data_ptr<int>() : nullptr,
                                 (int)val, (int)batch, (int)x.size(1), cur_stream()));
  else
    check(b200_top_k_mask_logits(x.data_ptr<float>(), out.data_ptr<float>(), arr ? a_t.data_ptr<int>() : nullptr,
                                 (int)val, (int)batch, (int)x.size(1), cur_stream()));
  return out;
}
torch::Tensor top_p_renorm_prob(torch::Tensor probs, std::optional<torch::Tensor> maybe_top_p_arr, double top_p_val) {
  return renorm<0>(probs, maybe_top_p_arr, top_p_val);
}
torch::Tensor top_k_renorm_prob(torch::Tensor probs, std::optional<torch::Tensor> maybe_top_k_arr, int64_t top_k_val) {
  return renorm<1>(probs, maybe_top_k_arr, (double)top_k_val);
}
torch::Tensor top_k_mask_logits(torch::Tensor logits, std::optional<torch::Tensor> maybe_top_k_arr, int64_t top_k_val) {
  return renorm<2>(logits, maybe_top_k_arr, (double)top_k_val);
}

// ---- prefill attention over the paged cache (this repo's op; reference: prefix_prefill.py:696 context_attention_fwd) --
void context_attention_fwd(torch::Tensor& q, torch::Tensor& k, torch::Tensor& v, torch::Tensor& o,
                           const std::string& kv_cache_dtype, torch::Tensor& k_cache, torch::Tensor& v_cache,
                           torch::Tensor& b_loc, torch::Tensor& b_start_loc, torch::Tensor& b_seq_len,
                           torch::Tensor& b_ctx_len, int64_t max_input_len, double k_scale, double v_scale,
                           const std::optional<torch::Tensor>& alibi_slopes, int64_t sliding_window) {
  const at::cuda::OptionalCUDAGuard guard(device_of(q));
  TORCH_CHECK(q.dim() == 3 && k.dim() == 3 && v.dim() == 3 && o.dim() == 3, "q / k / v / o must be [tokens, heads, head_size]");
  TORCH_CHECK(q.size(2) == k.size(2) && k.size(2) == v.size(2), "q, k and v must have the same head size");
  TORCH_CHECK(k.scalar_type() == q.scalar_type() && v.scalar_type() == q.scalar_type() && o.scalar_type() == q.scalar_type(),
              "q / k / v / o dtype mismatch");
  TORCH_CHECK(q.stride(2) == 1 && k.stride(2) == 1 && v.stride(2) == 1 && o.stride(2) == 1, "head dimension must be contiguous");
  TORCH_CHECK(k_cache.dim() == 5 && v_cache.dim() == 4, "k_cache [NB, Hkv, D/x, BS, x], v_cache [NB, Hkv, D, BS]");
  TORCH_CHECK(k_cache.stride(4) == 1 && k_cache.stride(3) == k_cache.size(4) &&
              k_cache.stride(2) == k_cache.size(3) * k_cache.size(4) && v_cache.stride(3) == 1 &&
              v_cache.stride(2) == v_cache.size(3), "the caches must be contiguous inside a (block, head)");
  TORCH_CHECK(b_loc.scalar_type() == at::kInt && b_start_loc.scalar_type() == at::kInt &&
              b_seq_len.scalar_type() == at::kInt && b_ctx_len.scalar_type() == at::kInt && b_loc.stride(1) == 1,
              "block table, start locations, sequence and context lengths must be int32");
  const int kv = b200_parse_kv_cache_dtype(kv_cache_dtype.c_str());
  TORCH_CHECK(kv >= 0, "Unsupported FP8 dtype: ", kv_cache_dtype);
  if (kv == B200_KV_AUTO) {
    TORCH_CHECK(k_cache.scalar_type() == q.scalar_type() && v_cache.scalar_type() == q.scalar_type(),
                "kv_cache_dtype='auto' unsupported for FP8 KV Cache prefill kernel");
  } else {
    TORCH_CHECK(k_cache.element_size() == 1 && v_cache.element_size() == 1, "fp8 kv cache must be stored as 1-byte elements");
  }
  const int D = (int)q.size(2);
  check(b200_context_attention_fwd(
      q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(),
      b_loc.data_ptr<int>(), b_start_loc.data_ptr<int>(), b_seq_len.data_ptr<int>(), b_ctx_len.data_ptr<int>(),
      alibi_slopes ? alibi_slopes->data_ptr<float>() : nullptr, (int)b_seq_len.size(0), (int)q.size(1), (int)k.size(1), D,
      (int)k_cache.size(3), (int)k_cache.size(4), (int)max_input_len, q.stride(0), q.stride(1), k.stride(0), k.stride(1),
      v.stride(0), v.stride(1), o.stride(0), o.stride(1), k_cache.stride(0), k_cache.stride(1), v_cache.stride(0),
      v_cache.stride(1), b_loc.stride(0), 1.0f / sqrtf((float)D), (float)k_scale, (float)v_scale, (int)sliding_window,
      dtype_code(q, "context_attention_fwd"), kv, cur_stream()));
}

}  // namespace

TORCH_LIBRARY_FRAGMENT(_C, ops) {
  // W8A8 GEMM, symmetric per-tensor or per-row/column quantisation (torch_bindings.cpp:235-253)
  ops.def(
      "cutlass_scaled_mm(Tensor! out, Tensor a,"
      "                  Tensor b, Tensor a_scales,"
      "                  Tensor b_scales, Tensor? bias) -> ()");
  ops.impl("cutlass_scaled_mm", torch::kCUDA, &cutlass_scaled_mm);
  ops.def("cutlass_scaled_mm_supports_fp8(int cuda_device_capability) -> bool");
  ops.impl("cutlass_scaled_mm_supports_fp8", &cutlass_scaled_mm_supports_fp8);
  ops.def(
      "cutlass_scaled_mm_azp(Tensor! out, Tensor a,"
      "                  Tensor b, Tensor a_scales,"
      "                  Tensor b_scales, Tensor azp_adj,"
      "                  Tensor? azp, Tensor? bias) -> ()");
  ops.impl("cutlass_scaled_mm_azp", torch::kCUDA, &cutlass_scaled_mm_azp);

  // sampling kernels (torch_bindings.cpp:294-350)
  ops.def(
      "sampling_from_probs(Tensor probs, Tensor uniform_samples, bool "
      "deterministic) -> Tensor");
  ops.impl("sampling_from_probs", torch::kCUDA, &sampling_from_probs);
  ops.def(
      "top_k_sampling_from_probs(Tensor probs, Tensor uniform_samples,"
      "                          Tensor? maybe_top_k_arr, int top_k_val,"
      "                          bool deterministic) -> Tensor[]");
  ops.impl("top_k_sampling_from_probs", torch::kCUDA, &top_k_sampling_from_probs);
  ops.def(
      "min_p_sampling_from_probs(Tensor probs, Tensor uniform_samples,"
      "                          Tensor? maybe_min_p_arr, float min_p_val,"
      "                          bool deterministic) -> Tensor[]");
  ops.impl("min_p_sampling_from_probs", torch::kCUDA, &min_p_sampling_from_probs);
  ops.def(
      "top_p_sampling_from_probs(Tensor probs, Tensor uniform_samples,"
      "                          Tensor? maybe_top_p_arr, float top_p_val,"
      "                          bool deterministic) -> Tensor[]");
  ops.impl("top_p_sampling_from_probs", torch::kCUDA, &top_p_sampling_from_probs);
  ops.def(
      "top_k_top_p_sampling_from_probs(Tensor probs, Tensor uniform_samples,"
      "                          Tensor? maybe_top_k_arr, float top_k_val,"
      "                          Tensor? maybe_top_p_arr, float top_p_val,"
      "                          bool deterministic) -> Tensor[]");
  ops.impl("top_k_top_p_sampling_from_probs", torch::kCUDA, &top_k_top_p_sampling_from_probs);
  ops.def(
      "top_k_renorm_prob(Tensor probs, Tensor? maybe_top_k_arr, int top_k_val) "
      "-> Tensor");
  ops.impl("top_k_renorm_prob", torch::kCUDA, &top_k_renorm_prob);
  ops.def(
      "top_p_renorm_prob(Tensor probs, Tensor? maybe_top_p_arr, float "
      "top_p_val) "
      "-> Tensor");
  ops.impl("top_p_renorm_prob", torch::kCUDA, &top_p_renorm_prob);
  ops.def(
      "top_k_mask_logits(Tensor logits, Tensor? maybe_top_k_arr, int "
      "top_k_val) -> Tensor");
  ops.impl("top_k_mask_logits", torch::kCUDA, &top_k_mask_logits);

  // fp8 activation quantisation (torch_bindings.cpp:374-390)
  ops.def(
      "static_scaled_fp8_quant(Tensor! out, Tensor input, Tensor scale) -> ()");
  ops.impl("static_scaled_fp8_quant", torch::kCUDA, &static_scaled_fp8_quant);
  ops.def(
      "dynamic_scaled_fp8_quant(Tensor! out, Tensor input, Tensor! scale) -> "
      "()");
  ops.impl("dynamic_scaled_fp8_quant", torch::kCUDA, &dynamic_scaled_fp8_quant);
  ops.def(
      "dynamic_per_token_scaled_fp8_quant(Tensor! out, Tensor input, "
      "Tensor! scale, Tensor? scale_ub) -> "
      "()");
  ops.impl("dynamic_per_token_scaled_fp8_quant", torch::kCUDA, &dynamic_per_token_scaled_fp8_quant);
}

// This repo's own op for the Triton-implemented prefill attention of the reference (argument list = the Python
// function context_attention_fwd's, prefix_prefill.py:696-711)
TORCH_LIBRARY_FRAGMENT(_C_b200, ext) {
  ext.def(
      "context_attention_fwd(Tensor q, Tensor k, Tensor v, Tensor! o, str kv_cache_dtype, Tensor k_cache, Tensor v_cache, "
      "Tensor b_loc, Tensor b_start_loc, Tensor b_seq_len, Tensor b_ctx_len, int max_input_len, float k_scale, "
      "float v_scale, Tensor? alibi_slopes, int sliding_window) -> ()");
  ext.impl("context_attention_fwd", torch::kCUDA, &context_attention_fwd);
}
