// TEST INFRASTRUCTURE (oracle/_ref build only) — never part of the product path.
//
// Registers the reference's OWN CUDA entry points (compiled where they lie under /root/reference/kernels by
// oracle/build_ref_cuda.py) under torch.ops._ref_cuda_C.*, so that GPU tests and tests/bench_vs_ref_cuda.py can run the
// reference kernels and this repo's kernels in the same process on the same tensors. The reference's own
// kernels/torch_bindings.cpp cannot be used: it references every op of the tree (CUTLASS, Machete, Mamba, ...),
// most of which are out of scope and do not build offline.
//
// Prototypes come from the reference headers (included by path; nothing is copied). Schemas are inferred from the C++
// signatures. gptq_marlin_gemm takes the weight type as four integers and builds the reference's ScalarTypeTorch
// in C++, so that the `_core_C.ScalarType` custom class does not have to be registered twice in one process.
#include <torch/library.h>
#include <torch/all.h>

#include "ops.h"                      // /root/reference/kernels/ops.h
#include "cache.h"                    // /root/reference/kernels/cache.h
#include "quantization/quant_ops.h"   // gptq_marlin_gemm, repack, awq_dequantize
#include "moe/moe_ops.h"              // topk_softmax
#include "moe/marlin_moe_ops.h"       // marlin_gemm_moe

static torch::Tensor ref_gptq_marlin_gemm(torch::Tensor a, torch::Tensor b_q_weight, torch::Tensor b_scales,
                                          torch::Tensor b_zeros, torch::Tensor g_idx, torch::Tensor perm,
                                          torch::Tensor workspace, int64_t type_exponent, int64_t type_mantissa,
                                          int64_t type_bias, bool type_signed, int64_t size_m, int64_t size_n,
                                          int64_t size_k, bool is_k_full, bool has_zp, bool use_fp32_reduce,
                                          bool is_zp_float) {
  // via the base-class constructor (exponent, mantissa, signed, bias): the four-integer constructor of
  // ScalarTypeTorch forwards (bias, signed) in swapped positions (kernels/core/scalar_type.hpp:331-333 vs :31-39)
  auto t = c10::make_intrusive<aphrodite::ScalarTypeTorch>(
      aphrodite::ScalarType((uint8_t)type_exponent, (uint8_t)type_mantissa, type_signed, (int32_t)type_bias));
  return gptq_marlin_gemm(a, b_q_weight, b_scales, b_zeros, g_idx, perm, workspace, t, size_m, size_n, size_k,
                          is_k_full, has_zp, use_fp32_reduce, is_zp_float);
}

static void ref_copy_blocks(std::vector<torch::Tensor> key_caches, std::vector<torch::Tensor> value_caches,
                            torch::Tensor block_mapping) {
  copy_blocks(key_caches, value_caches, block_mapping);
}

TORCH_LIBRARY(_ref_cuda_C, m) {
  m.def("paged_attention_v1", &paged_attention_v1);
  m.def("paged_attention_v2", &paged_attention_v2);
  m.def("rms_norm", &rms_norm);
  m.def("fused_add_rms_norm", &fused_add_rms_norm);
  m.def("rotary_embedding", &rotary_embedding);
  m.def("batched_rotary_embedding", &batched_rotary_embedding);
  m.def("silu_and_mul", &silu_and_mul);
  m.def("gelu_and_mul", &gelu_and_mul);
  m.def("gelu_tanh_and_mul", &gelu_tanh_and_mul);
  m.def("gelu_new", &gelu_new);
  m.def("gelu_fast", &gelu_fast);
  m.def("gelu_quick", &gelu_quick);
  m.def("reshape_and_cache", &reshape_and_cache);
  m.def("reshape_and_cache_flash", &reshape_and_cache_flash);
  m.def("copy_blocks", &ref_copy_blocks);
  m.def("convert_fp8", &convert_fp8);
  m.def("gptq_marlin_gemm", &ref_gptq_marlin_gemm);
  m.def("gptq_marlin_repack", &gptq_marlin_repack);
  m.def("awq_marlin_repack", &awq_marlin_repack);
  m.def("awq_dequantize", &awq_dequantize);
  m.def("moe_align_block_size", &moe_align_block_size);
  m.def("topk_softmax", &topk_softmax);
  m.def("marlin_gemm_moe", &marlin_gemm_moe);
  m.def("advance_step_flashattn", &advance_step_flashattn);
  m.def("permute_cols", &permute_cols);
  // round 2: fp8 activation quantisation (kernels/quantization/fp8/common.cu) and the sampling kernels (kernels/sampling/sampling.cu)
  m.def("static_scaled_fp8_quant", &static_scaled_fp8_quant);
  m.def("dynamic_scaled_fp8_quant", &dynamic_scaled_fp8_quant);
  m.def("dynamic_per_token_scaled_fp8_quant", &dynamic_per_token_scaled_fp8_quant);
  m.def("sampling_from_probs", &sampling_from_probs);
  m.def("top_p_sampling_from_probs", &top_p_sampling_from_probs);
  m.def("top_k_sampling_from_probs", &top_k_sampling_from_probs);
  m.def("min_p_sampling_from_probs", &min_p_sampling_from_probs);
  m.def("top_k_top_p_sampling_from_probs", &top_k_top_p_sampling_from_probs);
  m.def("top_p_renorm_prob", &top_p_renorm_prob);
  m.def("top_k_renorm_prob", &top_k_renorm_prob);
  m.def("top_k_mask_logits", &top_k_mask_logits);
}
