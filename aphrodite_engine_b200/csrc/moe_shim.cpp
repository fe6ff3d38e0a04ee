// moe_shim.cpp — `_moe_C` op namespace of the reference (kernels/moe/torch_bindings.cpp:9-27), forwarding
// to the C ABI. Built as `_moe_C.abi3.so` with PyInit__moe_C (the reference imports `aphrodite._moe_C`,
// aphrodite/_custom_ops.py:22-24).
#include <Python.h>

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/all.h>
#include <torch/library.h>

#include "b200_decode.h"

namespace {
void topk_softmax(torch::Tensor& topk_weights, torch::Tensor& topk_indices,
                  torch::Tensor& token_expert_indices, torch::Tensor& gating_output) {
  const int num_experts = (int)gating_output.size(-1);
  const int num_tokens = (int)(gating_output.numel() / num_experts);
  const int topk = (int)topk_weights.size(-1);
  TORCH_CHECK(gating_output.scalar_type() == at::kFloat, "gating_output must be float32");
  TORCH_CHECK(gating_output.is_contiguous(), "gating_output must be contiguous");
  const at::cuda::OptionalCUDAGuard guard(device_of(gating_output));
  const int rc = b200_topk_softmax(topk_weights.data_ptr<float>(), topk_indices.data_ptr<int>(),
                                   token_expert_indices.data_ptr<int>(), gating_output.data_ptr<float>(),
                                   num_tokens, num_experts, topk,
                                   (void*)at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, b200_last_error());
}

// marlin_gemm_moe (kernels/moe/marlin_moe_ops.cu:1482-1546): allocates and returns c [size_m, topk, size_n]
torch::Tensor marlin_gemm_moe(const torch::Tensor& a, const torch::Tensor& b_q_weights,
                              const torch::Tensor& sorted_ids, const torch::Tensor& topk_weights,
                              const torch::Tensor& topk_ids, const torch::Tensor& b_scales,
                              const torch::Tensor& g_idx, const torch::Tensor& perm, torch::Tensor& workspace,
                              int64_t size_m, int64_t size_n, int64_t size_k, bool is_k_full, int64_t num_experts,
                              int64_t topk, int64_t moe_block_size, bool replicate_input, bool apply_weights) {
  (void)workspace;   // the reference's lock workspace: this kernel never splits k across CTAs
  TORCH_CHECK(a.scalar_type() == at::kHalf || a.scalar_type() == at::kBFloat16,
              "marlin_gemm_moe only supports bfloat16 and float16");
  TORCH_CHECK(a.is_contiguous() && b_q_weights.is_contiguous() && b_scales.is_contiguous() &&
              sorted_ids.is_contiguous() && topk_ids.is_contiguous() && topk_weights.is_contiguous(),
              "marlin_gemm_moe: all tensors must be contiguous");
  TORCH_CHECK(b_scales.dim() == 3, "b_scales rank = ", b_scales.dim(), " is not 3");
  TORCH_CHECK(b_scales.size(2) == size_n, "b_scales dim 2 = ", b_scales.size(2), " is not size_n = ", size_n);
  TORCH_CHECK(b_scales.scalar_type() == a.scalar_type(), "b_scales must have the dtype of a");
  TORCH_CHECK(b_q_weights.dim() == 3 && b_q_weights.size(0) == num_experts && b_q_weights.size(1) == size_k / 16 &&
              b_q_weights.size(2) == size_n * 2, "b_q_weights must be int32 [E, size_k/16, size_n*2]");
  TORCH_CHECK(sorted_ids.scalar_type() == at::kInt && topk_ids.scalar_type() == at::kInt,
              "sorted_ids / topk_ids must be int32");
  TORCH_CHECK(topk_weights.scalar_type() == at::kFloat, "topk_weights must be float32");
  TORCH_CHECK(a.numel() == (replicate_input ? size_m : size_m * topk) * size_k,
              "marlin_gemm_moe: a must hold ", replicate_input ? "size_m" : "size_m * topk", " rows of size_k");
  const int64_t num_groups = b_scales.size(1);
  const bool has_act_order = g_idx.dim() == 2 && g_idx.size(1) != 0;
  if (has_act_order) {
    TORCH_CHECK(is_k_full, "act_order with a k-sharded weight (is_k_full = False) is not implemented in the B200 marlin kernel");
    TORCH_CHECK(num_groups > 1, "For act_order, num_groups must be > 1");
    TORCH_CHECK(perm.scalar_type() == at::kInt && perm.is_contiguous() && perm.numel() == num_experts * size_k,
                "perm must be int32 [E, size_k]");
  }
  if (num_groups > 1)
    TORCH_CHECK(size_k % num_groups == 0, "size_k = ", size_k, ", is not divisible by b_scales.size(0) = ", b_scales.size(0));
  const at::cuda::OptionalCUDAGuard guard(device_of(a));
  torch::Tensor c = torch::zeros({size_m, topk, size_n}, a.options());
  torch::Tensor a_sorted = torch::empty({sorted_ids.numel(), size_k}, a.options());
  torch::Tensor expert_offsets = torch::empty({num_experts + 1}, sorted_ids.options());
  const int dtype = a.scalar_type() == at::kHalf ? B200_F16 : B200_BF16;
  const int rc = b200_marlin_gemm_moe(
      a.data_ptr(), b_q_weights.data_ptr(), sorted_ids.data_ptr<int>(), sorted_ids.numel(),
      topk_weights.data_ptr<float>(), topk_ids.data_ptr<int>(), b_scales.data_ptr(),
      has_act_order ? perm.data_ptr<int>() : nullptr, c.data_ptr(), a_sorted.data_ptr(),
      expert_offsets.data_ptr<int>(), (int)size_m, (int)size_n, (int)size_k, (int)num_groups, (int)num_experts,
      (int)topk, (int)moe_block_size, replicate_input ? 1 : 0, apply_weights ? 1 : 0, dtype,
      (void*)at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, b200_last_error());
  return c;
}
}  // namespace

TORCH_LIBRARY(_moe_C, m) {
  m.def(
      "topk_softmax(Tensor! topk_weights, Tensor! topk_indices, Tensor! "
      "token_expert_indices, Tensor gating_output) -> ()");
  m.impl("topk_softmax", torch::kCUDA, &topk_softmax);
  m.def(
      "marlin_gemm_moe(Tensor! a, Tensor! b_q_weights, Tensor! sorted_ids, "
      "Tensor! topk_weights, Tensor! topk_ids, Tensor! b_scales, Tensor! "
      "g_idx, Tensor! perm, Tensor! workspace, int size_m, int size_n, int "
      "size_k, bool is_k_full, int num_experts, int topk, int moe_block_size, "
      "bool replicate_input, bool apply_weights) -> Tensor");
  m.impl("marlin_gemm_moe", torch::kCUDA, &marlin_gemm_moe);
}

PyMODINIT_FUNC PyInit__moe_C() {
  static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_moe_C", nullptr, 0, nullptr};
  return PyModule_Create(&module);
}
