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
}  // namespace

TORCH_LIBRARY(_moe_C, m) {
  m.def(
      "topk_softmax(Tensor! topk_weights, Tensor! topk_indices, Tensor! "
      "token_expert_indices, Tensor gating_output) -> ()");
  m.impl("topk_softmax", torch::kCUDA, &topk_softmax);
}

PyMODINIT_FUNC PyInit__moe_C() {
  static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_moe_C", nullptr, 0, nullptr};
  return PyModule_Create(&module);
}
