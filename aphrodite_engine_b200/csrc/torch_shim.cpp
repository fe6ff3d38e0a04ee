// torch_shim.cpp — the drop-in seam: registers the reference's torch ops (same namespaces, same
// schemas, same mutability annotations as kernels/torch_bindings.cpp:19-535 of the reference) and
// forwards each to the C ABI of libb200decode.so (include/b200_decode.h).
//
// Built as `_C.abi3.so` with PyInit__C, so `import aphrodite._C` (aphrodite/_custom_ops.py:13-24)
// can load this file in place of the reference's extension, or it can be injected with
// torch.ops.load_library() from an `aphrodite.general_plugins` entry point.
// Only marshalling lives here: device guard, current stream, strides, dtype codes, error mapping
// (non-zero C-ABI status -> c10::Error -> Python RuntimeError, like the reference's TORCH_CHECK).
#include <Python.h>

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/all.h>
#include <torch/library.h>
#include <torch/custom_class.h>
#include <ATen/core/stack.h>

#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "b200_decode.h"

namespace {

inline int dtype_code(const torch::Tensor& t, const char* what) {
  switch (t.scalar_type()) {
    case at::ScalarType::Float: return B200_F32;
    case at::ScalarType::Half: return B200_F16;
    case at::ScalarType::BFloat16: return B200_BF16;
    default: TORCH_CHECK(false, what, ": unsupported dtype ", t.scalar_type());
  }
  return -1;
}

inline int kv_code(const std::string& s) {
  const int c = b200_parse_kv_cache_dtype(s.c_str());
  TORCH_CHECK(c >= 0, "Unsupported data type of kv cache: ", s);
  return c;
}

inline void check(int rc) { TORCH_CHECK(rc == 0, b200_last_error()); }

inline void* cur_stream() { return (void*)at::cuda::getCurrentCUDAStream().stream(); }

// ---- attention -----------------------------------------------------------------------------
void paged_attention_v1(torch::Tensor& out, torch::Tensor& query, torch::Tensor& key_cache,
                        torch::Tensor& value_cache, int64_t num_kv_heads, double scale,
                        torch::Tensor& block_tables, torch::Tensor& seq_lens, int64_t block_size,
                        int64_t max_seq_len, const c10::optional<torch::Tensor>& alibi_slopes,
                        const std::string& kv_cache_dtype, double k_scale, double v_scale,
                        const int64_t tp_rank, const int64_t blocksparse_local_blocks,
                        const int64_t blocksparse_vert_stride,
                        const int64_t blocksparse_block_size,
                        const int64_t blocksparse_head_sliding_step) {
  const at::cuda::OptionalCUDAGuard guard(device_of(query));
  TORCH_CHECK(block_tables.scalar_type() == at::kInt && seq_lens.scalar_type() == at::kInt,
              "block_tables and seq_lens must be int32");
  check(b200_paged_attention_v1(
      out.data_ptr(), query.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr(),
      (int)query.size(0), (int)query.size(1), (int)num_kv_heads, (int)query.size(2),
      (int)block_size, (float)scale, block_tables.data_ptr<int>(), seq_lens.data_ptr<int>(),
      (int)block_tables.size(1), (int)max_seq_len,
      alibi_slopes ? reinterpret_cast<const float*>(alibi_slopes.value().data_ptr()) : nullptr,
      query.stride(0), key_cache.stride(0), key_cache.stride(1), dtype_code(query, "query"),
      kv_code(kv_cache_dtype), (float)k_scale, (float)v_scale, (int)tp_rank,
      (int)blocksparse_local_blocks, (int)blocksparse_vert_stride, (int)blocksparse_block_size,
      (int)blocksparse_head_sliding_step, cur_stream()));
}

void paged_attention_v2(torch::Tensor& out, torch::Tensor& exp_sums, torch::Tensor& max_logits,
                        torch::Tensor& tmp_out, torch::Tensor& query, torch::Tensor& key_cache,
                        torch::Tensor& value_cache, int64_t num_kv_heads, double scale,
                        torch::Tensor& block_tables, torch::Tensor& seq_lens, int64_t block_size,
                        int64_t max_seq_len, const c10::optional<torch::Tensor>& alibi_slopes,
                        const std::string& kv_cache_dtype, double k_scale, double v_scale,
                        const int64_t tp_rank, const int64_t blocksparse_local_blocks,
                        const int64_t blocksparse_vert_stride,
                        const int64_t blocksparse_block_size,
                        const int64_t blocksparse_head_sliding_step) {
  const at::cuda::OptionalCUDAGuard guard(device_of(query));
  TORCH_CHECK(block_tables.scalar_type() == at::kInt && seq_lens.scalar_type() == at::kInt,
              "block_tables and seq_lens must be int32");
  check(b200_paged_attention_v2(
      out.data_ptr(), exp_sums.data_ptr<float>(), max_logits.data_ptr<float>(), tmp_out.data_ptr(),
      query.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr(), (int)query.size(0),
      (int)query.size(1), (int)num_kv_heads, (int)query.size(2), (int)block_size, (float)scale,
      block_tables.data_ptr<int>(), seq_lens.data_ptr<int>(), (int)block_tables.size(1),
      (int)max_seq_len, (int)exp_sums.size(2),
      alibi_slopes ? reinterpret_cast<const float*>(alibi_slopes.value().data_ptr()) : nullptr,
      query.stride(0), key_cache.stride(0), key_cache.stride(1), dtype_code(query, "query"),
      kv_code(kv_cache_dtype), (float)k_scale, (float)v_scale, (int)tp_rank,
      (int)blocksparse_local_blocks, (int)blocksparse_vert_stride, (int)blocksparse_block_size,
      (int)blocksparse_head_sliding_step, cur_stream()));
}

// ---- activations / norm / rotary --------------------------------------------------------------
template <int ACT>
void act_and_mul(torch::Tensor& out, torch::Tensor& input) {
  const at::cuda::OptionalCUDAGuard guard(device_of(input));
  const int d = (int)(input.size(-1) / 2);
  const int64_t tokens = input.numel() / input.size(-1);
  check(b200_act_and_mul(out.data_ptr(), input.data_ptr(), (int)tokens, d, ACT,
                         dtype_code(input, "act_and_mul"), cur_stream()));
}
template <int ACT>
void activation(torch::Tensor& out, torch::Tensor& input) {
  const at::cuda::OptionalCUDAGuard guard(device_of(input));
  const int d = (int)input.size(-1);
  const int64_t tokens = input.numel() / d;
  check(b200_activation(out.data_ptr(), input.data_ptr(), (int)tokens, d, ACT,
                        dtype_code(input, "activation"), cur_stream()));
}

void rms_norm(torch::Tensor& out, torch::Tensor& input, torch::Tensor& weight, double epsilon) {
  const at::cuda::OptionalCUDAGuard guard(device_of(input));
  const int hidden = (int)input.size(-1);
  check(b200_rms_norm(out.data_ptr(), input.data_ptr(), weight.data_ptr(), (float)epsilon,
                      (int)(input.numel() / hidden), hidden, dtype_code(input, "rms_norm"),
                      cur_stream()));
}

void fused_add_rms_norm(torch::Tensor& input, torch::Tensor& residual, torch::Tensor& weight,
                        double epsilon) {
  const at::cuda::OptionalCUDAGuard guard(device_of(input));
  const int hidden = (int)input.size(-1);
  check(b200_fused_add_rms_norm(input.data_ptr(), residual.data_ptr(), weight.data_ptr(),
                                (float)epsilon, (int)(input.numel() / hidden), hidden,
                                dtype_code(input, "fused_add_rms_norm"), cur_stream()));
}

void rotary_embedding(torch::Tensor& positions, torch::Tensor& query, torch::Tensor& key,
                      int64_t head_size, torch::Tensor& cos_sin_cache, bool is_neox) {
  const at::cuda::OptionalCUDAGuard guard(device_of(query));
  const int64_t num_tokens = query.numel() / query.size(-1);
  check(b200_rotary_embedding(
      positions.data_ptr<int64_t>(), query.data_ptr(), key.data_ptr(), cos_sin_cache.data_ptr(),
      nullptr, (int)num_tokens, (int)(query.size(-1) / head_size), (int)(key.size(-1) / head_size),
      (int)head_size, (int)cos_sin_cache.size(1), query.stride(-2), key.stride(-2), is_neox ? 1 : 0,
      dtype_code(query, "rotary_embedding"), cur_stream()));
}

void batched_rotary_embedding(torch::Tensor& positions, torch::Tensor& query, torch::Tensor& key,
                              int64_t head_size, torch::Tensor& cos_sin_cache, bool is_neox,
                              int64_t rot_dim, torch::Tensor& cos_sin_cache_offsets) {
  const at::cuda::OptionalCUDAGuard guard(device_of(query));
  const int64_t num_tokens = cos_sin_cache_offsets.size(0);
  check(b200_rotary_embedding(
      positions.data_ptr<int64_t>(), query.data_ptr(), key.data_ptr(), cos_sin_cache.data_ptr(),
      cos_sin_cache_offsets.data_ptr<int64_t>(), (int)num_tokens,
      (int)(query.size(-1) / head_size), (int)(key.size(-1) / head_size), (int)head_size,
      (int)rot_dim, query.stride(-2), key.stride(-2), is_neox ? 1 : 0,
      dtype_code(query, "batched_rotary_embedding"), cur_stream()));
}

// ---- cache ops ------------------------------------------------------------------------------------
void swap_blocks(torch::Tensor& src, torch::Tensor& dst, const torch::Tensor& block_mapping) {
  const torch::Device sd = src.device(), dd = dst.device();
  int kind;
  if (sd.is_cuda() && dd.is_cuda()) {
    TORCH_CHECK(sd.index() == dd.index(), "src and dst must be on the same GPU");
    kind = 0;
  } else if (sd.is_cuda() && dd.is_cpu()) {
    kind = 1;
  } else if (sd.is_cpu() && dd.is_cuda()) {
    kind = 2;
  } else {
    TORCH_CHECK(false, "Invalid device combination");
  }
  TORCH_CHECK(block_mapping.device().is_cpu(), "block_mapping must be on CPU");
  TORCH_CHECK(block_mapping.scalar_type() == at::kLong, "block_mapping must be int64");
  const torch::Tensor bm = block_mapping.contiguous();
  const int64_t block_bytes = src.element_size() * src[0].numel();
  const at::cuda::OptionalCUDAGuard guard(sd.is_cuda() ? sd : dd);
  check(b200_swap_blocks(src.data_ptr(), dst.data_ptr(), bm.data_ptr<int64_t>(), (int)bm.size(0),
                         block_bytes, kind, cur_stream()));
}

void copy_blocks(std::vector<torch::Tensor> const& key_caches,
                 std::vector<torch::Tensor> const& value_caches,
                 const torch::Tensor& block_mapping) {
  const int num_layers = (int)key_caches.size();
  TORCH_CHECK(num_layers == (int)value_caches.size());
  if (num_layers == 0) return;
  const torch::Device dev = key_caches[0].device();
  TORCH_CHECK(dev.is_cuda());
  std::vector<int64_t> kp(num_layers), vp(num_layers);
  for (int i = 0; i < num_layers; ++i) {
    kp[i] = reinterpret_cast<int64_t>(key_caches[i].data_ptr());
    vp[i] = reinterpret_cast<int64_t>(value_caches[i].data_ptr());
  }
  // same host->device hand-off of the pointer tables as the reference (cache_kernels.cu:128-133)
  torch::Tensor kpt = torch::from_blob(kp.data(), {num_layers}, torch::kInt64).to(dev);
  torch::Tensor vpt = torch::from_blob(vp.data(), {num_layers}, torch::kInt64).to(dev);
  const int64_t block_bytes = key_caches[0][0].numel() * key_caches[0].element_size();
  const at::cuda::OptionalCUDAGuard guard(dev);
  check(b200_copy_blocks(kpt.data_ptr<int64_t>(), vpt.data_ptr<int64_t>(),
                         block_mapping.data_ptr<int64_t>(), num_layers, (int)block_mapping.size(0),
                         block_bytes, cur_stream()));
}

void reshape_and_cache(torch::Tensor& key, torch::Tensor& value, torch::Tensor& key_cache,
                       torch::Tensor& value_cache, torch::Tensor& slot_mapping,
                       const std::string& kv_cache_dtype, const double k_scale,
                       const double v_scale) {
  const at::cuda::OptionalCUDAGuard guard(device_of(key));
  check(b200_reshape_and_cache(
      key.data_ptr(), value.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr(),
      slot_mapping.data_ptr<int64_t>(), (int)key.size(0), (int)key.size(1), (int)key.size(2),
      (int)key_cache.size(3), (int)key_cache.size(4), key.stride(0), value.stride(0),
      dtype_code(key, "Unsupported input type of kv cache"), kv_code(kv_cache_dtype),
      (float)k_scale, (float)v_scale, cur_stream()));
}

void reshape_and_cache_flash(torch::Tensor& key, torch::Tensor& value, torch::Tensor& key_cache,
                             torch::Tensor& value_cache, torch::Tensor& slot_mapping,
                             const std::string& kv_cache_dtype, const double k_scale,
                             const double v_scale) {
  const at::cuda::OptionalCUDAGuard guard(device_of(key));
  TORCH_CHECK(key_cache.stride(0) == value_cache.stride(0));
  check(b200_reshape_and_cache_flash(
      key.data_ptr(), value.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr(),
      slot_mapping.data_ptr<int64_t>(), (int)key.size(0), (int)key.size(1), (int)key.size(2),
      (int)key_cache.size(1), key_cache.stride(0), key.stride(0), value.stride(0),
      dtype_code(key, "Unsupported input type of kv cache"), kv_code(kv_cache_dtype),
      (float)k_scale, (float)v_scale, cur_stream()));
}

void convert_fp8(torch::Tensor& dst_cache, torch::Tensor& src_cache, const double scale,
                 const std::string& kv_cache_dtype) {
  TORCH_CHECK(src_cache.device().is_cuda(), "src must be on a GPU");
  TORCH_CHECK(dst_cache.device().is_cuda(), "dst must be on a GPU");
  TORCH_CHECK(src_cache.device().index() == dst_cache.device().index(),
              "src and dst must be on the same GPU");
  const at::cuda::OptionalCUDAGuard guard(src_cache.device());
  auto code = [](const torch::Tensor& t) {
    switch (t.scalar_type()) {
      case at::ScalarType::Float: return (int)B200_F32;
      case at::ScalarType::Half: return (int)B200_F16;
      case at::ScalarType::BFloat16: return (int)B200_BF16;
      default: return -1;  // uint8 fp8 storage
    }
  };
  TORCH_CHECK(kv_cache_dtype == "auto" || kv_cache_dtype == "fp8" || kv_cache_dtype == "fp8_e4m3",
              "Unsupported data type: ", kv_cache_dtype);
  check(b200_convert_fp8(dst_cache.data_ptr(), src_cache.data_ptr(), src_cache.numel(),
                         code(src_cache), code(dst_cache), kv_code(kv_cache_dtype), (float)scale,
                         cur_stream()));
}

int64_t get_device_attribute(int64_t attribute, int64_t device_id) {
  return b200_get_device_attribute(attribute, device_id);
}
int64_t get_max_shared_memory_per_block_device_attribute(int64_t device_id) {
  return b200_get_max_shared_memory_per_block_device_attribute(device_id);
}


// ---- Marlin-format quantised GEMM ------------------------------------------------------------------
// `b_q_type` is a `_core_C.ScalarType` custom-class object. It is read through the class's registered
// property getters (size_bits, bias), NOT through a C++ type, so this op works with either this repo's
// `_core_C` (csrc/core_scalar_type.cpp) or the reference's own `_core_C` extension.
int64_t scalar_type_prop(const c10::IValue& st, const char* name) {
  auto obj = st.toObject();
  auto prop = obj->type()->getProperty(name);
  TORCH_CHECK(prop.has_value(), "b_q_type has no property '", name, "'");
  return (*prop->getter)({st}).toInt();
}

torch::Tensor gptq_marlin_gemm_impl(torch::Tensor& a, torch::Tensor& b_q_weight, torch::Tensor& b_scales,
                                    torch::Tensor& b_zeros, torch::Tensor& g_idx, torch::Tensor& perm,
                                    torch::Tensor& workspace, int64_t type_bits, int64_t type_bias,
                                    const std::string& type_str, int64_t size_m, int64_t size_n,
                                    int64_t size_k, bool is_k_full, bool has_zp, bool use_fp32_reduce,
                                    bool is_zp_float) {
  // argument contract of the reference entry point (gptq_marlin.cu:2255-2405), same messages
  if (has_zp) {
    TORCH_CHECK(type_bias == 0 && (type_bits == 4 || type_bits == 8),
                "b_q_type must be u4 or u8 when has_zp = True. Got = ", type_str);
  } else {
    TORCH_CHECK((type_bits == 4 && type_bias == 8) || (type_bits == 8 && type_bias == 128),
                "b_q_type must be uint4b8 or uint8b128 when has_zp = False. Got = ", type_str);
  }
  if (has_zp && is_zp_float)
    TORCH_CHECK(a.scalar_type() == at::kHalf, "Computation type must be float16 (half) when using float zero points.");
  const int64_t pack_factor = 32 / type_bits;
  TORCH_CHECK(a.size(0) == size_m, "Shape mismatch: a.size(0) = ", a.size(0), ", size_m = ", size_m);
  TORCH_CHECK(a.size(1) == size_k, "Shape mismatch: a.size(1) = ", a.size(1), ", size_k = ", size_k);
  TORCH_CHECK(size_k % 16 == 0, "size_k = ", size_k, " is not divisible by tile_size = 16");
  TORCH_CHECK((size_k / 16) == b_q_weight.size(0), "Shape mismatch: b_q_weight.size(0) = ",
              b_q_weight.size(0), ", size_k = ", size_k, ", tile_size = 16");
  TORCH_CHECK(b_q_weight.size(1) % 16 == 0, "b_q_weight.size(1) = ", b_q_weight.size(1),
              " is not divisible by tile_size = 16");
  const int64_t actual_size_n = (b_q_weight.size(1) / 16) * pack_factor;
  TORCH_CHECK(size_n == actual_size_n, "size_n = ", size_n, ", actual_size_n = ", actual_size_n);
  TORCH_CHECK(a.device().is_cuda(), "A is not on GPU");
  TORCH_CHECK(a.is_contiguous(), "A is not contiguous");
  TORCH_CHECK(b_q_weight.device().is_cuda() && b_q_weight.is_contiguous(), "b_q_weight must be a contiguous GPU tensor");
  TORCH_CHECK(b_scales.device().is_cuda() && b_scales.is_contiguous(), "b_scales must be a contiguous GPU tensor");
  TORCH_CHECK((g_idx.size(0) == 0 && perm.size(0) == 0) || (g_idx.size(0) == size_k && perm.size(0) == size_k),
              "Unexpected g_idx.size(0) = ", g_idx.size(0), " and perm.size(0) = ", perm.size(0),
              ", where size_k = ", size_k);
  const bool has_act_order = g_idx.size(0) != 0;
  TORCH_CHECK(!has_act_order || is_k_full,
              "act_order with a k-sharded weight (is_k_full = False) is not implemented in the B200 marlin kernel");
  TORCH_CHECK(b_scales.dim() == 2, "b_scales rank = ", b_scales.dim(), " is not 2");
  TORCH_CHECK(b_scales.size(1) == size_n, "b_scales dim 1 = ", b_scales.size(1), " is not size_n = ", size_n);
  const int64_t num_groups = b_scales.size(0);
  if (num_groups > 1)
    TORCH_CHECK(size_k % num_groups == 0, "size_k = ", size_k, ", is not divisible by b_scales.size(0) = ", num_groups);
  if (has_zp) {
    TORCH_CHECK(b_zeros.dim() == 2, "b_zeros rank = ", b_zeros.dim(), " is not 2");
    TORCH_CHECK(b_zeros.size(0) == num_groups, "b_zeros dim 0 = ", b_zeros.size(0), " is not num_groups = ", num_groups);
    if (is_zp_float) {
      TORCH_CHECK(b_zeros.size(1) == size_n, "b_zeros dim 1 = ", b_zeros.size(1), " is not size_n = ", size_n);
      TORCH_CHECK(b_zeros.scalar_type() == at::kHalf, "float zero points must be float16");
    } else {
      TORCH_CHECK(b_zeros.size(1) == size_n / pack_factor, "b_zeros dim 1 = ", b_zeros.size(1),
                  " is not size_n / pack_factor = ", size_n / pack_factor);
    }
    TORCH_CHECK(b_zeros.device().is_cuda() && b_zeros.is_contiguous(), "b_zeros must be a contiguous GPU tensor");
  }
  TORCH_CHECK(size_n % 64 == 0, "size_n = ", size_n, ", is not divisible by min_thread_n = 64");
  const int64_t min_workspace_size = (size_n / 64) * 16;
  TORCH_CHECK(workspace.numel() >= min_workspace_size, "workspace.numel = ", workspace.numel(),
              " is below min_workspace_size = ", min_workspace_size);
  TORCH_CHECK(a.scalar_type() == at::kHalf || a.scalar_type() == at::kBFloat16,
              "gpt_marlin_gemm only supports bfloat16 and float16");
  TORCH_CHECK(b_scales.scalar_type() == a.scalar_type(), "b_scales must have the dtype of a");

  const at::cuda::OptionalCUDAGuard guard(device_of(a));
  torch::Tensor c = torch::empty({size_m, size_n}, a.options());
  if (size_m == 0) return c;
  const int split = b200_marlin_gemm_plan((int)size_m, (int)size_n, (int)size_k, (int)num_groups);
  torch::Tensor c_tmp;
  float* c_tmp_ptr = nullptr;
  if (split > 1) {  // the reference's fp32 global-reduce buffer (gptq_marlin.cu:2317-2327): one slab per split
    c_tmp = torch::empty({split, size_m, size_n}, a.options().dtype(at::kFloat));
    c_tmp_ptr = c_tmp.data_ptr<float>();
  }
  TORCH_CHECK(workspace.scalar_type() == at::kInt, "workspace must be int32");
  (void)use_fp32_reduce;
  // act-order with the full k range: the checkpoint rows were sorted by group (gptq_marlin_repack with `perm`), so
  // gathering A's columns with the same permutation reduces the problem to the regular grouped GEMM — what the
  // reference does with its a_tmp buffer (gptq_marlin.cu:2145-2158, permute_cols_kernel)
  torch::Tensor a_perm;
  if (has_act_order) {
    TORCH_CHECK(perm.scalar_type() == at::kInt, "perm must be int32");
    a_perm = torch::empty_like(a);
    check(b200_permute_cols(a.data_ptr(), perm.data_ptr<int>(), a_perm.data_ptr(), size_m, (int)size_k, cur_stream()));
  }
  const torch::Tensor& a_used = has_act_order ? a_perm : a;
  check(b200_gptq_marlin_gemm(a_used.data_ptr(), b_q_weight.data_ptr(), b_scales.data_ptr(),
                              has_zp ? b_zeros.data_ptr() : nullptr, c.data_ptr(), c_tmp_ptr,
                              workspace.data_ptr<int>(), (int)size_m, (int)size_n, (int)size_k, (int)num_groups, (int)type_bits, has_zp ? (is_zp_float ? 2 : 1) : 0,
                              dtype_code(a, "gptq_marlin_gemm"), /*split_k=*/0, cur_stream()));
  return c;
}

void gptq_marlin_gemm_boxed(const c10::OperatorHandle&, torch::jit::Stack* stack) {
  auto args = torch::jit::last(*stack, 15);
  torch::Tensor a = args[0].toTensor(), bq = args[1].toTensor(), bs = args[2].toTensor(),
                bz = args[3].toTensor(), gi = args[4].toTensor(), pm = args[5].toTensor(),
                ws = args[6].toTensor();
  const c10::IValue st = args[7];
  const int64_t bits = scalar_type_prop(st, "size_bits"), bias = scalar_type_prop(st, "bias");
  const std::string name = "ScalarType(size_bits=" + std::to_string(bits) + ", bias=" + std::to_string(bias) + ")";
  torch::Tensor c = gptq_marlin_gemm_impl(a, bq, bs, bz, gi, pm, ws, bits, bias, name, args[8].toInt(),
                                          args[9].toInt(), args[10].toInt(), args[11].toBool(),
                                          args[12].toBool(), args[13].toBool(), args[14].toBool());
  torch::jit::drop(*stack, 15);
  torch::jit::push(*stack, std::move(c));
}

void moe_align_block_size(torch::Tensor topk_ids, int64_t num_experts, int64_t block_size,
                          torch::Tensor sorted_token_ids, torch::Tensor experts_ids,
                          torch::Tensor num_tokens_post_pad) {
  const at::cuda::OptionalCUDAGuard guard(device_of(topk_ids));
  TORCH_CHECK(topk_ids.scalar_type() == at::kInt || topk_ids.scalar_type() == at::kLong,
              "moe_align_block_size: topk_ids must be int32 or int64");
  TORCH_CHECK(topk_ids.is_contiguous(), "topk_ids must be contiguous");
  check(b200_moe_align_block_size(topk_ids.data_ptr(), topk_ids.scalar_type() == at::kLong ? 1 : 0,
                                  topk_ids.numel(), (int)num_experts, (int)block_size,
                                  sorted_token_ids.data_ptr<int32_t>(), experts_ids.data_ptr<int32_t>(),
                                  num_tokens_post_pad.data_ptr<int32_t>(), cur_stream()));
}

torch::Tensor permute_cols(torch::Tensor const& a, torch::Tensor const& perm) {
  const at::cuda::OptionalCUDAGuard guard(device_of(a));
  TORCH_CHECK(a.scalar_type() == at::kHalf || a.scalar_type() == at::kBFloat16, "Currently only 16bit types are supported");
  TORCH_CHECK(a.is_contiguous(), "A must be contiguous");
  TORCH_CHECK(a.size(-1) % 8 == 0, "A columns must be a multiple of 8 (128bits)");
  TORCH_CHECK(perm.scalar_type() == at::kInt && perm.numel() == a.size(-1), "perm must be int32 [size_k]");
  torch::Tensor out = torch::empty_like(a);
  check(b200_permute_cols(a.data_ptr(), perm.data_ptr<int>(), out.data_ptr(), a.numel() / a.size(-1),
                          (int)a.size(-1), cur_stream()));
  return out;
}

torch::Tensor awq_dequantize(torch::Tensor kernel, torch::Tensor scaling_factors, torch::Tensor zeros,
                             int64_t split_k_iters, int64_t thx, int64_t thy) {
  (void)split_k_iters; (void)thx; (void)thy;   // launch-shape hints of the reference kernel; not needed here
  TORCH_CHECK(kernel.dim() == 2, "awq_dequantize: only 2-D qweight is supported");
  TORCH_CHECK(scaling_factors.scalar_type() == at::kHalf, "awq_dequantize: scales must be float16");
  const int64_t in_c = kernel.size(0), qout_c = kernel.size(1);
  const int64_t G = in_c / scaling_factors.size(0);
  const at::cuda::OptionalCUDAGuard guard(device_of(scaling_factors));
  torch::Tensor out = torch::empty({in_c, qout_c * 8}, scaling_factors.options());
  check(b200_awq_dequantize(kernel.data_ptr(), scaling_factors.data_ptr(), zeros.data_ptr(), out.data_ptr(), in_c,
                            (int)qout_c, (int)G, cur_stream()));
  return out;
}

void advance_step_flashattn(int64_t num_seqs, int64_t num_queries, int64_t block_size, torch::Tensor& input_tokens,
                            torch::Tensor& sampled_token_ids, torch::Tensor& input_positions,
                            torch::Tensor& seq_lens, torch::Tensor& slot_mapping, torch::Tensor& block_tables) {
  TORCH_CHECK(input_tokens.scalar_type() == at::kLong && sampled_token_ids.scalar_type() == at::kLong &&
              input_positions.scalar_type() == at::kLong && slot_mapping.scalar_type() == at::kLong,
              "input_tokens / sampled_token_ids / input_positions / slot_mapping must be int64");
  TORCH_CHECK(seq_lens.scalar_type() == at::kInt && block_tables.scalar_type() == at::kInt,
              "seq_lens / block_tables must be int32");
  TORCH_CHECK(input_tokens.size(0) == num_seqs && sampled_token_ids.size(0) == num_queries &&
              input_positions.size(0) == num_seqs && seq_lens.size(0) == num_seqs &&
              slot_mapping.size(0) == num_seqs && block_tables.size(0) == num_seqs, "advance_step: bad tensor sizes");
  const at::cuda::OptionalCUDAGuard guard(device_of(input_tokens));
  check(b200_advance_step_flashattn((int)num_seqs, (int)num_queries, (int)block_size, input_tokens.data_ptr<int64_t>(),
                                    sampled_token_ids.data_ptr<int64_t>(), input_positions.data_ptr<int64_t>(),
                                    seq_lens.data_ptr<int>(), slot_mapping.data_ptr<int64_t>(),
                                    block_tables.data_ptr<int>(), block_tables.stride(0), cur_stream()));
}

torch::Tensor gptq_marlin_repack(torch::Tensor& b_q_weight, torch::Tensor& perm, c10::SymInt size_k_s,
                                 c10::SymInt size_n_s, int64_t num_bits) {
  const int64_t size_k = size_k_s.expect_int(), size_n = size_n_s.expect_int();
  TORCH_CHECK(num_bits == 4 || num_bits == 8, "num_bits must be 4 or 8. Got = ", num_bits);
  const int64_t pf = 32 / num_bits;
  TORCH_CHECK(size_k % 16 == 0, "size_k = ", size_k, " is not divisible by tile_k_size = 16");
  TORCH_CHECK(size_n % 64 == 0, "size_n = ", size_n, " is not divisible by tile_n_size = 64");
  TORCH_CHECK(b_q_weight.size(0) == size_k / pf, "Shape mismatch: b_q_weight.size(0) = ", b_q_weight.size(0),
              ", size_k = ", size_k, ", pack_factor = ", pf);
  TORCH_CHECK(b_q_weight.size(1) == size_n, "b_q_weight.size(1) = ", b_q_weight.size(1), " is not size_n = ", size_n);
  TORCH_CHECK(b_q_weight.device().is_cuda() && b_q_weight.is_contiguous(), "b_q_weight must be a contiguous GPU tensor");
  TORCH_CHECK(b_q_weight.dtype() == at::kInt, "b_q_weight type is not kInt");
  TORCH_CHECK(perm.device().is_cuda() && perm.is_contiguous(), "perm must be a contiguous GPU tensor");
  TORCH_CHECK(perm.dtype() == at::kInt, "perm type is not at::kInt");
  TORCH_CHECK(perm.numel() == 0 || perm.numel() == size_k, "perm must be empty or have size_k entries");
  const at::cuda::OptionalCUDAGuard guard(device_of(b_q_weight));
  torch::Tensor out = torch::empty({size_k / 16, size_n * 16 / pf}, b_q_weight.options());
  check(b200_gptq_marlin_repack(b_q_weight.data_ptr(), perm.numel() ? perm.data_ptr<int>() : nullptr,
                                out.data_ptr(), (int)size_k, (int)size_n, (int)num_bits, cur_stream()));
  return out;
}
torch::Tensor gptq_marlin_repack_meta(torch::Tensor& b_q_weight, torch::Tensor& perm, c10::SymInt size_k,
                                      c10::SymInt size_n, int64_t num_bits) {
  const int64_t pf = 32 / num_bits;
  return torch::empty_symint({size_k / 16, size_n * 16 / pf}, b_q_weight.options());
}

torch::Tensor awq_marlin_repack(torch::Tensor& b_q_weight, c10::SymInt size_k_s, c10::SymInt size_n_s,
                                int64_t num_bits) {
  const int64_t size_k = size_k_s.expect_int(), size_n = size_n_s.expect_int();
  TORCH_CHECK(num_bits == 4 || num_bits == 8, "num_bits must be 4 or 8. Got = ", num_bits);
  const int64_t pf = 32 / num_bits;
  TORCH_CHECK(size_k % 16 == 0, "size_k = ", size_k, " is not divisible by tile_k_size = 16");
  TORCH_CHECK(size_n % 64 == 0, "size_n = ", size_n, " is not divisible by tile_n_size = 64");
  TORCH_CHECK(b_q_weight.size(0) == size_k, "b_q_weight.size(0) = ", b_q_weight.size(0), " is not size_k = ", size_k);
  TORCH_CHECK(b_q_weight.size(1) == size_n / pf, "Shape mismatch: b_q_weight.size(1) = ", b_q_weight.size(1),
              ", size_n = ", size_n, ", pack_factor = ", pf);
  TORCH_CHECK(b_q_weight.device().is_cuda() && b_q_weight.is_contiguous(), "b_q_weight must be a contiguous GPU tensor");
  TORCH_CHECK(b_q_weight.dtype() == at::kInt, "b_q_weight type is not kInt");
  const at::cuda::OptionalCUDAGuard guard(device_of(b_q_weight));
  torch::Tensor out = torch::empty({size_k / 16, size_n * 16 / pf}, b_q_weight.options());
  check(b200_awq_marlin_repack(b_q_weight.data_ptr(), out.data_ptr(), (int)size_k, (int)size_n, (int)num_bits,
                               cur_stream()));
  return out;
}
torch::Tensor awq_marlin_repack_meta(torch::Tensor& b_q_weight, c10::SymInt size_k, c10::SymInt size_n,
                                     int64_t num_bits) {
  const int64_t pf = 32 / num_bits;
  return torch::empty_symint({size_k / 16, size_n * 16 / pf}, b_q_weight.options());
}

// ---- custom all-reduce (kernels/all_reduce/custom_all_reduce.cu:15-141) ---------------------------------------
using fptr_t = int64_t;

std::string pack_handles(const std::vector<std::string>& handles) {
  std::string blob;
  for (const auto& h : handles) {
    // torch <= 2.4 (the reference's pin) shares the raw 64-byte cudaIpcMemHandle_t; newer caching allocators
    // prefix it with {version, type}: type 'c' = cudaMalloc block (usable), 'e' = expandable segment (not IPC-able)
    if (h.size() == 66) {
      TORCH_CHECK(h[1] == 'c', "custom allreduce needs cudaMalloc-backed buffers (expandable_segments is not supported)");
      blob += h.substr(2);
    } else {
      TORCH_CHECK(h.size() == 64, "IPC handle must be 64 bytes, got ", h.size());
      blob += h;
    }
  }
  return blob;
}

fptr_t init_custom_ar(torch::Tensor& meta, torch::Tensor& rank_data, const std::vector<std::string>& handles,
                      const std::vector<int64_t>& offsets, int64_t rank, bool full_nvlink) {
  const int world_size = (int)offsets.size();
  if (world_size > 8) throw std::invalid_argument("world size > 8 is not supported");
  if (world_size % 2 != 0) throw std::invalid_argument("Odd num gpus is not supported for now");
  if (world_size != (int)handles.size()) throw std::invalid_argument("handles length should equal to offsets length");
  if (rank < 0 || rank >= world_size) throw std::invalid_argument("invalid rank passed in");
  const at::cuda::OptionalCUDAGuard guard(device_of(meta));
  const std::string blob = pack_handles(handles);
  const fptr_t fa = b200_car_init(meta.data_ptr(), rank_data.data_ptr(), rank_data.numel() * rank_data.element_size(),
                                  blob.data(), offsets.data(), world_size, (int)rank, full_nvlink ? 1 : 0);
  TORCH_CHECK(fa != 0, b200_last_error());
  return fa;
}

bool is_weak_contiguous(const torch::Tensor& t) {
  return t.is_contiguous() || (t.storage().nbytes() - t.storage_offset() * t.element_size() ==
                               (size_t)(t.numel() * t.element_size()));
}

void all_reduce_reg(fptr_t fa, torch::Tensor& inp, torch::Tensor& out) {
  const at::cuda::OptionalCUDAGuard guard(device_of(inp));
  TORCH_CHECK_EQ(inp.scalar_type(), out.scalar_type());
  TORCH_CHECK_EQ(inp.numel(), out.numel());
  TORCH_CHECK(is_weak_contiguous(out));
  check(b200_car_all_reduce(fa, inp.data_ptr(), out.data_ptr(), out.numel(), dtype_code(out, "custom allreduce"),
                            cur_stream()));
}

void all_reduce_unreg(fptr_t fa, torch::Tensor& inp, torch::Tensor& reg_buffer, torch::Tensor& out) {
  const at::cuda::OptionalCUDAGuard guard(device_of(inp));
  const size_t input_size = inp.numel() * inp.element_size();
  TORCH_CHECK_EQ(inp.scalar_type(), out.scalar_type());
  TORCH_CHECK_EQ(inp.numel(), out.numel());
  TORCH_CHECK(input_size <= (size_t)(reg_buffer.numel() * reg_buffer.element_size()),
              "registered buffer is too small to contain the input");
  TORCH_CHECK(is_weak_contiguous(out));
  AT_CUDA_CHECK(cudaMemcpyAsync(reg_buffer.data_ptr(), inp.data_ptr(), input_size, cudaMemcpyDeviceToDevice,
                                (cudaStream_t)cur_stream()));
  check(b200_car_all_reduce(fa, reg_buffer.data_ptr(), out.data_ptr(), out.numel(),
                            dtype_code(out, "custom allreduce"), cur_stream()));
}

void dispose(fptr_t fa) { b200_car_dispose(fa); }
int64_t meta_size() { return b200_car_meta_size(); }

void register_buffer(fptr_t fa, torch::Tensor& t, const std::vector<std::string>& handles,
                     const std::vector<int64_t>& offsets) {
  const at::cuda::OptionalCUDAGuard guard(device_of(t));
  const std::string blob = pack_handles(handles);
  check(b200_car_register_buffer(fa, t.data_ptr(), blob.data(), offsets.data()));
}

std::tuple<torch::Tensor, std::vector<int64_t>> get_graph_buffer_ipc_meta(fptr_t fa) {
  const int n = b200_car_get_graph_buffer_ipc_meta(fa, nullptr, nullptr, 0);
  TORCH_CHECK(n >= 0, b200_last_error());
  auto handles = torch::empty({(int64_t)n * 64}, torch::TensorOptions().dtype(torch::kUInt8).device(torch::kCPU));
  std::vector<int64_t> offsets(n);
  if (n > 0)
    TORCH_CHECK(b200_car_get_graph_buffer_ipc_meta(fa, handles.data_ptr(), offsets.data(), n) == n, b200_last_error());
  return {handles, std::move(offsets)};
}

void register_graph_buffers(fptr_t fa, const std::vector<std::string>& handles,
                            const std::vector<std::vector<int64_t>>& offsets) {
  const size_t ws = handles.size();
  TORCH_CHECK(ws == offsets.size() && ws > 0, "handles / offsets length mismatch");
  const size_t n = offsets[0].size();
  std::string blob;
  std::vector<int64_t> flat;
  for (size_t r = 0; r < ws; ++r) {
    TORCH_CHECK(handles[r].size() == n * 64 && offsets[r].size() == n, "per-rank graph buffer meta size mismatch");
    blob += handles[r];
    flat.insert(flat.end(), offsets[r].begin(), offsets[r].end());
  }
  check(b200_car_register_graph_buffers(fa, blob.data(), flat.data(), (int)n));
}

// ---- extensions beyond the reference's op set (namespace _C_b200) --------------------------------------------
// Fused TP exchange (csrc/tp_fused.cu). `block` is this rank's symmetric allocation; x / out are views into it.
void tp_allreduce_rows(int64_t mc_base, torch::Tensor& block, const std::vector<int64_t>& peer_bases,
                       torch::Tensor& x, torch::Tensor& out, const c10::optional<torch::Tensor>& residual,
                       const c10::optional<torch::Tensor>& weight, double epsilon, int64_t flag_off, int64_t rank,
                       int64_t world, int64_t algo) {
  const at::cuda::OptionalCUDAGuard guard(device_of(block));
  TORCH_CHECK(x.dim() == 2 && x.is_contiguous() && out.is_contiguous() && x.sizes() == out.sizes() &&
              x.scalar_type() == out.scalar_type(), "tp_allreduce_rows: x / out must be contiguous [T, H] of one dtype");
  char* base = static_cast<char*>(block.data_ptr());
  const int64_t nbytes = block.numel() * block.element_size();
  const int64_t in_off = static_cast<char*>(x.data_ptr()) - base, out_off = static_cast<char*>(out.data_ptr()) - base;
  const int64_t bytes = x.numel() * x.element_size();
  TORCH_CHECK(in_off >= 0 && in_off + bytes <= nbytes && out_off >= 0 && out_off + bytes <= nbytes,
              "tp_allreduce_rows: x / out must live inside the symmetric block");
  TORCH_CHECK((int64_t)peer_bases.size() == world, "tp_allreduce_rows: one peer base per rank");
  TORCH_CHECK(residual.has_value() == weight.has_value(), "tp_allreduce_rows: residual and weight go together");
  if (residual) {
    TORCH_CHECK(residual->is_contiguous() && residual->sizes() == x.sizes() && residual->scalar_type() == x.scalar_type() &&
                weight->is_contiguous() && weight->numel() == x.size(1) && weight->scalar_type() == x.scalar_type(),
                "tp_allreduce_rows: residual [T, H] / weight [H] of x's dtype");
  }
  check(b200_tp_allreduce_rows(reinterpret_cast<void*>(mc_base), base, peer_bases.data(), in_off, out_off, flag_off,
                               residual ? residual->data_ptr() : nullptr, weight ? weight->data_ptr() : nullptr,
                               (float)epsilon, (int)x.size(0), (int)x.size(1), (int)rank, (int)world,
                               dtype_code(x, "tp_allreduce_rows"), (int)algo, cur_stream()));
}
// rotary_embedding + reshape_and_cache in one launch (csrc/norm_rope_act.cu). query [T, Hq*D], key / value [T, Hkv*D]
// (views of the qkv GEMM output); caches in the reference's paged layout.
void rotary_embedding_and_cache(torch::Tensor& positions, torch::Tensor& query, torch::Tensor& key,
                                torch::Tensor& value, int64_t head_size, torch::Tensor& cos_sin_cache, bool is_neox,
                                torch::Tensor& key_cache, torch::Tensor& value_cache, torch::Tensor& slot_mapping,
                                const std::string& kv_cache_dtype, double k_scale, double v_scale) {
  const at::cuda::OptionalCUDAGuard guard(device_of(query));
  const int64_t num_tokens = query.numel() / query.size(-1);
  TORCH_CHECK(key.size(-1) == value.size(-1) && key.scalar_type() == query.scalar_type() &&
              value.scalar_type() == query.scalar_type(), "rotary_embedding_and_cache: q / k / v mismatch");
  TORCH_CHECK(query.stride(-1) == 1 && key.stride(-1) == 1 && value.stride(-1) == 1,
              "rotary_embedding_and_cache: rows must be contiguous");
  check(b200_rotary_embedding_and_cache(
      positions.data_ptr<int64_t>(), query.data_ptr(), key.data_ptr(), value.data_ptr(), cos_sin_cache.data_ptr(),
      key_cache.data_ptr(), value_cache.data_ptr(), slot_mapping.data_ptr<int64_t>(), (int)num_tokens,
      (int)(query.size(-1) / head_size), (int)(key.size(-1) / head_size), (int)head_size, (int)cos_sin_cache.size(1),
      query.stride(-2), key.stride(-2), value.stride(-2), is_neox ? 1 : 0, (int)key_cache.size(3),
      (int)key_cache.size(4), dtype_code(query, "rotary_embedding_and_cache"), kv_code(kv_cache_dtype), (float)k_scale,
      (float)v_scale, cur_stream()));
}
void moe_expert_scale_add(torch::Tensor& final_out, torch::Tensor& cur, torch::Tensor& topk_weights,
                          torch::Tensor& topk_ids, int64_t expert, bool first) {
  const at::cuda::OptionalCUDAGuard guard(device_of(cur));
  TORCH_CHECK(final_out.is_contiguous() && cur.is_contiguous() && final_out.sizes() == cur.sizes() && cur.dim() == 2 &&
              final_out.scalar_type() == cur.scalar_type(), "moe_expert_scale_add: final / cur must be contiguous [T, H]");
  TORCH_CHECK(topk_weights.scalar_type() == at::kFloat && topk_ids.scalar_type() == at::kInt &&
              topk_weights.is_contiguous() && topk_ids.is_contiguous() && topk_weights.sizes() == topk_ids.sizes() &&
              topk_ids.size(0) == cur.size(0), "moe_expert_scale_add: topk_weights f32 / topk_ids int32 [T, topk]");
  check(b200_moe_expert_scale_add(final_out.data_ptr(), cur.data_ptr(), topk_weights.data_ptr<float>(),
                                  topk_ids.data_ptr<int>(), (int)cur.size(0), (int)cur.size(1), (int)topk_ids.size(1),
                                  (int)expert, first ? 1 : 0, dtype_code(cur, "moe_expert_scale_add"), cur_stream()));
}
int64_t tp_flag_bytes() { return b200_tp_flag_bytes(); }

}  // namespace

TORCH_LIBRARY(_C, ops) {
  ops.def(
      "paged_attention_v1("
      "    Tensor! out, Tensor query, Tensor key_cache,"
      "    Tensor value_cache, int num_kv_heads, float scale,"
      "    Tensor block_tables, Tensor seq_lens, int block_size,"
      "    int max_seq_len, Tensor? alibi_slopes,"
      "    str kv_cache_dtype, float k_scale, float v_scale,"
      "    int tp_rank, int blocksparse_local_blocks,"
      "    int blocksparse_vert_stride, int blocksparse_block_size,"
      "    int blocksparse_head_sliding_step) -> ()");
  ops.impl("paged_attention_v1", torch::kCUDA, &paged_attention_v1);
  ops.def(
      "paged_attention_v2("
      "    Tensor! out, Tensor! exp_sums, Tensor! max_logits,"
      "    Tensor! tmp_out, Tensor query, Tensor key_cache,"
      "    Tensor value_cache, int num_kv_heads, float scale,"
      "    Tensor block_tables, Tensor seq_lens, int block_size,"
      "    int max_seq_len, Tensor? alibi_slopes,"
      "    str kv_cache_dtype, float k_scale, float v_scale,"
      "    int tp_rank, int blocksparse_local_blocks,"
      "    int blocksparse_vert_stride, int blocksparse_block_size,"
      "    int blocksparse_head_sliding_step) -> ()");
  ops.impl("paged_attention_v2", torch::kCUDA, &paged_attention_v2);

  ops.def("silu_and_mul(Tensor! out, Tensor input) -> ()");
  ops.impl("silu_and_mul", torch::kCUDA, &act_and_mul<0>);
  ops.def("gelu_and_mul(Tensor! out, Tensor input) -> ()");
  ops.impl("gelu_and_mul", torch::kCUDA, &act_and_mul<1>);
  ops.def("gelu_tanh_and_mul(Tensor! out, Tensor input) -> ()");
  ops.impl("gelu_tanh_and_mul", torch::kCUDA, &act_and_mul<2>);
  ops.def("gelu_new(Tensor! out, Tensor input) -> ()");
  ops.impl("gelu_new", torch::kCUDA, &activation<0>);
  ops.def("gelu_fast(Tensor! out, Tensor input) -> ()");
  ops.impl("gelu_fast", torch::kCUDA, &activation<1>);
  ops.def("gelu_quick(Tensor! out, Tensor input) -> ()");
  ops.impl("gelu_quick", torch::kCUDA, &activation<2>);

  ops.def("rms_norm(Tensor! out, Tensor input, Tensor weight, float epsilon) -> ()");
  ops.impl("rms_norm", torch::kCUDA, &rms_norm);
  ops.def(
      "fused_add_rms_norm(Tensor! input, Tensor! residual, Tensor weight, "
      "float epsilon) -> ()");
  ops.impl("fused_add_rms_norm", torch::kCUDA, &fused_add_rms_norm);

  ops.def(
      "rotary_embedding(Tensor positions, Tensor! query,"
      "                 Tensor! key, int head_size,"
      "                 Tensor cos_sin_cache, bool is_neox) -> ()");
  ops.impl("rotary_embedding", torch::kCUDA, &rotary_embedding);
  ops.def(
      "batched_rotary_embedding(Tensor positions, Tensor! query,"
      "                         Tensor! key, int head_size,"
      "                         Tensor cos_sin_cache, bool is_neox,"
      "                         int rot_dim,"
      "                         Tensor cos_sin_cache_offsets) -> ()");
  ops.impl("batched_rotary_embedding", torch::kCUDA, &batched_rotary_embedding);

  // Aligning the number of tokens to be processed by each expert (kernels/torch_bindings.cpp:394-399)
  ops.def(
      "moe_align_block_size(Tensor topk_ids, int num_experts,"
      "                     int block_size, Tensor! sorted_token_ids,"
      "                     Tensor! experts_ids,"
      "                     Tensor! num_tokens_post_pad) -> ()");
  ops.impl("moe_align_block_size", torch::kCUDA, &moe_align_block_size);

  // prepare_inputs advance_step (kernels/torch_bindings.cpp:77-82)
  ops.def(
      "advance_step_flashattn(int num_seqs, int num_queries, int block_size, "
      "Tensor! input_tokens, Tensor sampled_token_ids, "
      "Tensor! input_positions, Tensor! seq_lens, Tensor! slot_mapping, "
      "Tensor block_tables) -> ()");
  ops.impl("advance_step_flashattn", torch::kCUDA, &advance_step_flashattn);

  // AWQ dequantisation (kernels/torch_bindings.cpp:147-151) and act-order column gather (:218-219)
  ops.def(
      "awq_dequantize(Tensor _kernel, Tensor _scaling_factors, "
      "Tensor _zeros, int split_k_iters, int thx, int thy) -> Tensor");
  ops.impl("awq_dequantize", torch::kCUDA, &awq_dequantize);
  ops.def("permute_cols(Tensor A, Tensor perm) -> Tensor");
  ops.impl("permute_cols", torch::kCUDA, &permute_cols);

  // Marlin-format weight-only quantised GEMM + repack (kernels/torch_bindings.cpp:195-215)
  ops.def(
      "gptq_marlin_gemm(Tensor a, Tensor b_q_weight, Tensor b_scales, "
      "Tensor b_zeros, Tensor g_idx, Tensor perm, Tensor workspace, "
      "__torch__.torch.classes._core_C.ScalarType b_q_type, "
      "int size_m, int size_n, int size_k, bool is_k_full, "
      "bool has_zp, bool use_fp32_reduce, bool is_zp_float) -> Tensor");
  ops.impl("gptq_marlin_gemm", torch::dispatch(c10::DispatchKey::CUDA,
           torch::CppFunction::makeFromBoxedFunction<&gptq_marlin_gemm_boxed>()));
  ops.def(
      "gptq_marlin_repack(Tensor b_q_weight, Tensor perm, "
      "SymInt size_k, SymInt size_n, int num_bits) -> Tensor");
  ops.impl("gptq_marlin_repack", torch::kCUDA, &gptq_marlin_repack);
  ops.impl("gptq_marlin_repack", torch::kMeta, &gptq_marlin_repack_meta);
  ops.def(
      "awq_marlin_repack(Tensor b_q_weight, SymInt size_k, "
      "SymInt size_n, int num_bits) -> Tensor");
  ops.impl("awq_marlin_repack", torch::kCUDA, &awq_marlin_repack);
  ops.impl("awq_marlin_repack", torch::kMeta, &awq_marlin_repack_meta);
}

TORCH_LIBRARY(_C_cache_ops, cache_ops) {
  cache_ops.def("swap_blocks(Tensor src, Tensor! dst, Tensor block_mapping) -> ()");
  cache_ops.impl("swap_blocks", torch::kCUDA, &swap_blocks);
  cache_ops.def(
      "copy_blocks(Tensor(a!)[] key_caches, Tensor[](b!) value_caches, "
      "Tensor block_mapping) -> ()");
  cache_ops.impl("copy_blocks", torch::kCUDA, &copy_blocks);
  cache_ops.def(
      "reshape_and_cache(Tensor key, Tensor value,"
      "                  Tensor! key_cache, Tensor! value_cache,"
      "                  Tensor slot_mapping,"
      "                  str kv_cache_dtype,"
      "                  float k_scale, float v_scale) -> ()");
  cache_ops.impl("reshape_and_cache", torch::kCUDA, &reshape_and_cache);
  cache_ops.def(
      "reshape_and_cache_flash(Tensor key, Tensor value,"
      "                        Tensor! key_cache,"
      "                        Tensor! value_cache,"
      "                        Tensor slot_mapping,"
      "                        str kv_cache_dtype,"
      "                        float k_scale, float v_scale) -> ()");
  cache_ops.impl("reshape_and_cache_flash", torch::kCUDA, &reshape_and_cache_flash);
  cache_ops.def(
      "convert_fp8(Tensor! dst_cache, Tensor src_cache, float scale, "
      "str kv_cache_dtype) -> ()");
  cache_ops.impl("convert_fp8", torch::kCUDA, &convert_fp8);
}

TORCH_LIBRARY(_C_cuda_utils, cuda_utils) {
  cuda_utils.def("get_device_attribute(int attribute, int device_id) -> int");
  cuda_utils.impl("get_device_attribute", &get_device_attribute);
  cuda_utils.def("get_max_shared_memory_per_block_device_attribute(int device_id) -> int");
  cuda_utils.impl("get_max_shared_memory_per_block_device_attribute",
                  &get_max_shared_memory_per_block_device_attribute);
}

TORCH_LIBRARY(_C_custom_ar, custom_ar) {
  custom_ar.def(
      "init_custom_ar(Tensor meta, Tensor rank_data, "
      "str[] handles, int[] offsets, int rank, "
      "bool full_nvlink) -> int");
  custom_ar.impl("init_custom_ar", torch::kCUDA, &init_custom_ar);
  custom_ar.def("all_reduce_reg(int fa, Tensor inp, Tensor! out) -> ()");
  custom_ar.impl("all_reduce_reg", torch::kCUDA, &all_reduce_reg);
  custom_ar.def("all_reduce_unreg(int fa, Tensor inp, Tensor reg_buffer, Tensor! out) -> ()");
  custom_ar.impl("all_reduce_unreg", torch::kCUDA, &all_reduce_unreg);
  custom_ar.def("dispose", &dispose);
  custom_ar.def("meta_size", &meta_size);
  custom_ar.def(
      "register_buffer(int fa, Tensor t, str[] handles, "
      "int[] offsets) -> ()");
  custom_ar.impl("register_buffer", torch::kCUDA, &register_buffer);
  custom_ar.def("get_graph_buffer_ipc_meta", &get_graph_buffer_ipc_meta);
  custom_ar.def("register_graph_buffers", &register_graph_buffers);
}

// This repo's own additions (not part of the reference's op set; see INTEGRATION.md "extensions")
TORCH_LIBRARY(_C_b200, ext) {
  ext.def(
      "tp_allreduce_rows(int mc_base, Tensor block, int[] peer_bases, Tensor x, Tensor! out, "
      "Tensor!? residual, Tensor? weight, float epsilon, int flag_off, int rank, int world, int algo) -> ()");
  ext.impl("tp_allreduce_rows", torch::kCUDA, &tp_allreduce_rows);
  ext.def("tp_flag_bytes", &tp_flag_bytes);
  ext.def("moe_expert_scale_add(Tensor! final_out, Tensor cur, Tensor topk_weights, Tensor topk_ids, int expert, "
          "bool first) -> ()");
  ext.impl("moe_expert_scale_add", torch::kCUDA, &moe_expert_scale_add);
  ext.def(
      "rotary_embedding_and_cache(Tensor positions, Tensor! query, Tensor! key, Tensor value, int head_size, "
      "Tensor cos_sin_cache, bool is_neox, Tensor! key_cache, Tensor! value_cache, Tensor slot_mapping, "
      "str kv_cache_dtype, float k_scale, float v_scale) -> ()");
  ext.impl("rotary_embedding_and_cache", torch::kCUDA, &rotary_embedding_and_cache);
}

// `import <pkg>._C` support (kernels/core/registration.h:22-27 of the reference does the same)
PyMODINIT_FUNC PyInit__C() {
  static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_C", nullptr, 0, nullptr};
  return PyModule_Create(&module);
}
