"""Host-side entry points of the decode hot path, under the names the reference's callers use.

The reference reaches its kernels through thin Python functions in `aphrodite/_custom_ops.py` (activations :46-68,
paged attention :72-131, rotary :158-178, norms :182-189, advance_step / awq / permute :192-232 and :627-628, Marlin
:554-600, MoE :830-843, cache ops :846-892, device queries :895-902, custom all-reduce :906-941). A caller that does
`from aphrodite import _custom_ops as ops; ops.paged_attention_v1(...)` must be able to do the same against this
package, so every function below keeps the reference's NAME, PARAMETER ORDER, parameter names and defaults; nothing
else is shared. The functions are generated from one table (`_TABLE`): each entry names the torch op namespace the call
forwards to and spells the parameter list once. All of them forward to ops that this package registers from its own
`_C.abi3.so` / `_moe_C.abi3.so` (csrc/torch_shim.cpp, csrc/moe_shim.cpp -> C ABI -> hand-written sm_100a kernels).
Importing this module loads the native libraries and raises `NativeLibraryMissing` if they have not been built: there is
no fallback implementation.
"""
import torch

from . import _native

_native.load_torch_ops()

# (function name, torch.ops namespace, parameter list exactly as a caller may spell it)
_ATTN_TAIL = ("num_kv_heads, scale, block_tables, seq_lens, block_size, max_seq_len, alibi_slopes, kv_cache_dtype, "
              "k_scale, v_scale, tp_rank=0, blocksparse_local_blocks=0, blocksparse_vert_stride=0, "
              "blocksparse_block_size=64, blocksparse_head_sliding_step=0")
_KV_WRITE = "key, value, key_cache, value_cache, slot_mapping, kv_cache_dtype, k_scale, v_scale"
_TABLE = (
    # gated / plain activations: out[T, d] <- input[T, 2d] resp. [T, d]
    ("silu_and_mul", "_C", "out, x"),
    ("gelu_and_mul", "_C", "out, x"),
    ("gelu_tanh_and_mul", "_C", "out, x"),
    ("gelu_fast", "_C", "out, x"),
    ("gelu_new", "_C", "out, x"),
    ("gelu_quick", "_C", "out, x"),
    # decode attention over the paged KV cache (v2 = 512-token partitions + reduce)
    ("paged_attention_v1", "_C", "out, query, key_cache, value_cache, " + _ATTN_TAIL),
    ("paged_attention_v2", "_C", "out, exp_sum, max_logits, tmp_out, query, key_cache, value_cache, " + _ATTN_TAIL),
    # in-place rotary position embedding (batched = per-token cache offsets, LoRA long-context)
    ("rotary_embedding", "_C", "positions, query, key, head_size, cos_sin_cache, is_neox"),
    ("batched_rotary_embedding", "_C",
     "positions, query, key, head_size, cos_sin_cache, is_neox, rot_dim, cos_sin_cache_offsets"),
    ("rms_norm", "_C", "out, input, weight, epsilon"),
    ("fused_add_rms_norm", "_C", "input, residual, weight, epsilon"),
    # multi-step decode input preparation, AWQ dequant, column gather
    ("advance_step_flashattn", "_C",
     "num_seqs, num_queries, block_size, input_tokens, sampled_token_ids, input_positions, seq_lens, slot_mapping, "
     "block_tables"),
    ("awq_dequantize", "_C", "qweight, scales, zeros, split_k_iters, thx, thy"),
    ("permute_cols", "_C", "a, perm"),
    # Marlin-format weight-only quantised GEMM and checkpoint re-tiling
    ("gptq_marlin_repack", "_C", "b_q_weight, perm, size_k, size_n, num_bits"),
    ("awq_marlin_repack", "_C", "b_q_weight, size_k, size_n, num_bits"),
    ("gptq_marlin_gemm", "_C",
     "a, b_q_weight, b_scales, b_zeros, g_idx, perm, workspace, b_q_type, size_m, size_n, size_k, is_k_full, "
     "has_zp=False, use_fp32_reduce=False, is_zp_float=False"),
    # mixture-of-experts routing
    ("moe_align_block_size", "_C",
     "topk_ids, num_experts, block_size, sorted_token_ids, experts_ids, num_tokens_post_pad"),
    ("topk_softmax", "_moe_C", "topk_weights, topk_ids, token_expert_indicies, gating_output"),
    # KV-cache writers and movers
    ("reshape_and_cache", "_C_cache_ops", _KV_WRITE),
    ("reshape_and_cache_flash", "_C_cache_ops", _KV_WRITE),
    ("copy_blocks", "_C_cache_ops", "key_caches, value_caches, block_mapping"),
    ("swap_blocks", "_C_cache_ops", "src, dst, block_mapping"),
    ("convert_fp8", "_C_cache_ops", "output, input, scale=1.0, kv_dtype='fp8'"),
    ("get_device_attribute", "_C_cuda_utils", "attribute, device"),
    ("get_max_shared_memory_per_block_device_attribute", "_C_cuda_utils", "device"),
    # sampler: probability-space kernels (reference _custom_ops.py: sampling_from_probs ... top_k_mask_logits)
    ("sampling_from_probs", "_C", "probs, uniform_samples, deterministic=True"),
    ("top_k_renorm_prob", "_C", "probs, maybe_top_k_arr, top_k_val"),
    ("top_p_renorm_prob", "_C", "probs, maybe_top_p_arr, top_p_val"),
    ("top_k_mask_logits", "_C", "logits, maybe_top_k_arr, top_k_val"),
    ("cutlass_scaled_mm_supports_fp8", "_C", "cuda_device_capability"),
    # NVLink peer-memory all-reduce (opaque handle `fa`)
    ("init_custom_ar", "_C_custom_ar", "meta, rank_data, handles, offsets, rank, full_nvlink"),
    ("all_reduce_reg", "_C_custom_ar", "fa, inp, out"),
    ("all_reduce_unreg", "_C_custom_ar", "fa, inp, reg_buffer, out"),
    ("dispose", "_C_custom_ar", "fa"),
    ("meta_size", "_C_custom_ar", ""),
    ("register_buffer", "_C_custom_ar", "fa, t, handles, offsets"),
    ("get_graph_buffer_ipc_meta", "_C_custom_ar", "fa"),
    ("register_graph_buffers", "_C_custom_ar", "fa, handles, offsets"),
)


def _emit(name: str, namespace: str, params: str):
    """Builds `def name(params): return torch.ops.<namespace>.<name>(<params, positionally>)` with a real signature,
    so keyword calls and the reference's defaults work; the op handle is looked up once."""
    op = getattr(getattr(torch.ops, namespace), name)
    forwarded = ", ".join(p.split("=")[0].strip() for p in params.split(",") if p.strip())
    scope = {"_op": op}
    exec(f"def {name}({params}):\n    return _op({forwarded})\n", scope)    # noqa: S102 - table above is the only input
    fn = scope[name]
    fn.__module__ = __name__
    fn.__doc__ = f"torch.ops.{namespace}.{name}({forwarded}) — this package's sm_100a implementation."
    return fn


for _name, _ns, _params in _TABLE:
    globals()[_name] = _emit(_name, _ns, _params)
del _name, _ns, _params


# ---- wrappers that are more than a forward in the reference too (they allocate outputs / unpack pairs) ---------------
def top_k_sampling_from_probs(probs, uniform_samples, maybe_top_k_arr, top_k_val, deterministic=True):
    """(samples int32 [B], success bool [B]); uniform_samples [max_rounds, B] (reference _custom_ops.py + sampling.cu:125)."""
    return tuple(torch.ops._C.top_k_sampling_from_probs(probs, uniform_samples, maybe_top_k_arr, top_k_val, deterministic))


def top_p_sampling_from_probs(probs, uniform_samples, maybe_top_p_arr, top_p_val, deterministic=True):
    return tuple(torch.ops._C.top_p_sampling_from_probs(probs, uniform_samples, maybe_top_p_arr, top_p_val, deterministic))


def min_p_sampling_from_probs(probs, uniform_samples, maybe_min_p_arr, min_p_val, deterministic=True):
    return tuple(torch.ops._C.min_p_sampling_from_probs(probs, uniform_samples, maybe_min_p_arr, min_p_val, deterministic))


def top_k_top_p_sampling_from_probs(probs, uniform_samples, maybe_top_k_arr, top_k_val, maybe_top_p_arr, top_p_val,
                                    deterministic=True):
    return tuple(torch.ops._C.top_k_top_p_sampling_from_probs(probs, uniform_samples, maybe_top_k_arr, float(top_k_val),
                                                              maybe_top_p_arr, top_p_val, deterministic))


def scaled_fp8_quant(input, scale=None, num_token_padding=None, scale_ub=None, use_per_token_if_dynamic=False):
    """Quantise [tokens, hidden] to float8_e4m3fn; returns (output, scale). Static when `scale` is given, else dynamic
    per tensor or per token — the reference's function of the same name (aphrodite/_custom_ops.py:632-685)."""
    assert input.ndim == 2
    shape = input.shape
    if num_token_padding:
        shape = (max(num_token_padding, input.shape[0]), shape[1])
    output = torch.empty(shape, device=input.device, dtype=torch.float8_e4m3fn)
    if scale is None:
        if use_per_token_if_dynamic:
            scale = torch.empty((shape[0], 1), device=input.device, dtype=torch.float32)
            torch.ops._C.dynamic_per_token_scaled_fp8_quant(output, input, scale, scale_ub)
        else:
            scale = torch.zeros(1, device=input.device, dtype=torch.float32)
            torch.ops._C.dynamic_scaled_fp8_quant(output, input, scale)
    else:
        assert scale.numel() == 1 or num_token_padding is None
        torch.ops._C.static_scaled_fp8_quant(output, input, scale)
    return output, scale


def cutlass_scaled_mm(a, b, scale_a, scale_b, out_dtype, bias=None):
    """out[M, N] = out_dtype(scale_a * (scale_b * (a @ b)) + bias); a [M, K] row-major, b [K, N] column-major, fp8-e4m3
    or int8 (reference aphrodite/_custom_ops.py:496-513)."""
    assert b.shape[0] % 16 == 0 and b.shape[1] % 16 == 0
    assert out_dtype is torch.bfloat16 or out_dtype is torch.float16
    assert bias is None or (bias.shape[0] == b.shape[1] and bias.dtype == out_dtype)
    out = torch.empty((a.shape[0], b.shape[1]), dtype=out_dtype, device=a.device)
    torch.ops._C.cutlass_scaled_mm(out, a, b, scale_a, scale_b, bias)
    return out


# wrappers written out by hand (they allocate / unpack, in the reference as well); everything else comes from _TABLE
COMPOSITE = ("top_k_sampling_from_probs", "top_p_sampling_from_probs", "min_p_sampling_from_probs",
             "top_k_top_p_sampling_from_probs", "scaled_fp8_quant", "cutlass_scaled_mm")
__all__ = [t[0] for t in _TABLE] + ["top_k_sampling_from_probs", "top_p_sampling_from_probs", "min_p_sampling_from_probs",
                                   "top_k_top_p_sampling_from_probs", "scaled_fp8_quant", "cutlass_scaled_mm"]
