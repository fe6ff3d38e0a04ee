"""TEST INFRASTRUCTURE — op table over the reference's OWN CUDA kernels (oracle/_ref/_ref_cuda_C.so, built by
oracle/build_ref_cuda.py from /root/reference/kernels where they lie), with the function names and argument order of
`aphrodite/_custom_ops.py`, so that bench.py's `ref_cuda` leg can run the identical decode-step call pattern
(aphrodite/modeling/models/llama.py:234-261) over the reference's kernels and time it beside this repo's. Never
imported by the product package; only tests/ and bench.py's baseline legs use it."""
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "_ref_cuda_C.so")
REF_AR_SO = os.path.join(_HERE, "_ref", "_ref_cuda_ar_C.so")     # needs libcuda.so.1: loadable on a GPU box only
_loaded = None
_loaded_ar = None


def available() -> bool:
    return os.path.exists(REF_SO)


def load():
    global _loaded
    if _loaded is None:
        if not available():
            raise FileNotFoundError(f"{REF_SO} not built (python oracle/build_ref_cuda.py where /root/reference exists)")
        torch.ops.load_library(REF_SO)
        _loaded = torch.ops._ref_cuda_C
    return _loaded


def load_ar():
    global _loaded_ar
    if _loaded_ar is None:
        if not os.path.exists(REF_AR_SO):
            raise FileNotFoundError(f"{REF_AR_SO} not built")
        torch.ops.load_library(REF_AR_SO)
        _loaded_ar = torch.ops._ref_cuda_ar_C
    return _loaded_ar


class RefCudaOps:
    """`ops.<name>(...)` with the signatures of aphrodite/_custom_ops.py, forwarding to torch.ops._ref_cuda_C."""

    def __init__(self):
        self.r = load()

    def rms_norm(self, out, input, weight, epsilon):
        self.r.rms_norm(out, input, weight, epsilon)

    def fused_add_rms_norm(self, input, residual, weight, epsilon):
        self.r.fused_add_rms_norm(input, residual, weight, epsilon)

    def rotary_embedding(self, positions, query, key, head_size, cos_sin_cache, is_neox):
        self.r.rotary_embedding(positions, query, key, head_size, cos_sin_cache, is_neox)

    def silu_and_mul(self, out, x):
        self.r.silu_and_mul(out, x)

    def reshape_and_cache(self, key, value, key_cache, value_cache, slot_mapping, kv_cache_dtype, k_scale, v_scale):
        self.r.reshape_and_cache(key, value, key_cache, value_cache, slot_mapping, kv_cache_dtype, k_scale, v_scale)

    def paged_attention_v1(self, out, query, key_cache, value_cache, num_kv_heads, scale, block_tables, seq_lens,
                           block_size, max_seq_len, alibi_slopes, kv_cache_dtype, k_scale, v_scale, tp_rank=0,
                           blocksparse_local_blocks=0, blocksparse_vert_stride=0, blocksparse_block_size=64,
                           blocksparse_head_sliding_step=0):
        self.r.paged_attention_v1(out, query, key_cache, value_cache, num_kv_heads, scale, block_tables, seq_lens,
                                  block_size, max_seq_len, alibi_slopes, kv_cache_dtype, k_scale, v_scale, tp_rank,
                                  blocksparse_local_blocks, blocksparse_vert_stride, blocksparse_block_size,
                                  blocksparse_head_sliding_step)

    def paged_attention_v2(self, out, exp_sum, max_logits, tmp_out, query, key_cache, value_cache, num_kv_heads, scale,
                           block_tables, seq_lens, block_size, max_seq_len, alibi_slopes, kv_cache_dtype, k_scale,
                           v_scale, tp_rank=0, blocksparse_local_blocks=0, blocksparse_vert_stride=0,
                           blocksparse_block_size=64, blocksparse_head_sliding_step=0):
        self.r.paged_attention_v2(out, exp_sum, max_logits, tmp_out, query, key_cache, value_cache, num_kv_heads,
                                  scale, block_tables, seq_lens, block_size, max_seq_len, alibi_slopes, kv_cache_dtype,
                                  k_scale, v_scale, tp_rank, blocksparse_local_blocks, blocksparse_vert_stride,
                                  blocksparse_block_size, blocksparse_head_sliding_step)

    def gptq_marlin_gemm(self, a, b_q_weight, b_scales, b_zeros, g_idx, perm, workspace, b_q_type, size_m, size_n,
                         size_k, is_k_full, has_zp=False, use_fp32_reduce=False, is_zp_float=False):
        # the binding takes the weight type as (exponent, mantissa, bias, signed): oracle/ref_cuda_bindings.cpp
        return self.r.gptq_marlin_gemm(a, b_q_weight, b_scales, b_zeros, g_idx, perm, workspace, b_q_type.exponent,
                                       b_q_type.mantissa, b_q_type.bias, b_q_type.signed, size_m, size_n, size_k,
                                       is_k_full, has_zp, use_fp32_reduce, is_zp_float)

    @property
    def ar(self):
        return load_ar()

    # the reference's custom all-reduce entry points (kernels/all_reduce/custom_all_reduce.cu)
    def init_custom_ar(self, meta, rank_data, handles, offsets, rank, full_nvlink):
        return self.ar.init_custom_ar(meta, rank_data, [_raw64(h) for h in handles], offsets, rank, full_nvlink)

    def register_buffer(self, fa, t, handles, offsets):
        self.ar.register_buffer(fa, t, [_raw64(h) for h in handles], offsets)

    def all_reduce_reg(self, fa, inp, out):
        self.ar.all_reduce_reg(fa, inp, out)

    def all_reduce_unreg(self, fa, inp, reg_buffer, out):
        self.ar.all_reduce_unreg(fa, inp, reg_buffer, out)

    def dispose(self, fa):
        self.ar.dispose(fa)

    def meta_size(self):
        return self.ar.meta_size()

    def get_graph_buffer_ipc_meta(self, fa):
        return self.ar.get_graph_buffer_ipc_meta(fa)

    def register_graph_buffers(self, fa, handles, offsets):
        self.ar.register_graph_buffers(fa, handles, offsets)


def _raw64(h):
    """torch >= 2.5 prefixes the 64-byte cudaIpcMemHandle_t with {version, type}; the reference (torch 2.4) memcpy's
    the first 64 bytes of whatever it is given (custom_all_reduce.cu:27-30)."""
    h = bytes(h) if not isinstance(h, (bytes, str)) else h
    return h[2:] if len(h) == 66 else h


def make_attention_cls(op_table):
    """The package's PagedAttention glue (V1/V2 rule, cache views — the reference's own logic) over `op_table`."""
    from aphrodite_engine_b200.attention.paged_attn import PagedAttention

    class RefPagedAttention(PagedAttention):
        _ops = op_table
    return RefPagedAttention


def make_custom_allreduce(cpu_group, device, op_table):
    """The package's CustomAllreduce host protocol (the reference's, custom_all_reduce.py:40-296) driving the
    REFERENCE's kernels."""
    from aphrodite_engine_b200.distributed.custom_all_reduce import CustomAllreduce

    class RefCustomAllreduce(CustomAllreduce):
        _ops = op_table
    return RefCustomAllreduce(cpu_group, device)
