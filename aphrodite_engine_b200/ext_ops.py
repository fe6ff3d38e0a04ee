"""Extension ops: launches this package adds BEYOND the reference's op set (torch namespace `_C_b200`, registered by
csrc/torch_shim.cpp). They fuse neighbouring ops of the reference's decode layer; each is defined as "the same outputs
as the reference ops it replaces, run back to back", so a caller may swap them in without changing results
(INTEGRATION.md, section "extensions").

    rotary_embedding_and_cache   = ops.rotary_embedding(positions, q, k, ...) ; ops.reshape_and_cache(k, v, ...)
                                   (aphrodite/modeling/layers/rotary_embedding.py:152-173 then
                                   aphrodite/attention/ops/paged_attn.py:74-95)
    tp_allreduce_rows            = tensor_model_parallel_all_reduce(x) ; ops.fused_add_rms_norm(x, residual, w, eps)
                                   (see distributed/nvls.py, which owns the symmetric memory it needs)
"""
import torch

from . import _native

_native.load_torch_ops()


def rotary_embedding_and_cache(positions: torch.Tensor, query: torch.Tensor, key: torch.Tensor, value: torch.Tensor,
                               head_size: int, cos_sin_cache: torch.Tensor, is_neox: bool, key_cache: torch.Tensor,
                               value_cache: torch.Tensor, slot_mapping: torch.Tensor, kv_cache_dtype: str,
                               k_scale: float, v_scale: float) -> None:
    """query [T, Hq*D], key / value [T, Hkv*D] (may be views of the fused qkv output); in place on query and key."""
    torch.ops._C_b200.rotary_embedding_and_cache(positions, query, key, value, head_size, cos_sin_cache, is_neox,
                                                 key_cache, value_cache, slot_mapping.flatten(), kv_cache_dtype,
                                                 k_scale, v_scale)
