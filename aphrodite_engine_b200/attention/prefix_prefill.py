"""Prefill attention over the paged KV cache under the reference's Python interface.

The reference implements this op in Triton (`context_attention_fwd`, aphrodite/attention/ops/prefix_prefill.py:696-858) and
calls it from `PagedAttention.forward_prefix` (attention/ops/paged_attn.py:192-228). There is no torch op to replace, so
the drop-in seam is this function: same name, same positional parameters, same defaults; it forwards to this package's
sm_100a kernel (csrc/prefill_attention.cu) through the `_C_b200::context_attention_fwd` op. INTEGRATION.md shows the
one-line rebinding a maintainer adds (`aphrodite.attention.ops.prefix_prefill.context_attention_fwd = <this>`)."""
from typing import Optional

import torch

from .. import _native

_native.load_torch_ops()


def context_attention_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor, kv_cache_dtype: str,
                          k_cache: torch.Tensor, v_cache: torch.Tensor, b_loc: torch.Tensor, b_start_loc: torch.Tensor,
                          b_seq_len: torch.Tensor, b_ctx_len: torch.Tensor, max_input_len: int, k_scale: float = 1.0,
                          v_scale: float = 1.0, alibi_slopes: Optional[torch.Tensor] = None,
                          sliding_window: Optional[int] = None) -> None:
    """q / o [tokens, heads, D], k / v [tokens, kv_heads, D]; k_cache [NB, kv_heads, D/x, BS, x], v_cache
    [NB, kv_heads, D, BS] (uint8 storage when kv_cache_dtype is an fp8 flavour, like the reference :721-734);
    b_loc = block table, b_start_loc = first query token of each sequence, b_seq_len = context + query length,
    b_ctx_len = cached context length. Writes o in place."""
    if "fp8" in kv_cache_dtype:
        assert k_cache.dtype == torch.uint8 and v_cache.dtype == torch.uint8
    elif k_cache.dtype == torch.uint8 or v_cache.dtype == torch.uint8:
        raise ValueError("kv_cache_dtype='auto' unsupported for FP8 KV Cache prefill kernel")
    assert q.shape[-1] == k.shape[-1] == v.shape[-1]
    window = 0 if sliding_window is None or sliding_window <= 0 else int(sliding_window)
    torch.ops._C_b200.context_attention_fwd(q, k, v, o, kv_cache_dtype, k_cache, v_cache, b_loc, b_start_loc, b_seq_len,
                                            b_ctx_len, int(max_input_len), float(k_scale), float(v_scale), alibi_slopes,
                                            window)
