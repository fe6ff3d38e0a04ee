"""Attention glue with the reference's interface (`PagedAttention` in aphrodite/attention/ops/paged_attn.py:
cache shape :49-56, views :58-72, writer :74-95, decode dispatch :97-190, prefix prefill :192-228), over this package's ops.

Kept from the reference because callers and checkpointed caches depend on it: the `[2, num_blocks, block*heads*dim]`
allocation viewed as K `[blocks, kv_heads, dim/x, block, x]` (x = 16 bytes of elements) and V `[blocks, kv_heads, dim,
block]`; the 512-token partition of the split-KV kernel; and the rule for choosing the single-pass kernel — at most
8192 cached tokens AND (one partition OR more than 512 (sequence, head) pairs to fill the GPU with) — otherwise the
partitioned kernel with its (exp_sum, max_logit, tmp_out) scratch."""
from typing import Optional, Tuple

import torch

from .. import _custom_ops as ops

_PARTITION_SIZE = 512          # tokens per split-KV partition; equals kPartitionSize in csrc/paged_attention.cu
_V1_MAX_CONTEXT = 8192
_V1_MIN_PAIRS = 512


def _wants_single_pass(max_seq_len: int, partitions: int, seq_head_pairs: int) -> bool:
    if max_seq_len > _V1_MAX_CONTEXT:
        return False
    return partitions == 1 or seq_head_pairs > _V1_MIN_PAIRS


class PagedAttention:
    # the op table the two kernel-calling methods use; a subclass may point it at another implementation of the same
    # op names (bench.py's reference-CUDA arm does, to time the reference's kernels under the identical call pattern)
    _ops = ops

    @staticmethod
    def get_kv_cache_shape(num_blocks: int, block_size: int, num_kv_heads: int, head_size: int) -> Tuple[int, ...]:
        per_block = block_size * num_kv_heads * head_size
        return (2, num_blocks, per_block)

    @staticmethod
    def split_kv_cache(kv_cache: torch.Tensor, num_kv_heads: int,
                       head_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
        lanes = 16 // kv_cache.element_size()            # elements per 16-byte run of the K layout
        blocks = kv_cache.shape[1]
        k_plane, v_plane = kv_cache[0], kv_cache[1]
        return (k_plane.view(blocks, num_kv_heads, head_size // lanes, -1, lanes),
                v_plane.view(blocks, num_kv_heads, head_size, -1))

    @classmethod
    def write_to_paged_cache(cls, key, value, key_cache, value_cache, slot_mapping, kv_cache_dtype: str, k_scale: float,
                             v_scale: float) -> None:
        slots = slot_mapping.flatten()
        cls._ops.reshape_and_cache(key, value, key_cache, value_cache, slots, kv_cache_dtype, k_scale, v_scale)

    @classmethod
    def forward_decode(cls, query: torch.Tensor, key_cache: torch.Tensor, value_cache: torch.Tensor,
                       block_tables: torch.Tensor, seq_lens: torch.Tensor, max_seq_len: int, kv_cache_dtype: str,
                       num_kv_heads: int, scale: float, alibi_slopes: Optional[torch.Tensor], k_scale: float,
                       v_scale: float, tp_rank: int = 0, blocksparse_local_blocks: int = 0,
                       blocksparse_vert_stride: int = 0, blocksparse_block_size: int = 64,
                       blocksparse_head_sliding_step: int = 0,
                       output: Optional[torch.Tensor] = None) -> torch.Tensor:
        block_size = value_cache.shape[3]
        sparse = blocksparse_vert_stride is not None and blocksparse_vert_stride > 1
        if sparse and not (blocksparse_block_size > 0 and blocksparse_block_size % block_size == 0):
            raise AssertionError(f"{blocksparse_block_size=} needs to be a multiple of {block_size=} used in block_tables.")
        out = torch.empty_like(query) if output is None else output
        num_seqs, num_heads, head_size = query.shape
        partitions = -(-max_seq_len // _PARTITION_SIZE)
        common = (num_kv_heads, scale, block_tables, seq_lens, block_size, max_seq_len, alibi_slopes, kv_cache_dtype,
                  k_scale, v_scale, tp_rank, blocksparse_local_blocks, blocksparse_vert_stride, blocksparse_block_size,
                  blocksparse_head_sliding_step)
        if _wants_single_pass(max_seq_len, partitions, num_seqs * num_heads):
            cls._ops.paged_attention_v1(out, query, key_cache, value_cache, *common)
            return out
        assert _PARTITION_SIZE % block_size == 0
        stats_shape = (num_seqs, num_heads, partitions)
        exp_sums = torch.empty(stats_shape, dtype=torch.float32, device=out.device)
        max_logits = torch.empty_like(exp_sums)
        tmp_out = torch.empty(stats_shape + (head_size, ), dtype=out.dtype, device=out.device)
        cls._ops.paged_attention_v2(out, exp_sums, max_logits, tmp_out, query, key_cache, value_cache, *common)
        return out

    @staticmethod
    def forward_prefix(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, kv_cache_dtype: str,
                       key_cache: torch.Tensor, value_cache: torch.Tensor, block_tables: torch.Tensor,
                       query_start_loc: torch.Tensor, seq_lens_tensor: torch.Tensor, context_lens: torch.Tensor,
                       max_query_len: int, alibi_slopes: Optional[torch.Tensor], sliding_window: Optional[int],
                       k_scale: float, v_scale: float) -> torch.Tensor:
        """Prefill with cached context: query tokens attend to the paged cache and, causally, to this step's keys.
        query_start_loc has batch + 1 entries (the reference drops the last one the same way, :216-217)."""
        from .prefix_prefill import context_attention_fwd
        output = torch.empty_like(query)
        context_attention_fwd(query, key, value, output, kv_cache_dtype, key_cache, value_cache, block_tables,
                              query_start_loc[:-1], seq_lens_tensor, context_lens, max_query_len, k_scale, v_scale,
                              alibi_slopes, sliding_window)
        return output
