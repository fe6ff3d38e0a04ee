"""Host-side mirror of the reference's Marlin mixture-of-experts layer glue
(aphrodite/modeling/layers/fused_moe/fused_moe.py:174-229 `moe_align_block_size`, :369-402 `fused_topk`,
:438-542 `fused_marlin_moe`): same names, arguments and tensor contracts; every device step is one of this
repo's ops (`topk_softmax`, `moe_align_block_size`, `marlin_gemm_moe`, `silu_and_mul`). Nothing is computed on
the CPU and there is no fallback: the ops raise if the CUDA extension is missing.
"""
from typing import Optional, Tuple

import torch

from aphrodite_engine_b200 import _custom_ops as ops
from aphrodite_engine_b200 import _native


def marlin_moe_block_size(num_tokens: int, num_experts: int) -> int:
    """BLOCK_SIZE_M of the reference's default config for the Marlin path (fused_moe.py:324-340)."""
    return 16 if (num_tokens <= num_experts or num_tokens <= 32) else 64


def moe_align_block_size(topk_ids: torch.Tensor, block_size: int,
                         num_experts: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Sorted (by expert) flat token-slot ids, padded per expert to `block_size`; padding = topk_ids.numel()."""
    max_num_tokens_padded = topk_ids.numel() + num_experts * (block_size - 1)
    sorted_ids = torch.full((max_num_tokens_padded, ), topk_ids.numel(), dtype=torch.int32, device=topk_ids.device)
    expert_ids = torch.empty((-(-max_num_tokens_padded // block_size), ), dtype=torch.int32, device=topk_ids.device)
    num_tokens_post_pad = torch.empty((1, ), dtype=torch.int32, device=topk_ids.device)
    ops.moe_align_block_size(topk_ids, num_experts, block_size, sorted_ids, expert_ids, num_tokens_post_pad)
    return sorted_ids, expert_ids, num_tokens_post_pad


def fused_topk(hidden_states: torch.Tensor, gating_output: torch.Tensor, topk: int,
               renormalize: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    assert hidden_states.shape[0] == gating_output.shape[0], "Number of tokens mismatch"
    M = hidden_states.shape[0]
    dev = hidden_states.device
    topk_weights = torch.empty(M, topk, dtype=torch.float32, device=dev)
    topk_ids = torch.empty(M, topk, dtype=torch.int32, device=dev)
    token_expert_indices = torch.empty(M, topk, dtype=torch.int32, device=dev)
    ops.topk_softmax(topk_weights, topk_ids, token_expert_indices, gating_output.float())
    if renormalize:
        topk_weights = topk_weights / topk_weights.sum(dim=-1, keepdim=True)
    return topk_weights, topk_ids


def fused_marlin_moe(hidden_states: torch.Tensor,
                     w1: torch.Tensor,
                     w2: torch.Tensor,
                     gating_output: torch.Tensor,
                     g_idx1: torch.Tensor,
                     g_idx2: torch.Tensor,
                     rand_perm1: torch.Tensor,
                     rand_perm2: torch.Tensor,
                     topk: int,
                     renormalize: bool = True,
                     w1_scale: Optional[torch.Tensor] = None,
                     w2_scale: Optional[torch.Tensor] = None,
                     block_size_m: Optional[int] = None) -> torch.Tensor:
    """out[t] = sum_k w[t,k] * down_{e(t,k)}( silu(gate) * up ) with 4-bit Marlin experts.

    w1: int32 [E, K/16, 2N*2] (gate|up), w2: int32 [E, N/16, K*2], w*_scale: [E, groups, 2N] / [E, groups, K].
    """
    assert hidden_states.shape[0] == gating_output.shape[0], "Number of tokens mismatch"
    assert hidden_states.shape[1] == w1.shape[1] * 16, "Hidden size mismatch w1"
    assert hidden_states.shape[1] == w2.shape[2] // 2, "Hidden size mismatch w2"
    assert gating_output.shape[1] == w1.shape[0], "Number of experts mismatch"
    assert hidden_states.is_contiguous() and w1.is_contiguous() and w2.is_contiguous()
    _native.load_torch_ops()
    M, K = hidden_states.shape
    E = w1.shape[0]
    N = w2.shape[1] * 16
    topk_weights, topk_ids = fused_topk(hidden_states, gating_output, topk, renormalize)
    if block_size_m is None:
        block_size_m = marlin_moe_block_size(M, E)
    sorted_token_ids, _, _ = moe_align_block_size(topk_ids, block_size_m, E)
    workspace = torch.zeros(((M + 255) // 256) * (max(2 * N, K) // 64) * 16, dtype=torch.int32,
                            device=hidden_states.device)
    gate_up = torch.ops._moe_C.marlin_gemm_moe(hidden_states, w1, sorted_token_ids, topk_weights, topk_ids, w1_scale,
                                               g_idx1, rand_perm1, workspace, M, 2 * N, K, True, E, topk,
                                               block_size_m, True, False)
    act = torch.empty((M * topk, N), device=hidden_states.device, dtype=hidden_states.dtype)
    ops.silu_and_mul(act, gate_up.view(-1, 2 * N))
    down = torch.ops._moe_C.marlin_gemm_moe(act, w2, sorted_token_ids, topk_weights, topk_ids, w2_scale, g_idx2,
                                            rand_perm2, workspace, M, K, N, True, E, topk, block_size_m, False, True)
    return torch.sum(down, dim=1)
