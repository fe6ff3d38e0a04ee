"""Mixtral's quantised mixture-of-experts block with the reference's semantics for `-q awq / gptq` checkpoints.

The reference routes quantised Mixtral to `aphrodite/modeling/models/mixtral_quant.py` (model_loader/utils.py:23-30):
experts are split ACROSS tensor-parallel ranks (`np.array_split(range(E), tp)[rank]`, :109-110), every local expert
runs DENSELY on all tokens through ordinary quantised linears (w1, w3 -> SiLU(w1 x) * (w3 x) -> w2, :82-88), its
output is scaled by the token's routing weight for that expert (zero when the expert was not selected, :143-148),
the local experts are summed and one all-reduce follows (:154). BASELINE configs[4] (Mixtral-8x7B AWQ, TP=4,
bs=128) is this block 32 times.

Here: the same arithmetic on this package's kernels — the AWQ-Marlin W4A16 GEMM (`gptq_marlin_gemm` with integer zero
points; w1 and w3 fused into one [2I] GEMM, which changes no output value), `silu_and_mul`, the router's
`topk_softmax`, and `moe_expert_scale_add` for the mask / scale / accumulate of the expert loop. Weights are
random-init in the Marlin layout (what `awq_marlin_repack` + `marlin_permute_scales` + `awq_to_marlin_zero_points`
produce at load time, aphrodite/quantization/awq_marlin.py); no checkpoint I/O.
"""
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn.functional as F

from . import _custom_ops as ops
from . import _native


@dataclass
class MixtralShape:
    name: str = "Mixtral-8x7B"
    hidden: int = 4096
    intermediate: int = 14336
    num_experts: int = 8
    topk: int = 2
    group_size: int = 128


def local_experts(num_experts: int, tp_size: int, tp_rank: int) -> List[int]:
    """np.array_split(range(E), tp)[rank] of the reference (mixtral_quant.py:109-110), without numpy."""
    base, extra = divmod(num_experts, tp_size)
    start = tp_rank * base + min(tp_rank, extra)
    return list(range(start, start + base + (1 if tp_rank < extra else 0)))


class MixtralQuantMoE:
    def __init__(self, shape: MixtralShape, device, dtype=torch.bfloat16, tp_rank: int = 0, tp_size: int = 1,
                 seed: int = 4321, op_table=None, fused_scale_add: Optional[bool] = None, share_from=None):
        if tp_size > shape.num_experts:
            raise ValueError(f"Tensor parallel size {tp_size} is greater than the number of experts {shape.num_experts}.")
        self.s, self.device, self.dtype = shape, device, dtype
        self.tp_rank, self.tp_size = tp_rank, tp_size
        self.experts = local_experts(shape.num_experts, tp_size, tp_rank)
        self.ops = ops if op_table is None else op_table
        # the expert loop's mask / scale / accumulate: one kernel here, torch element-wise ops in the reference
        self.fused_scale_add = (op_table is None) if fused_scale_add is None else fused_scale_add
        from .scalar_type import scalar_types
        self._qtype = scalar_types.uint4
        self._empty = torch.empty(0, dtype=torch.int32, device=device)
        if share_from is not None:
            self.gate, self.w13, self.w2 = share_from.gate, share_from.w13, share_from.w2
            return
        H, I, G = shape.hidden, shape.intermediate, shape.group_size
        g = torch.Generator(device=device).manual_seed(seed)
        self.gate = (torch.randn(shape.num_experts, H, generator=g, device=device, dtype=torch.float32) * 0.02).to(dtype)

        def awq_marlin(n_out, k_in, gen):
            return dict(q=torch.randint(-2**31, 2**31 - 1, (k_in // 16, n_out * 2), generator=gen, device=device,
                                        dtype=torch.int32),
                        s=(torch.rand(k_in // G, n_out, generator=gen, device=device) * 0.004 + 0.001).to(dtype),
                        z=torch.randint(-2**31, 2**31 - 1, (k_in // G, n_out // 8), generator=gen, device=device,
                                        dtype=torch.int32),
                        ws=torch.zeros((n_out // 64) * 16, dtype=torch.int32, device=device), n=n_out, k=k_in)
        self.w13, self.w2 = {}, {}
        for e in range(shape.num_experts):        # every rank draws all experts' seeds, keeps its own: TP=N == TP=1
            ge = torch.Generator(device=device).manual_seed(seed + 1 + e)
            if e in self.experts:
                self.w13[e] = awq_marlin(2 * I, H, ge)
                self.w2[e] = awq_marlin(H, I, ge)

    def _linear(self, x, wt):
        """apply_awq_marlin_linear (aphrodite/quantization/utils/marlin_utils.py:278-315) with integer zero points."""
        return self.ops.gptq_marlin_gemm(x, wt["q"], wt["s"], wt["z"], self._empty, self._empty, wt["ws"], self._qtype,
                                         x.shape[0], wt["n"], wt["k"], True, True, True, False)

    def route(self, hidden_states: torch.Tensor):
        """softmax(fp32) -> top-k -> renormalise (mixtral_quant.py:133-139). Returns (weights f32, ids int32) [T, k]."""
        logits = F.linear(hidden_states, self.gate)
        T, k = hidden_states.shape[0], self.s.topk
        w = torch.empty(T, k, dtype=torch.float32, device=self.device)
        ids = torch.empty(T, k, dtype=torch.int32, device=self.device)
        src = torch.empty(T, k, dtype=torch.int32, device=self.device)
        ops.topk_softmax(w, ids, src, logits.float())
        return w / w.sum(dim=-1, keepdim=True), ids

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        """This rank's partial sum over its experts, BEFORE the tensor-parallel all-reduce."""
        T = hidden_states.shape[0]
        weights, ids = self.route(hidden_states)
        final = None
        for e in self.experts:
            gate_up = self._linear(hidden_states, self.w13[e])
            act = torch.empty(T, self.s.intermediate, dtype=self.dtype, device=self.device)
            self.ops.silu_and_mul(act, gate_up)
            cur = self._linear(act, self.w2[e])
            if self.fused_scale_add:
                if final is None:
                    final = torch.empty_like(cur)
                    _native.load_torch_ops()
                    torch.ops._C_b200.moe_expert_scale_add(final, cur, weights, ids, e, True)
                else:
                    torch.ops._C_b200.moe_expert_scale_add(final, cur, weights, ids, e, False)
            else:                                  # the reference's element-wise sequence (mixtral_quant.py:143-152)
                expert_weights = (weights * (ids == e)).sum(dim=-1, keepdim=True)
                cur = cur.mul_(expert_weights)
                final = cur if final is None else final.add_(cur)
        return final
