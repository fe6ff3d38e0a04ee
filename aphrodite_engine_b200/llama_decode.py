"""Decode-step harness: the reference's Llama call pattern over this package's ops.

One decode step issues, per layer, exactly the op sequence of the reference's model code
(aphrodite/modeling/models/llama.py:169-181 attention, :88-92 MLP, :234-261 decoder layer) and of its
attention backend (aphrodite/attention/backends/xformers.py:550 reshape_and_cache, :635
PagedAttention.forward_decode):

    fused_add_rms_norm -> qkv GEMM -> rotary_embedding -> reshape_and_cache -> paged_attention
    -> o_proj GEMM [-> TP all-reduce] -> fused_add_rms_norm -> gate_up GEMM -> silu_and_mul
    -> down GEMM [-> TP all-reduce]            ... final rms_norm -> lm_head GEMM -> greedy token

Unquantised GEMMs are library calls (cuBLAS through F.linear) exactly as in the reference
(aphrodite/modeling/layers/linear.py:146); everything else is this package's hand-written kernels
reached through the reference-named wrappers in `_custom_ops`. Tensor parallelism is the reference's
head / column-row split (linear.py:258,991; two row-parallel all-reduces per layer, llama.py:72-92,
132-147) over torch.distributed NCCL. Weights are random-init (the reference's `dummy` load format,
modeling/model_loader/weight_utils.py:595-625); this module holds no checkpoint I/O.
"""
from dataclasses import dataclass
from typing import Callable, List, Optional

import torch
import torch.nn.functional as F

from . import _custom_ops as ops
from .attention.paged_attn import PagedAttention


@dataclass
class LlamaShape:
    name: str = "Llama-3-8B"
    hidden: int = 4096
    layers: int = 32
    heads: int = 32
    kv_heads: int = 8
    head_size: int = 128
    intermediate: int = 14336
    vocab: int = 128256
    rope_theta: float = 500000.0
    max_position: int = 8192
    rms_eps: float = 1e-5


class DecodeState:
    """Static device buffers of one decode batch (what the reference's model_runner prepares per step
    and keeps alive for CUDA-graph replay, aphrodite/worker/model_runner.py:1682+)."""

    def __init__(self, batch: int, max_blocks: int, device):
        self.input_ids = torch.zeros(batch, dtype=torch.long, device=device)
        self.positions = torch.zeros(batch, dtype=torch.long, device=device)
        self.slot_mapping = torch.zeros(batch, dtype=torch.long, device=device)
        self.seq_lens = torch.zeros(batch, dtype=torch.int32, device=device)
        self.block_tables = torch.zeros(batch, max_blocks, dtype=torch.int32, device=device)
        self.next_tokens = torch.zeros(batch, dtype=torch.long, device=device)
        self.max_seq_len = 0

    def tensors(self):
        return [self.input_ids, self.positions, self.slot_mapping, self.seq_lens, self.block_tables]


class LlamaDecoder:
    def __init__(self, shape: LlamaShape, batch: int, block_size: int, num_blocks: int,
                 device, dtype=torch.bfloat16, kv_cache_dtype: str = "auto", tp_rank: int = 0,
                 tp_size: int = 1, group=None, seed: int = 1234, layers: Optional[int] = None,
                 kv_fill: bool = True, quant: Optional[str] = None, group_size: int = 128,
                 custom_ar=None):
        assert shape.heads % tp_size == 0 and shape.kv_heads % tp_size == 0
        assert shape.intermediate % tp_size == 0 and shape.vocab % tp_size == 0
        self.s, self.batch, self.block_size = shape, batch, block_size
        self.device, self.dtype, self.kv_cache_dtype = device, dtype, kv_cache_dtype
        self.tp_rank, self.tp_size, self.group = tp_rank, tp_size, group
        self.custom_ar = custom_ar      # distributed.CustomAllreduce or None (-> NCCL)
        self.n_layers = shape.layers if layers is None else layers
        self.heads = shape.heads // tp_size
        self.kv_heads = shape.kv_heads // tp_size
        self.q_size = self.heads * shape.head_size
        self.kv_size = self.kv_heads * shape.head_size
        self.inter = shape.intermediate // tp_size
        self.vocab_local = shape.vocab // tp_size
        self.scale = shape.head_size ** -0.5
        g = torch.Generator(device=device).manual_seed(seed + tp_rank)

        def w(*sz, std=0.02):
            return (torch.randn(*sz, generator=g, device=device, dtype=torch.float32) * std).to(dtype)

        self.quant = quant
        if quant == "gptq":
            # GPTQ 4-bit (uint4b8, group 128) in Marlin layout — BASELINE configs[2]. Random packed words and
            # scales stand in for a checkpoint (the `dummy` load format); the layouts are what
            # gptq_marlin_repack / marlin_permute_scales would produce (process_weights_after_loading,
            # aphrodite/quantization/kernels/marlin.py:89-109).
            from .scalar_type import scalar_types
            self._qtype = scalar_types.uint4b8
            self._empty = torch.empty(0, dtype=torch.int32, device=device)

            def qw(n_out, k_in):
                return dict(q=torch.randint(-2**31, 2**31 - 1, (k_in // 16, n_out * 2), generator=g,
                                            device=device, dtype=torch.int32),
                            s=(torch.rand(k_in // group_size, n_out, generator=g, device=device) * 0.004 + 0.001).to(dtype),
                            ws=torch.zeros((n_out // 64) * 16, dtype=torch.int32, device=device), n=n_out, k=k_in)
            w_lin = qw
        else:
            def w_lin(n_out, k_in):
                return w(n_out, k_in)

        H = shape.hidden
        self.embed = w(shape.vocab, H)
        self.layers = []
        for _ in range(self.n_layers):
            self.layers.append(dict(
                ln1=torch.ones(H, dtype=dtype, device=device),
                ln2=torch.ones(H, dtype=dtype, device=device),
                qkv=w_lin(self.q_size + 2 * self.kv_size, H),
                o=w_lin(H, self.q_size),
                gate_up=w_lin(2 * self.inter, H),
                down=w_lin(H, self.inter),
            ))
        self.norm = torch.ones(H, dtype=dtype, device=device)
        self.lm_head = w(self.vocab_local, H)
        # rotary cache [max_pos, rot_dim] = cat(cos, sin) (modeling/layers/rotary_embedding.py:105-120)
        inv = 1.0 / (shape.rope_theta ** (torch.arange(0, shape.head_size, 2, dtype=torch.float32) /
                                          shape.head_size))
        fr = torch.einsum("i,j->ij", torch.arange(shape.max_position, dtype=torch.float32), inv)
        self.cos_sin = torch.cat((fr.cos(), fr.sin()), dim=-1).to(dtype).to(device)
        # KV cache: one [2, num_blocks, block_size*kv_heads*head_size] tensor per layer
        # (aphrodite/worker/cache_engine.py:65-84)
        store = dtype if kv_cache_dtype == "auto" else torch.uint8
        self.kv_caches = []
        for _ in range(self.n_layers):
            kv = torch.empty(PagedAttention.get_kv_cache_shape(num_blocks, block_size, self.kv_heads,
                                                              shape.head_size),
                             dtype=store, device=device)
            if kv_fill:
                if store == torch.uint8:
                    kv.random_(0, 120, generator=g)
                else:
                    kv.uniform_(-self.scale, self.scale, generator=g)
            self.kv_caches.append(kv)
        self.kv_views = [PagedAttention.split_kv_cache(kv, self.kv_heads, shape.head_size)
                         for kv in self.kv_caches]
        self.attn_hook: Optional[Callable] = None   # bench instrumentation around the attention op
        self.my_kernel_launches_per_step = (10 if quant == "gptq" else 6) * self.n_layers + 1

    def _linear(self, x, wt):
        """F.linear for bf16 weights (cuBLAS, as the reference's UnquantizedLinearMethod) or the Marlin-format
        W4A16 GEMM (apply_gptq_marlin_linear, aphrodite/quantization/utils/marlin_utils.py:241-275)."""
        if isinstance(wt, dict):
            return ops.gptq_marlin_gemm(x, wt["q"], wt["s"], self._empty, self._empty, self._empty, wt["ws"],
                                        self._qtype, x.shape[0], wt["n"], wt["k"], True, False, True, False)
        return F.linear(x, wt)

    def _all_reduce(self, x):
        """Row-parallel reduction (GroupCoordinator._all_reduce, aphrodite/distributed/parallel_state.py:353-379):
        the NVLink peer-memory kernel when it applies (out of place), else NCCL in place."""
        if self.tp_size > 1:
            if self.custom_ar is not None:
                out = self.custom_ar.custom_all_reduce(x)
                if out is not None:
                    return out
            torch.distributed.all_reduce(x, group=self.group)
        return x

    def forward(self, st: DecodeState) -> torch.Tensor:
        s = self.s
        hidden = F.embedding(st.input_ids, self.embed)
        residual = None
        for li, L in enumerate(self.layers):
            if residual is None:
                residual = hidden
                normed = torch.empty_like(hidden)
                ops.rms_norm(normed, hidden, L["ln1"], s.rms_eps)
                hidden = normed
            else:
                ops.fused_add_rms_norm(hidden, residual, L["ln1"], s.rms_eps)
            qkv = self._linear(hidden, L["qkv"])
            q, k, v = qkv.split([self.q_size, self.kv_size, self.kv_size], dim=-1)
            ops.rotary_embedding(st.positions, q, k, s.head_size, self.cos_sin, True)
            kc, vc = self.kv_views[li]
            PagedAttention.write_to_paged_cache(k.view(-1, self.kv_heads, s.head_size),
                                                v.view(-1, self.kv_heads, s.head_size), kc, vc,
                                                st.slot_mapping, self.kv_cache_dtype, 1.0, 1.0)
            qv = q.view(-1, self.heads, s.head_size)
            attn_out = torch.empty(self.batch, self.heads, s.head_size, dtype=self.dtype,
                                   device=self.device)
            if self.attn_hook is not None:
                self.attn_hook(li, True)
            PagedAttention.forward_decode(qv, kc, vc, st.block_tables, st.seq_lens, st.max_seq_len,
                                          self.kv_cache_dtype, self.kv_heads, self.scale, None, 1.0,
                                          1.0, output=attn_out)
            if self.attn_hook is not None:
                self.attn_hook(li, False)
            hidden = self._all_reduce(self._linear(attn_out.view(self.batch, -1), L["o"]))
            ops.fused_add_rms_norm(hidden, residual, L["ln2"], s.rms_eps)
            gate_up = self._linear(hidden, L["gate_up"])
            act = torch.empty(self.batch, self.inter, dtype=self.dtype, device=self.device)
            ops.silu_and_mul(act, gate_up)
            hidden = self._all_reduce(self._linear(act, L["down"]))
        ops.fused_add_rms_norm(hidden, residual, self.norm, s.rms_eps)
        logits = F.linear(hidden, self.lm_head)
        if self.tp_size == 1:
            torch.argmax(logits, dim=-1, out=st.next_tokens)
        else:
            # greedy over vocab shards: local (max, argmax) then a tiny all-gather
            mx, idx = logits.float().max(dim=-1)
            idx = idx + self.tp_rank * self.vocab_local
            pair = torch.stack((mx, idx.float()), dim=-1).contiguous()
            gathered = [torch.empty_like(pair) for _ in range(self.tp_size)]
            torch.distributed.all_gather(gathered, pair, group=self.group)
            allp = torch.stack(gathered, dim=0)                # [tp, B, 2]
            best = allp[..., 0].argmax(dim=0)                  # [B]
            st.next_tokens.copy_(allp[best, torch.arange(self.batch, device=self.device), 1].long())
        return st.next_tokens


def make_synthetic_batch(batch: int, ctx: int, block_size: int, vocab_hi: int = 10000, seed: int = 0):
    """Host-side (pinned) inputs of one decode step at a fixed context length: every sequence owns
    distinct blocks (block_tables = randperm), the new token sits at position ctx-1.
    Token ids in [0, 10000) as tests/benchmarks/engine/latency.py:60-62 of the reference."""
    g = torch.Generator().manual_seed(seed)
    nb_per = (ctx + block_size - 1) // block_size
    num_blocks = batch * nb_per
    bt = torch.randperm(num_blocks, generator=g).view(batch, nb_per).to(torch.int32)
    pos = torch.full((batch,), ctx - 1, dtype=torch.long)
    slot = bt[:, (ctx - 1) // block_size].long() * block_size + (ctx - 1) % block_size
    host = dict(
        input_ids=torch.randint(0, vocab_hi, (batch,), generator=g, dtype=torch.long),
        positions=pos,
        slot_mapping=slot,
        seq_lens=torch.full((batch,), ctx, dtype=torch.int32),
        block_tables=bt,
    )
    return {k: v.pin_memory() if torch.cuda.is_available() else v for k, v in host.items()}, num_blocks


def upload(st: DecodeState, host: dict, non_blocking: bool = True) -> int:
    """Host -> device copy of one step's inputs into the static buffers; returns bytes copied."""
    n = 0
    for name in ("input_ids", "positions", "slot_mapping", "seq_lens", "block_tables"):
        dst = getattr(st, name)
        dst.copy_(host[name], non_blocking=non_blocking)
        n += dst.numel() * dst.element_size()
    st.max_seq_len = int(host["seq_lens"].max())
    return n
