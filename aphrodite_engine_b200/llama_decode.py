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
132-147). The exchange after a row-parallel GEMM has three implementations, chosen by the caller:

    tp_mode "nccl"    torch.distributed all_reduce, then fused_add_rms_norm           (the reference's fallback path)
    tp_mode "p2p"     the IPC peer-memory all-reduce kernel (distributed.CustomAllreduce), then fused_add_rms_norm
                      (the reference's custom all-reduce path, parallel_state.py:353-379)
    tp_mode "nvls"    ONE kernel: NVSwitch multicast reduce + residual add + RMSNorm + multicast store
                      (distributed.NvlsTensorParallel / csrc/tp_fused.cu); the residual stream is token-sharded

Weights are random-init (the reference's `dummy` load format, modeling/model_loader/weight_utils.py:595-625); every
rank draws the FULL tensors from one rank-independent generator and keeps its shard, so a TP=N model is the same model
as TP=1 (replicated parameters are identical, sharded ones are slices). This module holds no checkpoint I/O.
"""
from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.nn.functional as F

from . import _custom_ops as ops
from . import ext_ops
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
                 custom_ar=None, nvls=None, op_table=None, attention_cls=None, share_from: "LlamaDecoder" = None,
                 fuse_rope_cache: Optional[bool] = None, share_kv_from: "LlamaDecoder" = None,
                 share_weights_from: "LlamaDecoder" = None):
        assert shape.heads % tp_size == 0 and shape.kv_heads % tp_size == 0
        assert shape.intermediate % tp_size == 0 and shape.vocab % tp_size == 0
        self.s, self.batch, self.block_size = shape, batch, block_size
        self.device, self.dtype, self.kv_cache_dtype = device, dtype, kv_cache_dtype
        self.tp_rank, self.tp_size, self.group = tp_rank, tp_size, group
        self.custom_ar = custom_ar      # distributed.CustomAllreduce or None
        self.nvls = nvls                # distributed.NvlsTensorParallel or None
        self.tp_mode = "none" if tp_size == 1 else ("nvls" if nvls is not None else
                                                   ("p2p" if custom_ar is not None else "nccl"))
        # the op table (this package's kernels unless a checker substitutes the reference's) and its attention glue
        self.ops = ops if op_table is None else op_table
        self.attn = PagedAttention if attention_cls is None else attention_cls
        # rotary + cache write as one launch (ext_ops; same outputs) unless a substituted op table is being timed
        self.fuse_rope_cache = (op_table is None) if fuse_rope_cache is None else fuse_rope_cache
        self.n_layers = shape.layers if layers is None else layers
        self.heads = shape.heads // tp_size
        self.kv_heads = shape.kv_heads // tp_size
        self.q_size = self.heads * shape.head_size
        self.kv_size = self.kv_heads * shape.head_size
        self.inter = shape.intermediate // tp_size
        self.vocab_local = shape.vocab // tp_size
        self.scale = shape.head_size ** -0.5
        self.quant = quant
        self.attn_hook: Optional[Callable] = None   # bench instrumentation around the attention op
        self.last_hidden: Optional[torch.Tensor] = None   # final normed hidden state of the last forward()
        if quant == "gptq":
            from .scalar_type import scalar_types
            self._qtype = scalar_types.uint4b8
            self._empty = torch.empty(0, dtype=torch.int32, device=device)
        if share_from is not None:      # a second decoder over the SAME weights and KV cache (another op table / TP mode)
            o = share_from
            assert (o.tp_size, o.tp_rank, o.quant, o.n_layers) == (tp_size, tp_rank, quant, self.n_layers)
            self.embed, self.layers, self.norm, self.lm_head = o.embed, o.layers, o.norm, o.lm_head
            self.cos_sin, self.kv_caches, self.kv_views = o.cos_sin, o.kv_caches, o.kv_views
        else:
            # share_weights_from: same parameters over another KV cache (other batch / context / cache dtype);
            # share_kv_from: same cache, embedding, head and norms, other linear weights (e.g. the quantised variant)
            if share_weights_from is not None:
                o = share_weights_from
                assert (o.tp_size, o.tp_rank, o.quant, o.n_layers) == (tp_size, tp_rank, quant, self.n_layers)
                self.embed, self.layers, self.norm, self.lm_head, self.cos_sin = o.embed, o.layers, o.norm, o.lm_head, o.cos_sin
            else:
                self._init_weights(seed, group_size, reuse=share_kv_from)
            if share_kv_from is not None:
                o = share_kv_from
                assert (o.tp_size, o.tp_rank, o.n_layers, o.kv_cache_dtype) == (tp_size, tp_rank, self.n_layers, kv_cache_dtype)
                self.kv_caches, self.kv_views = o.kv_caches, o.kv_views
            else:
                self._init_kv_cache(num_blocks, kv_fill, seed)
        # norm x2 (or 2 fused exchanges), rope, cache write (one launch when fused), attention, act (+ 4 W4A16 GEMMs)
        per_layer = 6 - (1 if self.fuse_rope_cache else 0) + (4 if quant == "gptq" else 0)
        if self.tp_mode == "p2p":
            per_layer += 2
        self.my_kernel_launches_per_step = per_layer * self.n_layers + 1

    # ------------------------------------------------------------------------------------------------------------
    def _init_weights(self, seed: int, group_size: int, reuse: "LlamaDecoder" = None):
        s, dev, dtype, r, n = self.s, self.device, self.dtype, self.tp_rank, self.tp_size
        g = torch.Generator(device=dev).manual_seed(seed)
        H = s.hidden

        def full(n_out, k_in, std=0.02):
            return (torch.randn(n_out, k_in, generator=g, device=dev, dtype=torch.float32) * std).to(dtype)

        def rows(t, parts):
            """Column-parallel shard: `t` is a concatenation of `parts` row blocks; keep this rank's slice of each."""
            if n == 1:
                return t
            out, at = [], 0
            for p in parts:
                step = p // n
                out.append(t[at + r * step: at + (r + 1) * step])
                at += p
            return torch.cat(out, dim=0).contiguous()

        def cols(t):
            step = t.shape[1] // n
            return t[:, r * step:(r + 1) * step].contiguous() if n > 1 else t

        if self.quant == "gptq":
            # GPTQ 4-bit (uint4b8, group 128) in Marlin layout — BASELINE configs[2]. Random packed words and scales
            # stand in for a checkpoint (the `dummy` load format); the layouts are what gptq_marlin_repack /
            # marlin_permute_scales produce (process_weights_after_loading, aphrodite/quantization/kernels/marlin.py:
            # 89-109). Marlin tiles are 16 k x 64 n, so an n-shard is a column range and a k-shard a row range.
            def qfull(n_out, k_in):
                q = torch.randint(-2**31, 2**31 - 1, (k_in // 16, n_out * 2), generator=g, device=dev, dtype=torch.int32)
                sc = (torch.rand(k_in // group_size, n_out, generator=g, device=dev) * 0.004 + 0.001).to(dtype)
                return q, sc

            def pack(q, sc):
                n_out, k_in = sc.shape[1], q.shape[0] * 16
                return dict(q=q.contiguous(), s=sc.contiguous(), n=n_out, k=k_in,
                            ws=torch.zeros((n_out // 64) * 16, dtype=torch.int32, device=dev))

            def col_parallel(n_out, k_in, parts):
                q, sc = qfull(n_out, k_in)
                qs, ss, at = [], [], 0
                for p in parts:
                    step = p // n
                    lo, hi = at + r * step, at + (r + 1) * step
                    qs.append(q[:, lo * 2:hi * 2]); ss.append(sc[:, lo:hi])
                    at += p
                return pack(torch.cat(qs, dim=1), torch.cat(ss, dim=1))

            def row_parallel(n_out, k_in):
                q, sc = qfull(n_out, k_in)
                step = k_in // n
                return pack(q[r * step // 16:(r + 1) * step // 16], sc[r * step // group_size:(r + 1) * step // group_size])
        else:
            def col_parallel(n_out, k_in, parts):
                return rows(full(n_out, k_in), parts)

            def row_parallel(n_out, k_in):
                return cols(full(n_out, k_in))

        self.embed = full(s.vocab, H) if reuse is None else reuse.embed
        QS, KS, I = s.heads * s.head_size, s.kv_heads * s.head_size, s.intermediate
        self.layers = []
        for _ in range(self.n_layers):
            self.layers.append(dict(
                ln1=torch.ones(H, dtype=dtype, device=dev),
                ln2=torch.ones(H, dtype=dtype, device=dev),
                qkv=col_parallel(QS + 2 * KS, H, (QS, KS, KS)),
                o=row_parallel(H, QS),
                gate_up=col_parallel(2 * I, H, (I, I)),
                down=row_parallel(H, I),
            ))
        self.norm = torch.ones(H, dtype=dtype, device=dev)
        self.lm_head = rows(full(s.vocab, H), (s.vocab,)) if reuse is None else reuse.lm_head
        # rotary cache [max_pos, rot_dim] = cat(cos, sin) (modeling/layers/rotary_embedding.py:105-120)
        inv = 1.0 / (s.rope_theta ** (torch.arange(0, s.head_size, 2, dtype=torch.float32) / s.head_size))
        fr = torch.einsum("i,j->ij", torch.arange(s.max_position, dtype=torch.float32), inv)
        self.cos_sin = torch.cat((fr.cos(), fr.sin()), dim=-1).to(dtype).to(dev)

    def _init_kv_cache(self, num_blocks: int, kv_fill: bool, seed: int):
        """One [2, num_blocks, block_size*kv_heads*head_size] tensor per layer (aphrodite/worker/cache_engine.py:65-84);
        under TP a rank holds its kv-heads' slice of the full cache (per-rank num_kv_heads, cache_engine.py:37-38)."""
        s, dev = self.s, self.device
        store = self.dtype if self.kv_cache_dtype == "auto" else torch.uint8
        g = torch.Generator(device=dev).manual_seed(seed + 1)
        self.kv_caches = []
        per_head = self.block_size * s.head_size
        for _ in range(self.n_layers):
            shape = self.attn.get_kv_cache_shape(num_blocks, self.block_size, self.kv_heads, s.head_size)
            if not kv_fill:
                kv = torch.empty(shape, dtype=store, device=dev)
            else:
                fullkv = torch.empty(2, num_blocks, s.kv_heads, per_head, dtype=store, device=dev)
                if store == torch.uint8:
                    # fp8 cache = convert_fp8 of a U(-s, s) 16-bit cache, as the reference's benchmarks fill theirs
                    # (aphrodite/common/utils.py:604-621), in slabs of blocks to bound the 16-bit temporary
                    step = max(1, (1 << 28) // (s.kv_heads * per_head))
                    for plane in range(2):
                        for b0 in range(0, num_blocks, step):
                            dst = fullkv[plane, b0:b0 + step]
                            src = torch.empty(dst.shape, dtype=self.dtype, device=dev).uniform_(-self.scale, self.scale,
                                                                                               generator=g)
                            ops.convert_fp8(dst, src, 1.0, self.kv_cache_dtype)
                else:
                    fullkv.uniform_(-self.scale, self.scale, generator=g)
                if self.tp_size > 1:
                    kv = fullkv[:, :, self.tp_rank * self.kv_heads:(self.tp_rank + 1) * self.kv_heads].reshape(shape).contiguous()
                    del fullkv
                else:
                    kv = fullkv.view(shape)
            self.kv_caches.append(kv)
        self.kv_views = [self.attn.split_kv_cache(kv, self.kv_heads, s.head_size) for kv in self.kv_caches]

    # ------------------------------------------------------------------------------------------------------------
    def _linear(self, x, wt, out=None):
        """F.linear for bf16 weights (cuBLAS, as the reference's UnquantizedLinearMethod) or the Marlin-format
        W4A16 GEMM (apply_gptq_marlin_linear, aphrodite/quantization/utils/marlin_utils.py:241-275). `out`: the
        symmetric buffer a row-parallel result must land in for the fused exchange."""
        if isinstance(wt, dict):
            y = self.ops.gptq_marlin_gemm(x, wt["q"], wt["s"], self._empty, self._empty, self._empty, wt["ws"],
                                          self._qtype, x.shape[0], wt["n"], wt["k"], True, False, True, False)
            if out is not None:
                out.copy_(y)
                return out
            return y
        if out is not None:
            return torch.matmul(x, wt.t(), out=out)
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

    def _row_parallel_then_norm(self, x, wt, residual, ln_weight):
        """Row-parallel GEMM -> sum over ranks -> residual add -> RMSNorm. Returns the normed hidden state; `residual`
        is updated in place (tp_mode "nvls": only this rank's rows of it, the stream is token-sharded)."""
        if self.tp_mode == "nvls":
            self._linear(x, wt, out=self.nvls.x(self.batch))
            return self.nvls.allreduce_add_rms_norm(self.batch, residual, ln_weight, self.s.rms_eps)
        hidden = self._all_reduce(self._linear(x, wt))
        self.ops.fused_add_rms_norm(hidden, residual, ln_weight, self.s.rms_eps)
        return hidden

    def forward(self, st: DecodeState) -> torch.Tensor:
        s, o = self.s, self.ops
        hidden = F.embedding(st.input_ids, self.embed)
        residual = hidden
        normed = torch.empty_like(hidden)
        o.rms_norm(normed, hidden, self.layers[0]["ln1"], s.rms_eps)
        hidden = normed
        for li, L in enumerate(self.layers):
            qkv = self._linear(hidden, L["qkv"])
            q, k, v = qkv.split([self.q_size, self.kv_size, self.kv_size], dim=-1)
            kc, vc = self.kv_views[li]
            if self.fuse_rope_cache:
                ext_ops.rotary_embedding_and_cache(st.positions, q, k, v, s.head_size, self.cos_sin, True, kc, vc,
                                                   st.slot_mapping, self.kv_cache_dtype, 1.0, 1.0)
            else:
                o.rotary_embedding(st.positions, q, k, s.head_size, self.cos_sin, True)
                self.attn.write_to_paged_cache(k.view(-1, self.kv_heads, s.head_size),
                                               v.view(-1, self.kv_heads, s.head_size), kc, vc,
                                               st.slot_mapping, self.kv_cache_dtype, 1.0, 1.0)
            qv = q.view(-1, self.heads, s.head_size)
            attn_out = torch.empty(self.batch, self.heads, s.head_size, dtype=self.dtype,
                                   device=self.device)
            if self.attn_hook is not None:
                self.attn_hook(li, True)
            self.attn.forward_decode(qv, kc, vc, st.block_tables, st.seq_lens, st.max_seq_len,
                                     self.kv_cache_dtype, self.kv_heads, self.scale, None, 1.0,
                                     1.0, output=attn_out)
            if self.attn_hook is not None:
                self.attn_hook(li, False)
            hidden = self._row_parallel_then_norm(attn_out.view(self.batch, -1), L["o"], residual, L["ln2"])
            gate_up = self._linear(hidden, L["gate_up"])
            act = torch.empty(self.batch, self.inter, dtype=self.dtype, device=self.device)
            o.silu_and_mul(act, gate_up)
            next_ln = self.layers[li + 1]["ln1"] if li + 1 < self.n_layers else self.norm
            hidden = self._row_parallel_then_norm(act, L["down"], residual, next_ln)
        self.last_hidden = hidden
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
