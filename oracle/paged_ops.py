"""CPU restatement of the reference's decode-path algorithms — TEST INFRASTRUCTURE ONLY.

Nothing under aphrodite_engine_b200/ may import this package. It is used by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs as the CHECKER
(and as the timed CPU baseline when oracle/_ref is unusable on the host), never as the product.

Every function cites the reference file:line it restates (paths under the reference tree).
Arithmetic is done in float32 on CPU tensors with the reference's rounding points made explicit.

Pinning: these restatements are checked against (a) the reference's own CPU kernels compiled from
/root/reference/kernels/cpu into oracle/_ref (tests/test_oracle_cpu.py, live when the .so loads)
and (b) committed golden vectors generated from those kernels (tests/golden/*.npz, made by
tests/golden/make_golden.py). The reference ships no stored golden vectors for this path
(SURVEY.md §8c); its tests are generative with the tolerances repeated in tests/tolerances.py.
"""
from typing import List, Optional, Tuple

import torch

PARTITION_SIZE = 512  # aphrodite/attention/ops/paged_attn.py:13


# ------------------------------------------------------------------------------------------------
# fp8 (kernels/quantization/fp8/nvidia/quant_utils.cuh:296-360 dequant, :466-498 quant)
# ------------------------------------------------------------------------------------------------
def _fp8_torch_dtype(kv_cache_dtype: str):
    if kv_cache_dtype in ("fp8", "fp8_e4m3", "auto"):
        return torch.float8_e4m3fn
    if kv_cache_dtype == "fp8_e5m2":
        return torch.float8_e5m2
    raise ValueError(f"Unsupported data type of kv cache: {kv_cache_dtype}")


def fp8_dequant(u8: torch.Tensor, scale: float, out_dtype: torch.dtype,
                kv_cache_dtype: str = "fp8") -> torch.Tensor:
    """fp8 -> half (exact) -> float * scale -> out_dtype (round-to-nearest-even)."""
    f = u8.view(_fp8_torch_dtype(kv_cache_dtype)).to(torch.float16).to(torch.float32)
    return (f * scale).to(out_dtype)


def fp8_quant(x: torch.Tensor, scale: float, kv_cache_dtype: str = "fp8") -> torch.Tensor:
    """float(x) / scale -> fp8 with saturation to the largest finite value (__NV_SATFINITE)."""
    dt = _fp8_torch_dtype(kv_cache_dtype)
    f = x.to(torch.float32) / scale
    fmax = torch.finfo(dt).max
    f = torch.where(torch.isnan(f), f, f.clamp(-fmax, fmax))
    return f.to(dt).view(torch.uint8)


def convert_fp8(dst: torch.Tensor, src: torch.Tensor, scale: float = 1.0,
                kv_cache_dtype: str = "fp8") -> None:
    """kernels/cache_kernels.cu:334-410 (direction is chosen by which side is uint8)."""
    if src.dtype == torch.uint8:
        dst.copy_(fp8_dequant(src, scale, dst.dtype, kv_cache_dtype))
    else:
        dst.copy_(fp8_quant(src, scale, kv_cache_dtype))


# ------------------------------------------------------------------------------------------------
# KV cache layout helpers (aphrodite/common/utils.py:686-738; attention/ops/paged_attn.py:49-62)
# ------------------------------------------------------------------------------------------------
def gather_kv(key_cache: torch.Tensor, value_cache: torch.Tensor, block_table: List[int],
              seq_len: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """key_cache [NB,Hkv,D/x,BS,x], value_cache [NB,Hkv,D,BS] -> keys, values [seq_len,Hkv,D]."""
    nb, hkv, dx, bs, x = key_cache.shape
    d = dx * x
    nblk = (seq_len + bs - 1) // bs
    idx = torch.as_tensor(block_table[:nblk], dtype=torch.long)
    k = key_cache[idx]                                    # [nblk,Hkv,D/x,BS,x]
    k = k.permute(0, 3, 1, 2, 4).reshape(nblk * bs, hkv, d)[:seq_len]
    v = value_cache[idx]                                  # [nblk,Hkv,D,BS]
    v = v.permute(0, 3, 1, 2).reshape(nblk * bs, hkv, d)[:seq_len]
    return k, v


def _blocksparse_mask(seq_len: int, block_size: int, head: int, kv_head: int, num_heads: int,
                      num_kv_heads: int, tp_rank: int, local_blocks: int, vert_stride: int,
                      bs_block_size: int, head_sliding_step: int) -> torch.Tensor:
    """kernels/attention/attention_kernels.cu:210-257 — True where the token's KV block is attended."""
    q_bs = (seq_len - 1) // bs_block_size
    if head_sliding_step >= 0:
        off = (tp_rank * num_heads + head) * head_sliding_step + 1
    else:
        off = (tp_rank * num_kv_heads + kv_head) * (-head_sliding_step) + 1
    tok = torch.arange(seq_len)
    kb = (tok // block_size) * block_size // bs_block_size
    remote = (kb + off) % vert_stride == 0
    local = kb > q_bs - local_blocks
    return remote | local


def paged_attention(
    query: torch.Tensor,            # [S,Hq,D]
    key_cache: torch.Tensor,        # [NB,Hkv,D/x,BS,x]
    value_cache: torch.Tensor,      # [NB,Hkv,D,BS]
    block_tables: torch.Tensor,     # int32 [S,max_blocks]
    seq_lens: torch.Tensor,         # int32 [S]
    scale: float,
    alibi_slopes: Optional[torch.Tensor] = None,
    kv_cache_dtype: str = "auto",
    k_scale: float = 1.0,
    v_scale: float = 1.0,
    tp_rank: int = 0,
    blocksparse_local_blocks: int = 0,
    blocksparse_vert_stride: int = 0,
    blocksparse_block_size: int = 64,
    blocksparse_head_sliding_step: int = 0,
    round_probs: bool = True,
) -> torch.Tensor:
    """softmax(scale * q.K^T [+ alibi]) . V per (seq, head) over the block-table-indexed cache.

    Restates kernels/attention/attention_kernels.cu:87-496: logits and softmax in fp32
    (:293-343, normaliser sum + 1e-6), probabilities rounded to the activation dtype before P.V
    (:395-397), fp32 accumulation; fp8 K/V dequantised to the activation dtype first (:271-276).
    v1 and v2 compute the same function (v2 only partitions the sequence, :529-669).
    """
    S, Hq, D = query.shape
    Hkv = value_cache.shape[1]
    BS = value_cache.shape[3]
    G = Hq // Hkv
    dt = query.dtype
    out = torch.zeros(S, Hq, D, dtype=dt)
    if kv_cache_dtype != "auto":
        key_cache = fp8_dequant(key_cache, k_scale, dt, kv_cache_dtype)
        value_cache = fp8_dequant(value_cache, v_scale, dt, kv_cache_dtype)
    bt = block_tables.tolist()
    sl = seq_lens.tolist()
    sparse = blocksparse_vert_stride is not None and blocksparse_vert_stride > 1
    for i in range(S):
        n = int(sl[i])
        if n == 0:
            continue
        k, v = gather_kv(key_cache, value_cache, bt[i], n)      # [n,Hkv,D]
        kf = k.float().repeat_interleave(G, dim=1)                # [n,Hq,D]
        vf = v.float().repeat_interleave(G, dim=1)
        logits = scale * torch.einsum("hd,nhd->hn", query[i].float(), kf)
        if alibi_slopes is not None:
            pos = (torch.arange(n) - n + 1).float()
            logits = logits + alibi_slopes.float().view(-1, 1) * pos.view(1, -1)
        if sparse:
            for h in range(Hq):
                m = _blocksparse_mask(n, BS, h, h // G, Hq, Hkv, tp_rank, blocksparse_local_blocks,
                                      blocksparse_vert_stride, blocksparse_block_size,
                                      blocksparse_head_sliding_step)
                logits[h, ~m] = float("-inf")
        mx = logits.max(dim=-1, keepdim=True).values
        e = torch.exp(logits - mx)
        p = e / (e.sum(dim=-1, keepdim=True) + 1e-6)
        if round_probs:
            p = p.to(dt).float()
        out[i] = torch.einsum("hn,nhd->hd", p, vf).to(dt)
    return out


# ------------------------------------------------------------------------------------------------
# cache writers / movers (kernels/cache_kernels.cu)
# ------------------------------------------------------------------------------------------------
def reshape_and_cache(key: torch.Tensor, value: torch.Tensor, key_cache: torch.Tensor,
                      value_cache: torch.Tensor, slot_mapping: torch.Tensor,
                      kv_cache_dtype: str = "auto", k_scale: float = 1.0,
                      v_scale: float = 1.0) -> None:
    """kernels/cache_kernels.cu:152-204: K index :184-187, V index :188-191, slot < 0 skipped :166."""
    T, H, D = key.shape
    BS, x = key_cache.shape[3], key_cache.shape[4]
    slots = slot_mapping.tolist()
    if kv_cache_dtype != "auto":
        key = fp8_quant(key, k_scale, kv_cache_dtype)
        value = fp8_quant(value, v_scale, kv_cache_dtype)
    for t in range(T):
        s = int(slots[t])
        if s < 0:
            continue
        b, o = s // BS, s % BS
        key_cache[b, :, :, o, :] = key[t].reshape(H, D // x, x)
        value_cache[b, :, :, o] = value[t]


def reshape_and_cache_flash(key, value, key_cache, value_cache, slot_mapping,
                            kv_cache_dtype: str = "auto", k_scale: float = 1.0,
                            v_scale: float = 1.0) -> None:
    """kernels/cache_kernels.cu:206-245: cache layout [NB, BS, H, D]."""
    T = key.shape[0]
    BS = key_cache.shape[1]
    slots = slot_mapping.tolist()
    if kv_cache_dtype != "auto":
        key = fp8_quant(key, k_scale, kv_cache_dtype)
        value = fp8_quant(value, v_scale, kv_cache_dtype)
    for t in range(T):
        s = int(slots[t])
        if s < 0:
            continue
        key_cache[s // BS, s % BS] = key[t]
        value_cache[s // BS, s % BS] = value[t]


def copy_blocks(key_caches: List[torch.Tensor], value_caches: List[torch.Tensor],
                block_mapping: torch.Tensor) -> None:
    """kernels/cache_kernels.cu:67-99: for every layer, dst block := src block (K and V)."""
    for src, dst in block_mapping.tolist():
        for kc in key_caches:
            kc[dst].copy_(kc[src])
        for vc in value_caches:
            vc[dst].copy_(vc[src])


def swap_blocks(src: torch.Tensor, dst: torch.Tensor, block_mapping: torch.Tensor) -> None:
    """kernels/cache_kernels.cu:24-63."""
    for s, d in block_mapping.tolist():
        dst[d].copy_(src[s])


# ------------------------------------------------------------------------------------------------
# normalisation / rotary / activations
# ------------------------------------------------------------------------------------------------
def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """kernels/layernorm_kernels.cu:24-50: out = T(T(x * rsqrt(mean(x^2)+eps)) * w)."""
    xf = x.float()
    s = torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)
    y = (xf * s).to(x.dtype)
    return (y.float() * weight.float()).to(x.dtype)


def fused_add_rms_norm(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor,
                       eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """kernels/layernorm_kernels.cu:204-286: z = T(x + r) (:231-234), variance from z, out as rms_norm.
    Returns (new x, new residual)."""
    z = (x.float() + residual.float()).to(x.dtype)
    return rms_norm(z, weight, eps), z


def rotary_embedding(positions: torch.Tensor, query: torch.Tensor, key: torch.Tensor,
                     head_size: int, cos_sin_cache: torch.Tensor, is_neox: bool,
                     rot_dim: Optional[int] = None,
                     cos_sin_cache_offsets: Optional[torch.Tensor] = None
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """kernels/pos_encoding_kernels.cu:10-93: every *, -, + rounds to the tensor dtype (:31-34).
    query [T, Hq*D], key [T, Hkv*D]; returns rotated copies."""
    dt = query.dtype
    rot = cos_sin_cache.shape[1] if rot_dim is None else rot_dim
    emb = rot // 2
    pos = positions.flatten().long()
    if cos_sin_cache_offsets is not None:
        pos = pos + cos_sin_cache_offsets.flatten().long()
    cs = cos_sin_cache[pos]                       # [T, rot]
    cos, sin = cs[:, :emb], cs[:, emb:rot]

    def mul(a, b):
        return (a.float() * b.float()).to(dt)

    def rot_one(t):
        T = t.shape[0]
        th = t.reshape(T, -1, head_size).clone()
        c, s = cos.unsqueeze(1), sin.unsqueeze(1)
        if is_neox:
            xs, ys = th[..., :emb], th[..., emb:rot]
        else:
            xs, ys = th[..., 0:rot:2], th[..., 1:rot:2]
        nx = (mul(xs, c).float() - mul(ys, s).float()).to(dt)
        ny = (mul(ys, c).float() + mul(xs, s).float()).to(dt)
        if is_neox:
            th[..., :emb], th[..., emb:rot] = nx, ny
        else:
            th[..., 0:rot:2], th[..., 1:rot:2] = nx, ny
        return th.reshape(t.shape)

    return rot_one(query), rot_one(key)


def _tmul(a, b, dt):
    return (a.float() * b.float()).to(dt)


def silu_and_mul(x: torch.Tensor) -> torch.Tensor:
    """kernels/activation_kernels.cu:13-31: T(T(x / (1 + exp(-x))) * y)."""
    d = x.shape[-1] // 2
    a, b = x[..., :d], x[..., d:]
    af = a.float()
    return _tmul((af / (1.0 + torch.exp(-af))).to(x.dtype), b, x.dtype)


def gelu_and_mul(x: torch.Tensor) -> torch.Tensor:
    """kernels/activation_kernels.cu:33-41."""
    d = x.shape[-1] // 2
    a, b = x[..., :d], x[..., d:]
    af = a.float()
    g = (af * 0.5 * (1.0 + torch.erf(af * 0.7071067811865476))).to(x.dtype)
    return _tmul(g, b, x.dtype)


def gelu_tanh_and_mul(x: torch.Tensor) -> torch.Tensor:
    """kernels/activation_kernels.cu:43-52."""
    d = x.shape[-1] // 2
    a, b = x[..., :d], x[..., d:]
    af = a.float()
    beta = 1.4142135623730951 * 1.1283791670955126 * 0.5
    g = (0.5 * af * (1.0 + torch.tanh(beta * (af + 0.044715 * af * af * af)))).to(x.dtype)
    return _tmul(g, b, x.dtype)


def gelu_new(x: torch.Tensor) -> torch.Tensor:
    """kernels/activation_kernels.cu:120-125 (operations in T round at each step)."""
    dt = x.dtype
    x3 = _tmul(_tmul(x, x, dt), x, dt).float()
    inner = (x.float() + (0.044715 * x3).to(dt).float()).to(dt)
    t = torch.tanh((0.79788456 * inner.float()).to(dt).float()).to(dt)
    half_x = _tmul(torch.tensor(0.5, dtype=dt), x, dt)
    return _tmul(half_x, (1.0 + t.float()).to(dt), dt)


def gelu_fast(x: torch.Tensor) -> torch.Tensor:
    """kernels/activation_kernels.cu:127-134."""
    dt = x.dtype
    f = x.float()
    a = (f * 0.79788456).to(dt)
    b = (1.0 + _tmul((0.044715 * f).to(dt), x, dt).float()).to(dt)
    t = torch.tanh(_tmul(a, b, dt).float()).to(dt)
    half_x = _tmul(torch.tensor(0.5, dtype=dt), x, dt)
    return _tmul(half_x, (1.0 + t.float()).to(dt), dt)


def gelu_quick(x: torch.Tensor) -> torch.Tensor:
    """kernels/activation_kernels.cu:136-140."""
    f = x.float()
    return (f / (1.0 + torch.exp(-1.702 * f))).to(x.dtype)


# ------------------------------------------------------------------------------------------------
# synthetic inputs (tests/kernels/test_attention.py:147-173, aphrodite/common/utils.py:686-738)
# ------------------------------------------------------------------------------------------------
def make_kv_cache(num_blocks: int, block_size: int, num_kv_heads: int, head_size: int,
                  dtype: torch.dtype, kv_cache_dtype: str = "auto", seed: int = 0,
                  device: str = "cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    g = torch.Generator(device="cpu").manual_seed(seed)
    s = head_size ** -0.5
    store = dtype if kv_cache_dtype == "auto" else torch.uint8
    x = 16 // torch.tensor([], dtype=store).element_size()
    kshape = (num_blocks, num_kv_heads, head_size // x, block_size, x)
    vshape = (num_blocks, num_kv_heads, head_size, block_size)
    k = torch.empty(kshape, dtype=torch.float32).uniform_(-s, s, generator=g)
    v = torch.empty(vshape, dtype=torch.float32).uniform_(-s, s, generator=g)
    if kv_cache_dtype == "auto":
        k, v = k.to(dtype), v.to(dtype)
    else:  # fp8 cache produced by quantising a half tensor (utils.py:604-621)
        k = fp8_quant(k.to(torch.float16), 1.0, kv_cache_dtype)
        v = fp8_quant(v.to(torch.float16), 1.0, kv_cache_dtype)
    return k.to(device), v.to(device)


# ------------------------------------------------------------------------------------------------
# MoE routing (kernels/moe/align_block_size_kernel.cu:22-110, kernels/moe/softmax.cu:107-161)
# ------------------------------------------------------------------------------------------------
def moe_align_block_size(topk_ids: torch.Tensor, num_experts: int, block_size: int,
                         max_num_tokens_padded: Optional[int] = None):
    """Stable counting sort of the flat (token, k) slots by expert, every expert segment padded to a
    multiple of block_size. Returns (sorted_token_ids, expert_ids, num_tokens_post_pad) with untouched
    entries equal to the caller-side presets (numel / -1), as fused_moe.py:214-228 allocates them."""
    flat = topk_ids.flatten().tolist()
    numel = len(flat)
    if max_num_tokens_padded is None:
        max_num_tokens_padded = numel + num_experts * (block_size - 1)
    sorted_ids = torch.full((max_num_tokens_padded,), numel, dtype=torch.int32)
    max_blocks = (max_num_tokens_padded + block_size - 1) // block_size
    expert_ids = torch.full((max_blocks,), -1, dtype=torch.int32)
    off = 0
    for e in range(num_experts):
        idx = [i for i, v in enumerate(flat) if v == e]       # increasing order == stable
        sorted_ids[off:off + len(idx)] = torch.tensor(idx, dtype=torch.int32)
        padded = (len(idx) + block_size - 1) // block_size * block_size
        expert_ids[off // block_size:(off + padded) // block_size] = e
        off += padded
    return sorted_ids, expert_ids, torch.tensor([off], dtype=torch.int32)


def topk_softmax(gating_output: torch.Tensor, topk: int):
    """softmax over experts in fp32, then k rounds of arg-max (ties -> lowest expert id);
    weights are the un-renormalised probabilities; token_expert_indices[t, k] = k * T + t."""
    T, E = gating_output.shape
    p = torch.softmax(gating_output.float(), dim=-1)
    w = torch.empty(T, topk, dtype=torch.float32)
    ids = torch.empty(T, topk, dtype=torch.int32)
    src = torch.empty(T, topk, dtype=torch.int32)
    work = p.clone()
    for k in range(topk):
        best = work.max(dim=-1).values
        first = (work == best.unsqueeze(-1)).float().argmax(dim=-1)   # lowest index among ties
        w[:, k] = best
        ids[:, k] = first.int()
        src[:, k] = k * T + torch.arange(T, dtype=torch.int32)
        work[torch.arange(T), first] = -1.0
    return w, ids, src
