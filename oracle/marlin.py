"""CPU restatement of the Marlin W4A16 / W8A16 weight formats and GEMM — TEST INFRASTRUCTURE ONLY.

Restates (reference paths):
  aphrodite/quantization/utils/quant_utils.py:123-205   quantize_weights
  aphrodite/quantization/utils/quant_utils.py:334-442   pack_rows / pack_cols / gptq_pack / awq_pack
  aphrodite/quantization/utils/marlin_utils_test.py:30-125  marlin tile permutation + packing, marlin_quantize
  aphrodite/quantization/utils/marlin_utils.py:172-238  scale / zero-point permutations
  kernels/quantization/gptq_marlin/gptq_marlin.cu:156-360 dequant numerics ((q - bias) exact, then * scale
                                                          rounded once to the activation dtype)
Pinned against the reference's own Python (imported from /root/reference in the build container) by
tests/golden/make_golden_marlin.py -> tests/golden/marlin_*.npz.
"""
from typing import Optional, Tuple

import numpy as np
import torch

TILE = 16  # GPTQ_MARLIN_TILE


def pack_factor(num_bits: int) -> int:
    return 32 // num_bits


def _interleave(num_bits: int) -> np.ndarray:
    if num_bits == 4:
        return np.array([0, 2, 4, 6, 1, 3, 5, 7])
    if num_bits == 8:
        return np.array([0, 2, 1, 3])
    raise ValueError(f"num_bits must be 4 or 8, got {num_bits}")


def quantize_weights(w: torch.Tensor, num_bits: int, group_size: int, bias: int,
                     zero_points: bool = False):
    """Symmetric-with-bias (GPTQ uint4b8 / uint8b128) or asymmetric (AWQ uint4 / uint8 + zp) group
    quantisation. Returns (w_ref, q_w int32 [K,N] stored values, scales [K/g,N], zp int32 [K/g,N] | None).
    w_ref = (q - zp_or_bias) * s computed in w.dtype, as quant_utils.py:176-181."""
    K, N = w.shape
    dt = w.dtype
    g = K if group_size == -1 else group_size
    wg = w.reshape(K // g, g, N)
    max_val = wg.max(dim=1, keepdim=True).values
    min_val = wg.min(dim=1, keepdim=True).values
    if zero_points:
        q_max, q_min = (1 << num_bits) - 1, 0
        s = (max_val - min_val).clamp(min=1e-5) / q_max
        zp = torch.round(torch.abs(min_val / s)).clamp(q_min, q_max).int()
        q = torch.round(wg / s).int() + zp
        q = torch.clamp(q, q_min, q_max)
        w_ref = (q - zp).to(dt) * s
        stored = q
    else:
        q_max, q_min = (1 << num_bits) - 1 - bias, -bias
        s = torch.max((max_val / q_max).abs(), (min_val / q_min).abs())
        q = torch.clamp(torch.round(wg / s).int(), q_min, q_max)
        w_ref = q.to(dt) * s
        stored = q + bias
        zp = None
    return (w_ref.reshape(K, N).contiguous(), stored.reshape(K, N).contiguous().int(),
            s.reshape(K // g, N).contiguous(), None if zp is None else zp.reshape(K // g, N).contiguous())


def pack_rows(q_w: torch.Tensor, num_bits: int) -> torch.Tensor:
    """GPTQ checkpoint layout: int32 [K/pf, N], pf consecutive k per word (quant_utils.py:334-354)."""
    pf = pack_factor(num_bits)
    q = q_w.numpy().astype(np.uint32)
    out = np.zeros((q.shape[0] // pf, q.shape[1]), dtype=np.uint32)
    for i in range(pf):
        out |= q[i::pf, :] << (num_bits * i)
    return torch.from_numpy(out.astype(np.int32))


def pack_cols(q_w: torch.Tensor, num_bits: int) -> torch.Tensor:
    pf = pack_factor(num_bits)
    q = q_w.numpy().astype(np.uint32)
    out = np.zeros((q.shape[0], q.shape[1] // pf), dtype=np.uint32)
    for i in range(pf):
        out |= q[:, i::pf] << (num_bits * i)
    return torch.from_numpy(out.astype(np.int32))


def unpack_cols(p: torch.Tensor, num_bits: int, K: int, N: int) -> torch.Tensor:
    pf = pack_factor(num_bits)
    a = p.numpy().astype(np.uint32).copy()
    out = np.zeros((K, N), dtype=np.uint32)
    mask = (1 << num_bits) - 1
    for i in range(pf):
        out[:, i::pf] = a & mask
        a >>= num_bits
    return torch.from_numpy(out.astype(np.int32))


def awq_pack(q_w: torch.Tensor, num_bits: int) -> torch.Tensor:
    """AWQ checkpoint layout: int32 [K, N/pf] with the column interleave (quant_utils.py:425-442)."""
    K, N = q_w.shape
    il = _interleave(num_bits)
    q = q_w.reshape(-1, len(il))[:, il].reshape(K, N).contiguous()
    return pack_cols(q, num_bits)


def weight_perm(num_bits: int) -> torch.Tensor:
    """marlin_utils_test.py:65-92: thread i of a warp owns, for each of the four 16x16 tiles of a
    16x64 block, rows {2(i%4), +1, +8, +9} x columns {i/4, i/4+8} (its mma.sync B fragments)."""
    perm = []
    for i in range(32):
        p1 = []
        col = i // 4
        for block in (0, 1):
            for row in (2 * (i % 4), 2 * (i % 4) + 1, 2 * (i % 4 + 4), 2 * (i % 4 + 4) + 1):
                p1.append(16 * row + col + 8 * block)
        for j in range(4):
            perm.extend(p + 256 * j for p in p1)
    perm = np.array(perm)
    il = _interleave(num_bits)
    return torch.from_numpy(perm.reshape(-1, len(il))[:, il].ravel())


def marlin_weights(q_w: torch.Tensor, num_bits: int) -> torch.Tensor:
    """[K,N] stored values -> Marlin layout int32 [K/16, N*16/pf] (marlin_utils_test.py:30-62)."""
    K, N = q_w.shape
    perm = weight_perm(num_bits)
    q = q_w.reshape(K // TILE, TILE, N // TILE, TILE).permute(0, 2, 1, 3).reshape(K // TILE, N * TILE)
    q = q.reshape(-1, perm.numel())[:, perm].reshape(K // TILE, N * TILE).contiguous()
    return pack_cols(q, num_bits)


def scale_perms():
    sp = []
    for i in range(8):
        sp.extend(i + 8 * j for j in range(8))
    sps = []
    for i in range(4):
        sps.extend(2 * i + j for j in (0, 1, 8, 9, 16, 17, 24, 25))
    return sp, sps


def marlin_permute_scales(s: torch.Tensor, K: int, N: int, group_size: int) -> torch.Tensor:
    sp, sps = scale_perms()
    if group_size < K and group_size != -1:
        s = s.reshape(-1, len(sp))[:, sp]
    else:
        s = s.reshape(-1, len(sps))[:, sps]
    return s.reshape(-1, N).contiguous()


def marlin_zero_points(zp: torch.Tensor, K: int, N: int, num_bits: int) -> torch.Tensor:
    sp, _ = scale_perms()
    z = zp.reshape(-1, len(sp))[:, sp]
    il = _interleave(num_bits)
    z = z.reshape(-1, len(il))[:, il].reshape(-1, N).contiguous()
    return pack_cols(z, num_bits)


def awq_to_marlin_zero_points(q_zp_packed: torch.Tensor, K: int, N: int, num_bits: int) -> torch.Tensor:
    z = unpack_cols(q_zp_packed, num_bits, K, N)
    undo = np.argsort(_interleave(num_bits))
    z = z.reshape(-1, len(undo))[:, undo].reshape(-1, N).contiguous()
    return marlin_zero_points(z, K, N, num_bits)


def gptq_marlin_repack(b_q_weight: torch.Tensor, perm: Optional[torch.Tensor], K: int, N: int,
                       num_bits: int) -> torch.Tensor:
    """kernels/quantization/gptq_marlin/gptq_marlin_repack.cu:271-343: GPTQ [K/pf,N] (rows optionally
    gathered by `perm`, the act-order sort indices) -> Marlin layout."""
    pf = pack_factor(num_bits)
    a = b_q_weight.numpy().astype(np.uint32)
    mask = (1 << num_bits) - 1
    q = np.zeros((K, N), dtype=np.int32)
    for i in range(pf):
        q[i::pf, :] = (a >> (num_bits * i)) & mask
    q = torch.from_numpy(q)
    if perm is not None and perm.numel() > 0:
        q = q[perm.long()]
    return marlin_weights(q, num_bits)


def awq_marlin_repack(b_q_weight: torch.Tensor, K: int, N: int, num_bits: int) -> torch.Tensor:
    """kernels/quantization/gptq_marlin/awq_marlin_repack.cu:208-268: AWQ [K,N/pf] -> Marlin layout."""
    z = unpack_cols(b_q_weight, num_bits, K, N)
    undo = np.argsort(_interleave(num_bits))
    z = z.reshape(-1, len(undo))[:, undo].reshape(K, N).contiguous()
    return marlin_weights(z, num_bits)


def marlin_quantize(w: torch.Tensor, num_bits: int, group_size: int
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """GPTQ-style (uint4b8 / uint8b128, no act-order): returns (w_ref, marlin_q_w, marlin_scales)."""
    K, N = w.shape
    bias = 1 << (num_bits - 1)
    w_ref, q_w, s, _ = quantize_weights(w, num_bits, group_size, bias, False)
    g = K if group_size == -1 else group_size
    return w_ref, marlin_weights(q_w, num_bits), marlin_permute_scales(s, K, N, g)


def awq_marlin_quantize(w: torch.Tensor, num_bits: int, group_size: int):
    """AWQ-style (uint4 / uint8 + integer zero points): (w_ref, marlin_q_w, marlin_scales, marlin_zp)."""
    K, N = w.shape
    w_ref, q_w, s, zp = quantize_weights(w, num_bits, group_size, 0, True)
    g = K if group_size == -1 else group_size
    return (w_ref, marlin_weights(q_w, num_bits), marlin_permute_scales(s, K, N, g),
            marlin_zero_points(zp, K, N, num_bits))


def marlin_gemm(a: torch.Tensor, w_ref: torch.Tensor) -> torch.Tensor:
    """C = A . W with W already dequantised to the activation dtype; fp32 accumulation
    (tests/kernels/test_marlin_gemm.py:236-254 compares against `a @ w_ref`)."""
    return (a.float() @ w_ref.float()).to(a.dtype)


def marlin_quantize_act_order(w: torch.Tensor, num_bits: int, group_size: int, seed: int = 0):
    """GPTQ act-order checkpoint (aphrodite/quantization/utils/quant_utils.py:208-240 + :314-331, sort_weights;
    marlin_utils_test.py:95-125): rows of the quantised matrix are randomly permuted (simulated activation order),
    g_idx records each row's group, then rows are sorted by group for the kernel.
    Returns (w_ref [checkpoint row order], marlin_q_w, marlin_scales, g_idx_sorted int32, sort_indices int32)."""
    K, N = w.shape
    bias = 1 << (num_bits - 1)
    w_ref, q_w, s, _ = quantize_weights(w, num_bits, group_size, bias, False)
    g_idx = (torch.arange(K) // group_size).int()
    rand_perm = torch.randperm(K, generator=torch.Generator().manual_seed(seed))
    g_idx, q_w, w_ref = g_idx[rand_perm].contiguous(), q_w[rand_perm].contiguous(), w_ref[rand_perm].contiguous()
    sort_indices = torch.argsort(g_idx, stable=True).int()
    q_sorted = q_w[sort_indices.long()].contiguous()
    return (w_ref, marlin_weights(q_sorted, num_bits), marlin_permute_scales(s, K, N, group_size),
            g_idx[sort_indices.long()].contiguous(), sort_indices)


def awq_dequantize(qweight: torch.Tensor, scales: torch.Tensor, zeros: torch.Tensor) -> torch.Tensor:
    """kernels/quantization/awq/gemm_kernels.cu:720-780: out[k,n] = fp16((q - z)) * s in fp16 arithmetic."""
    K, N8 = qweight.shape
    N = N8 * 8
    G = K // scales.shape[0]
    undo = np.argsort(_interleave(4))
    q = unpack_cols(qweight, 4, K, N).reshape(-1, 8)[:, undo].reshape(K, N)
    z = unpack_cols(zeros, 4, K // G, N).reshape(-1, 8)[:, undo].reshape(K // G, N)
    d = (q - z.repeat_interleave(G, dim=0)).to(torch.float16)
    return (d.float() * scales.repeat_interleave(G, dim=0).float()).to(torch.float16)


def hqq_marlin_quantize(q_w: torch.Tensor, s: torch.Tensor, zp: torch.Tensor, group_size: int):
    """HQQ float zero points (aphrodite/quantization/hqq_marlin.py:183-226): codes q [K,N] uint4, fp16 scale / zero
    [K/g, N]; W = fp16(fp16(q - zero) * scale) — the kernel's sub_zpf + scale chain (gptq_marlin.cu:393-403, 366-379).
    Returns (w_ref fp16, marlin_q_w, marlin_scales, marlin_zeros): zeros use the SCALE permutation."""
    K, N = q_w.shape
    s, zp = s.half(), zp.half()
    rep = lambda t: t.repeat_interleave(group_size, dim=0)
    w_ref = (q_w.half() - rep(zp)) * rep(s)
    return (w_ref, marlin_weights(q_w.int(), 4), marlin_permute_scales(s, K, N, group_size),
            marlin_permute_scales(zp, K, N, group_size))


def marlin_gemm_moe(a: torch.Tensor, w_refs, topk_ids: torch.Tensor, topk_weights: torch.Tensor,
                    replicate_input: bool, apply_weights: bool) -> torch.Tensor:
    """kernels/moe/marlin_moe_ops.cu:1482-1546 semantics with dequantised expert weights w_refs[e] [K, N]:
    row r = t * topk + j of C is a[t] (replicate_input) or a[r] times W[topk_ids[t, j]], rounded to the activation
    dtype, then optionally multiplied by topk_weights[t, j] in fp32 and rounded again (:945-956).
    Returns C [M, topk, N]; rows routed to an invalid expert stay zero."""
    M, topk = topk_ids.shape
    N = w_refs[0].shape[1]
    out = torch.zeros(M * topk, N, dtype=a.dtype)
    flat = topk_ids.reshape(-1).long()
    rows = torch.arange(M * topk)
    for e, w in enumerate(w_refs):
        sel = rows[flat == e]
        if sel.numel() == 0:
            continue
        src = a[sel // topk] if replicate_input else a[sel]
        o = (src.float() @ w.float()).to(a.dtype)
        if apply_weights:
            o = (topk_weights.reshape(-1)[sel].float()[:, None] * o.float()).to(a.dtype)
        out[sel] = o
    return out.view(M, topk, N)


def fused_marlin_moe(a: torch.Tensor, w1_refs, w2_refs, topk_weights: torch.Tensor,
                     topk_ids: torch.Tensor) -> torch.Tensor:
    """aphrodite/modeling/layers/fused_moe/fused_moe.py:529-542: gate_up GEMM, silu_and_mul, down GEMM with the
    routing weights applied, sum over the topk slots."""
    from oracle import paged_ops as po
    M, topk = topk_ids.shape
    gate_up = marlin_gemm_moe(a, w1_refs, topk_ids, topk_weights, True, False)
    act = po.silu_and_mul(gate_up.view(M * topk, -1))
    down = marlin_gemm_moe(act, w2_refs, topk_ids, topk_weights, False, True)
    return torch.sum(down, dim=1)


def mixtral_quant_moe(x: torch.Tensor, gate_w: torch.Tensor, w13_refs: dict, w2_refs: dict, topk: int,
                      experts) -> torch.Tensor:
    """aphrodite/modeling/models/mixtral_quant.py:128-152 (one rank's partial sum, before the all-reduce), with the
    experts' dequantised weights: w13_refs[e] [H, 2I] (w1 | w3 side by side), w2_refs[e] [I, H].
    router: softmax in fp32 -> top-k -> renormalise (:133-139); per local expert: dense MLP on ALL tokens (:82-88,
    SiLU(w1 x) * (w3 x) with every op rounding to the activation dtype), `.mul_(expert_weights)` (:148, the product
    formed in fp32, rounded once), `final.add_(current)` (:152, an activation-dtype add)."""
    from oracle import paged_ops as po
    dt = x.dtype
    logits = (x.float() @ gate_w.float().t()).to(dt)
    p = torch.softmax(logits.float(), dim=1)
    w, ids = torch.topk(p, topk, dim=-1)
    w = w / w.sum(dim=-1, keepdim=True)
    final = None
    for e in experts:
        gate_up = marlin_gemm(x, w13_refs[e])
        act = po.silu_and_mul(gate_up)
        cur = marlin_gemm(act, w2_refs[e])
        ew = (w * (ids == e)).sum(dim=-1, keepdim=True)
        cur = (cur.float() * ew).to(dt)
        final = cur if final is None else (final.float() + cur.float()).to(dt)
    return final
