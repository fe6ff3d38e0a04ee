"""CPU restatement of the SURVEY §8(f) rows — TEST INFRASTRUCTURE ONLY (never imported by aphrodite_engine_b200/).

  prefix-aware prefill attention   aphrodite/attention/ops/prefix_prefill.py:13-255 (_fwd_kernel), :448-694 (alibi)
  fp8 activation quantisation      kernels/quantization/fp8/common.cu:46-62, 71-105, 182-256
  W8A8 scaled GEMM                 kernels/quantization/cutlass_w8a8/scaled_mm_entry.cu:92-140 (+ the epilogue of
                                   scaled_mm_c3x.cu: a_scales * (b_scales * acc) [+ bias])
  sampling kernels                 kernels/sampling/sampling.cuh:186-690 (rejection samplers), :909-1310 (renorm / mask)

Pinning: the prefill restatement is checked against golden vectors produced by the reference's OWN Triton kernel run under
Triton's CPU interpreter in the build container (tests/golden/make_golden_prefill.py -> tests/golden/prefill_*.npz);
fp8 quantisation and the sampling kernels are checked on the GPU against the reference's own CUDA kernels recompiled for
sm_100a (oracle/build_ref_cuda.py, tests/test_gpu_vs_ref_cuda.py). The W8A8 GEMM of the reference is a CUTLASS build that
needs a network fetch (CMakeLists.txt:231-241) and has no sm_100 kernel: treated as unbuildable; its restatement follows
the reference's own test baseline (tests/kernels/test_cutlass.py: baseline_scaled_mm) — "parity unpinned" for that op
beyond the arithmetic definition.
"""
import math
from typing import Optional

import numpy as np
import torch

from . import paged_ops as po

FP8_MAX = 448.0


# ----------------------------------------------------------------------------------------------------------------------
# prefill attention over the paged cache (prefix_prefill.py)
# ----------------------------------------------------------------------------------------------------------------------
def context_attention(q, k, v, key_cache, value_cache, block_tables, start_loc, seq_lens, ctx_lens, kv_cache_dtype="auto",
                      k_scale=1.0, v_scale=1.0, alibi_slopes=None, sliding_window=0):
    """q [T, Hq, D], k / v [T, Hkv, D]; key_cache [NB, Hkv, D/x, BS, x], value_cache [NB, Hkv, D, BS] (uint8 when fp8).
    Returns out [T, Hq, D] in q's dtype. fp32 softmax, probabilities rounded to q's dtype before P.V like the kernel
    (prefix_prefill.py:170 `p = p.to(v.dtype)`); scale 1/sqrt(D) (:742)."""
    T, Hq, D = q.shape
    Hkv = k.shape[1]
    qpk = Hq // Hkv
    NB, _, _, BS, x = key_cache.shape
    dt = q.dtype
    out = torch.zeros_like(q)
    scale = 1.0 / math.sqrt(D)
    sw = sliding_window if sliding_window and sliding_window > 0 else 0
    for b in range(len(seq_lens)):
        ctx, seq = int(ctx_lens[b]), int(seq_lens[b])
        ql, s0 = seq - ctx, int(start_loc[b])
        if ql <= 0:
            continue
        # gather the context from the paged layouts
        pos = torch.arange(ctx)
        blk = block_tables[b][pos // BS].long()
        off = pos % BS
        kc = key_cache[blk, :, :, off, :]                      # [ctx, Hkv, D/x, x]
        kc = kc.reshape(ctx, Hkv, D)
        vc = value_cache[blk, :, :, off]                       # [ctx, Hkv, D]
        if kv_cache_dtype != "auto":
            kc = po.fp8_dequant(kc.contiguous(), k_scale, dt, kv_cache_dtype)       # fp8 -> fp32 * scale -> q.dtype (:126-129)
            vc = po.fp8_dequant(vc.contiguous(), v_scale, dt, kv_cache_dtype)
        keys = torch.cat([kc.to(dt), k[s0:s0 + ql]], 0).float()    # [seq, Hkv, D]
        vals = torch.cat([vc.to(dt), v[s0:s0 + ql]], 0).float()
        qpos = ctx + torch.arange(ql)
        kpos = torch.arange(seq)
        rel = kpos[None, :] - qpos[:, None]                    # <= 0 where visible
        for h in range(Hq):
            kvh = h // qpk
            s = (q[s0:s0 + ql, h].float() @ keys[:, kvh].T) * scale       # [ql, seq]
            if alibi_slopes is not None:
                s = s + float(alibi_slopes[h]) * rel.float()
            if sw:
                s = torch.where(-rel < sw, s, torch.full_like(s, -10000.0))   # finite mask of the reference (:147-149)
            s = torch.where(rel <= 0, s, torch.full_like(s, float("-inf")))
            p = torch.softmax(s, dim=-1)
            out[s0:s0 + ql, h] = (p.to(dt).float() @ vals[:, kvh]).to(dt)
    return out


# ----------------------------------------------------------------------------------------------------------------------
# fp8 activation quantisation (fp8/common.cu)
# ----------------------------------------------------------------------------------------------------------------------
def _to_e4m3(x: torch.Tensor) -> torch.Tensor:
    r = torch.clamp(x, -FP8_MAX, FP8_MAX)                       # fmax(-MAX, fmin(x, MAX)) (:55)
    r = torch.where(torch.isnan(x), torch.full_like(x, FP8_MAX), r)   # fmin(NaN, MAX) = MAX
    return r.to(torch.float8_e4m3fn)


def static_scaled_fp8_quant(x: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    inv = (torch.ones(1, dtype=torch.float32) / scale.float().reshape(1))          # one IEEE division (:191)
    return _to_e4m3(x.float() * inv)


def dynamic_scaled_fp8_quant(x: torch.Tensor):
    scale = (x.float().abs().max() / FP8_MAX).reshape(1)                            # segmented_max_reduction (:71-105)
    return static_scaled_fp8_quant(x, scale), scale


def dynamic_per_token_scaled_fp8_quant(x: torch.Tensor, scale_ub: Optional[torch.Tensor] = None):
    amax = x.float().abs().amax(dim=-1, keepdim=True)
    if scale_ub is not None:
        amax = torch.minimum(amax, scale_ub.float().reshape(1, 1))
    scale = torch.maximum(amax / FP8_MAX, torch.tensor(1.0 / (FP8_MAX * 512.0)))    # (:226-237)
    return _to_e4m3(x.float() / scale), scale                                       # true division (:246-255)


# ----------------------------------------------------------------------------------------------------------------------
# W8A8 scaled GEMM (scaled_mm_entry.cu / scaled_mm_c3x.cu epilogue)
# ----------------------------------------------------------------------------------------------------------------------
def scaled_mm(a: torch.Tensor, b: torch.Tensor, a_scales: torch.Tensor, b_scales: torch.Tensor, out_dtype,
              bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a [M, K] (fp8 / int8), b [K, N]; fp64 accumulation stands in for the exact sum, epilogue in fp32 in the
    reference's order: a_scale * (b_scale * acc) (+ bias, fused multiply-add)."""
    acc = (a.double() @ b.double()).float()
    tmp = b_scales.float().reshape(1, -1) * acc
    o = a_scales.float().reshape(-1, 1) * tmp
    if bias is not None:
        o = (a_scales.double().reshape(-1, 1) * tmp.double() + bias.double().reshape(1, -1)).float()    # one rounding = fma
    return o.to(out_dtype)


# ----------------------------------------------------------------------------------------------------------------------
# sampling (sampling.cuh). float32 running sums in index order: the kernels' sums differ from these only by the
# association order of fp32 additions, so tests compare ids exactly except where a cdf value lands within a few ulp of u.
# ----------------------------------------------------------------------------------------------------------------------
def _sample_above(row: np.ndarray, pivot: float, u: float):
    """first index i with row[i] > pivot whose inclusive running mass (entries > pivot) exceeds u; len-1 if none.
    Also returns the margin |cdf - u| at the decision (small margin = the kernel may legitimately pick a neighbour)."""
    y = np.where(row > pivot, row, 0.0).astype(np.float64)
    cdf = np.cumsum(y)
    hit = np.nonzero((cdf > u) & (row > pivot))[0]
    if hit.size == 0:
        return len(row) - 1, float(abs(cdf[-1] - u))
    i = int(hit[0])
    prev = cdf[i - 1] if i > 0 else 0.0
    return i, float(min(cdf[i] - u, u - prev if i > 0 else np.inf))


def sampling_from_probs(probs: np.ndarray, uniform: np.ndarray):
    ids, margins = [], []
    for b in range(probs.shape[0]):
        i, m = _sample_above(probs[b], 0.0, float(uniform[b]))
        ids.append(i)
        margins.append(m)
    return np.array(ids, dtype=np.int32), np.array(margins)


def rejection_sampling(mode: str, probs: np.ndarray, uniform: np.ndarray, k=None, p=None):
    """mode in {"top_k", "top_p", "min_p", "top_k_top_p"}; uniform [rounds, B]; k / p scalars or per-row arrays.
    Returns (ids, success, min margin over the rounds)."""
    B, V = probs.shape
    R = uniform.shape[0]
    ids = np.zeros(B, dtype=np.int32)
    ok = np.zeros(B, dtype=bool)
    margins = np.full(B, np.inf)
    for b in range(B):
        row = probs[b].astype(np.float32)
        kb = int(k[b]) if isinstance(k, np.ndarray) else (int(k) if k is not None else 0)
        pb = float(p[b]) if isinstance(p, np.ndarray) else (float(p) if p is not None else 0.0)
        q, pivot = np.float32(1.0), np.float32(0.0)
        scaled = np.float32(row.max()) * np.float32(pb) if mode == "min_p" else None
        sid = V - 1
        for r in range(R):
            u = np.float32(uniform[r, b]) * q
            sid, m = _sample_above(row, float(pivot), float(u))
            margins[b] = min(margins[b], m)
            pivot = max(pivot, row[sid])
            if mode == "min_p" and pivot >= scaled:
                ok[b] = True
                break
            above = row > pivot
            q = np.float32(row[above].astype(np.float64).sum())
            cnt = int(above.sum())
            if mode == "top_k" and cnt < kb:
                ok[b] = True
                break
            if mode in ("top_p", "top_k_top_p"):
                margins[b] = min(margins[b], abs(float(q) - pb))      # a mass within rounding of p may stop a round apart
            if mode == "top_p" and q < np.float32(pb):
                ok[b] = True
                break
            if mode == "top_k_top_p" and cnt < kb and q < np.float32(pb):
                ok[b] = True
                break
        ids[b] = sid
    return ids, ok, margins


def top_p_renorm_prob(probs: np.ndarray, p):
    """keep x >= t*, t* = the largest value with mass(x >= t*) >= p; divide by the kept mass (sampling.cuh:909-1040)."""
    out = np.zeros_like(probs, dtype=np.float32)
    for b in range(probs.shape[0]):
        row = probs[b].astype(np.float64)
        pb = float(p[b]) if isinstance(p, np.ndarray) else float(p)
        vals = np.unique(row)[::-1]                               # distinct values, descending
        mass = np.array([row[row >= t].sum() for t in vals]) if len(vals) <= 4096 else None
        if mass is None:
            order = np.sort(row)[::-1]
            cs = np.cumsum(order)
            idx = int(np.searchsorted(cs, pb, side="left"))
            idx = min(idx, len(order) - 1)
            t = order[idx]
            kept_mass = row[row >= t].sum()
            crossed = cs[-1] >= pb
        else:
            hit = np.nonzero(mass >= pb)[0]
            crossed = hit.size > 0
            t = vals[hit[0]] if crossed else 0.0
            kept_mass = mass[hit[0]] if crossed else 1.0
        if not crossed:                                           # mass never reaches p: keep x > 0, no renormalisation
            out[b] = np.where(row > 0, row, 0.0)
        else:
            out[b] = np.where(row >= t, row / max(kept_mass, 1e-8), 0.0)
    return out


def _kth_largest(row: np.ndarray, k: int):
    k = max(int(k), 1)                                           # k = 0 behaves as k = 1 in the kernel's bisection
    return np.sort(row)[::-1][k - 1]


def top_k_renorm_prob(probs: np.ndarray, k):
    out = np.zeros_like(probs, dtype=np.float32)
    V = probs.shape[1]
    for b in range(probs.shape[0]):
        row = probs[b].astype(np.float64)
        kb = int(k[b]) if isinstance(k, np.ndarray) else int(k)
        if kb >= V:
            out[b] = row
            continue
        t = _kth_largest(row, kb)
        out[b] = np.where(row >= t, row / max(row[row >= t].sum(), 1e-8), 0.0)
    return out


def top_k_mask_logits(logits: np.ndarray, k):
    out = np.array(logits, dtype=np.float32, copy=True)
    V = logits.shape[1]
    for b in range(logits.shape[0]):
        kb = int(k[b]) if isinstance(k, np.ndarray) else int(k)
        if kb >= V:
            continue
        t = _kth_largest(logits[b], kb)
        out[b] = np.where(logits[b] >= t, logits[b], -np.inf)
    return out
