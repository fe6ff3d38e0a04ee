"""GPU parity: grouped W4A16 GEMM (`_moe_C.marlin_gemm_moe`) and the fused Marlin MoE layer vs the CPU oracle
(oracle/marlin.py `marlin_gemm_moe` / `fused_marlin_moe`, semantics of kernels/moe/marlin_moe_ops.cu:1482-1546 and
aphrodite/modeling/layers/fused_moe/fused_moe.py:438-542; layer-level check follows tests/kernels/test_moe.py:15-29)."""
import pytest
import torch

from oracle import marlin as om
from oracle import paged_ops as po

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _experts(E, K, N, group_size, dtype, seed, act_order=False):
    """E quantised experts -> (w_refs list [K,N], q [E,K/16,2N] int32, scales [E,G,N], perm [E,K] or [E,0])."""
    g = torch.Generator().manual_seed(seed)
    refs, qs, ss, perms, gidx = [], [], [], [], []
    for e in range(E):
        w = (torch.randn(K, N, generator=g) * 0.1).to(dtype)
        if act_order:
            w_ref, mq, ms, g_sorted, sort_idx = om.marlin_quantize_act_order(w, 4, group_size, seed=seed + e)
            perms.append(sort_idx)
            gidx.append(g_sorted)
        else:
            w_ref, mq, ms = om.marlin_quantize(w, 4, group_size)
        refs.append(w_ref)
        qs.append(mq)
        ss.append(ms)
    perm = torch.stack(perms).int() if act_order else torch.empty(E, 0, dtype=torch.int32)
    g_idx = torch.stack(gidx).int() if act_order else torch.empty(E, 0, dtype=torch.int32)
    return refs, torch.stack(qs).contiguous(), torch.stack(ss).contiguous(), g_idx, perm


def _route(M, E, topk, seed, skew=False):
    g = torch.Generator().manual_seed(seed)
    gate = torch.randn(M, E, generator=g)
    if skew:
        gate[:, E // 2:] -= 20.0          # half of the experts receive nothing
    w, ids, _ = po.topk_softmax(gate, topk)
    return gate, w, ids.int()


def _check(out, ref, tol=1e-2):
    out, ref = out.float().cpu(), ref.float()
    assert torch.isfinite(out).all()
    scale = ref.abs().max().clamp(min=1e-6)
    assert ((out - ref).abs().max() / scale) < tol, float((out - ref).abs().max() / scale)
    assert ((out - ref).abs().mean() / ref.abs().mean().clamp(min=1e-6)) < tol / 4


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("cfg", [  # (M, K, N, E, topk, block, group, replicate, apply, skew)
    (1, 128, 64, 8, 2, 16, -1, True, False, False),
    (33, 256, 192, 8, 2, 16, 128, True, True, False),
    (128, 512, 256, 8, 2, 64, 128, True, False, True),
    (128, 512, 256, 8, 2, 64, 64, False, True, False),
    (300, 256, 128, 4, 2, 64, 32, True, False, True),
    (77, 1024, 320, 60, 6, 16, 128, False, True, False),
    (600, 128, 64, 2, 1, 64, -1, True, True, False),
])
def test_marlin_gemm_moe(ops, dtype, cfg):
    from aphrodite_engine_b200 import fused_moe as fm
    M, K, N, E, topk, block, group, replicate, apply_w, skew = cfg
    refs, q, s, g_idx, perm = _experts(E, K, N, group, dtype, seed=M + N)
    _, tw, ids = _route(M, E, topk, seed=K + E, skew=skew)
    g = torch.Generator().manual_seed(M)
    a = torch.randn(M if replicate else M * topk, K, generator=g).to(dtype)
    sorted_ids, _, _ = fm.moe_align_block_size(ids.to(DEV), block, E)
    ws = torch.zeros(16, dtype=torch.int32, device=DEV)
    out = torch.ops._moe_C.marlin_gemm_moe(a.to(DEV), q.to(DEV), sorted_ids, tw.to(DEV), ids.to(DEV), s.to(DEV),
                                           g_idx.to(DEV), perm.to(DEV), ws, M, N, K, True, E, topk, block,
                                           replicate, apply_w)
    torch.cuda.synchronize()
    assert out.shape == (M, topk, N) and out.dtype == dtype
    _check(out, om.marlin_gemm_moe(a, refs, ids, tw, replicate, apply_w))


def test_marlin_gemm_moe_act_order(ops):
    from aphrodite_engine_b200 import fused_moe as fm
    M, K, N, E, topk, block, group = 50, 512, 128, 4, 2, 16, 128
    dtype = torch.float16
    refs, q, s, g_idx, perm = _experts(E, K, N, group, dtype, seed=3, act_order=True)
    _, tw, ids = _route(M, E, topk, seed=9)
    a = torch.randn(M, K, generator=torch.Generator().manual_seed(1)).to(dtype)
    sorted_ids, _, _ = fm.moe_align_block_size(ids.to(DEV), block, E)
    ws = torch.zeros(16, dtype=torch.int32, device=DEV)
    out = torch.ops._moe_C.marlin_gemm_moe(a.to(DEV), q.to(DEV), sorted_ids, tw.to(DEV), ids.to(DEV), s.to(DEV),
                                           g_idx.to(DEV), perm.to(DEV), ws, M, N, K, True, E, topk, block, True, False)
    torch.cuda.synchronize()
    _check(out, om.marlin_gemm_moe(a, refs, ids, tw, True, False))


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("cfg", [(1, 256, 128, 8, 2), (64, 512, 256, 8, 2), (200, 256, 512, 16, 4)])
def test_fused_marlin_moe_layer(ops, dtype, cfg):
    """hidden K -> experts' gate/up [K, 2N] -> silu_and_mul -> down [N, K]; Mixtral-style top-k routing."""
    from aphrodite_engine_b200 import fused_moe as fm
    M, K, N, E, topk = cfg
    r1, q1, s1, g1, p1 = _experts(E, K, 2 * N, 128, dtype, seed=11)
    r2, q2, s2, g2, p2 = _experts(E, N, K, 128, dtype, seed=12)
    gate, tw, ids = _route(M, E, topk, seed=M)
    tw = tw / tw.sum(dim=-1, keepdim=True)
    a = torch.randn(M, K, generator=torch.Generator().manual_seed(2)).to(dtype)
    out = fm.fused_marlin_moe(a.to(DEV), q1.to(DEV), q2.to(DEV), gate.to(DEV), g1.to(DEV), g2.to(DEV), p1.to(DEV),
                              p2.to(DEV), topk, renormalize=True, w1_scale=s1.to(DEV), w2_scale=s2.to(DEV))
    torch.cuda.synchronize()
    assert out.shape == (M, K)
    _check(out, om.fused_marlin_moe(a, r1, r2, tw, ids), tol=2e-2)
