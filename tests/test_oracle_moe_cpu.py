"""CPU checks of the MoE / HQQ parts of the oracle (oracle/marlin.py) against independent brute-force restatements
of the reference semantics: kernels/moe/marlin_moe_ops.cu:1482-1546 (row r = t*topk + j of C comes from token t and
expert topk_ids[t, j]), tests/kernels/test_moe.py:15-29 (`torch_moe`: layer = sum_j w_j * down_e(silu(gate) * up)),
aphrodite/quantization/hqq_marlin.py:183-226 (W = (q - zero) * scale, zeros permuted like the scales)."""
import pytest
import torch

from oracle import marlin as om
from oracle import paged_ops as po


def _route(M, E, topk, seed):
    g = torch.Generator().manual_seed(seed)
    w, ids, _ = po.topk_softmax(torch.randn(M, E, generator=g), topk)
    return w, ids.int()


@pytest.mark.parametrize("replicate,apply_w", [(True, False), (True, True), (False, True), (False, False)])
def test_marlin_gemm_moe_oracle_matches_row_by_row_loop(replicate, apply_w):
    M, K, N, E, topk = 13, 32, 24, 5, 2
    g = torch.Generator().manual_seed(3)
    w_refs = [torch.randn(K, N, generator=g).half() for _ in range(E)]
    tw, ids = _route(M, E, topk, 1)
    a = torch.randn(M if replicate else M * topk, K, generator=g).half()
    got = om.marlin_gemm_moe(a, w_refs, ids, tw, replicate, apply_w)
    assert got.shape == (M, topk, N) and got.dtype == a.dtype
    for t in range(M):
        for j in range(topk):
            r = t * topk + j
            src = a[t] if replicate else a[r]
            o = (src.float() @ w_refs[int(ids[t, j])].float()).half()
            if apply_w:
                o = (tw[t, j].float() * o.float()).half()
            # same arithmetic, different fp32 summation order (batched vs per-row matmul): <= 1 fp16 ulp
            torch.testing.assert_close(got[t, j].float(), o.float(), atol=8e-3, rtol=2e-3)


def test_marlin_gemm_moe_oracle_leaves_unrouted_rows_zero():
    M, K, N, E, topk = 4, 16, 8, 3, 2
    w_refs = [torch.ones(K, N).half() for _ in range(E)]
    ids = torch.tensor([[0, 1], [2, 7], [1, 0], [9, 9]], dtype=torch.int32)      # 7 and 9 are not experts
    tw = torch.ones(M, topk)
    a = torch.ones(M, K).half()
    got = om.marlin_gemm_moe(a, w_refs, ids, tw, True, False)
    assert torch.equal(got[1, 1], torch.zeros(N).half()) and torch.equal(got[3], torch.zeros(topk, N).half())
    assert torch.equal(got[0, 0], torch.full((N,), float(K)).half())


def test_fused_marlin_moe_oracle_matches_torch_moe_restatement():
    """tests/kernels/test_moe.py:15-29, with dequantised [K, N]-layout expert weights and fp32 accumulation."""
    M, K, N, E, topk = 9, 48, 40, 4, 2
    g = torch.Generator().manual_seed(5)
    w1 = [torch.randn(K, 2 * N, generator=g).half() * 0.2 for _ in range(E)]
    w2 = [torch.randn(N, K, generator=g).half() * 0.2 for _ in range(E)]
    tw, ids = _route(M, E, topk, 2)
    tw = tw / tw.sum(dim=-1, keepdim=True)
    a = torch.randn(M, K, generator=g).half()
    got = om.fused_marlin_moe(a, w1, w2, tw, ids)
    ref = torch.zeros(M, K)
    for t in range(M):
        for j in range(topk):
            e = int(ids[t, j])
            h = a[t].float() @ w1[e].float()
            act = torch.nn.functional.silu(h[:N]) * h[N:]
            ref[t] += tw[t, j] * (act @ w2[e].float())
    # the oracle rounds to fp16 where the kernels do (after each GEMM, after silu, after the weighting): small drift
    torch.testing.assert_close(got.float(), ref, atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("group_size", [64, 128])
def test_hqq_marlin_quantize_layout_and_reference_weights(group_size):
    K, N = 256, 128
    g = torch.Generator().manual_seed(group_size)
    q = torch.randint(0, 16, (K, N), generator=g)
    s = (torch.rand(K // group_size, N, generator=g) * 0.02 + 0.005).half()
    zp = (torch.rand(K // group_size, N, generator=g) * 4 + 6).half()
    w_ref, mq, ms, mz = om.hqq_marlin_quantize(q, s, zp, group_size)
    # (q - zero) * scale with both operations rounded to fp16 — the reference kernel's sub_zpf + scale chain
    exp = torch.empty(K, N, dtype=torch.float16)
    for k in range(K):
        gi = k // group_size
        exp[k] = (q[k].half() - zp[gi]) * s[gi]
    assert torch.equal(w_ref, exp)
    # zero points travel in the scales' Marlin permutation; codes are plain uint4 Marlin tiles
    assert torch.equal(mz, om.marlin_permute_scales(zp, K, N, group_size))
    assert torch.equal(ms, om.marlin_permute_scales(s, K, N, group_size))
    assert torch.equal(mq, om.marlin_weights(q.int(), 4))


def test_mixtral_quant_dense_per_expert_equals_per_token_routing():
    """oracle.marlin.mixtral_quant_moe (the reference's dense-per-expert loop, mixtral_quant.py:128-152) against the
    per-token definition of a top-k MoE in fp32: out[t] = sum_{k: expert(t,k) local} w[t,k] * MLP_e(x[t])."""
    import torch
    from oracle import marlin as om
    torch.manual_seed(0)
    T, H, I, E, topk = 9, 32, 48, 4, 2
    x = torch.randn(T, H)
    gate = torch.randn(E, H)
    w13 = {e: torch.randn(H, 2 * I) * 0.2 for e in range(E)}
    w2 = {e: torch.randn(I, H) * 0.2 for e in range(E)}
    p = torch.softmax(x @ gate.t(), dim=1)
    w, ids = torch.topk(p, topk, dim=-1)
    w = w / w.sum(-1, keepdim=True)
    for local in ([0, 1, 2, 3], [2, 3], [1]):
        got = om.mixtral_quant_moe(x, gate, w13, w2, topk, local)
        want = torch.zeros(T, H)
        for t in range(T):
            for k in range(topk):
                e = int(ids[t, k])
                if e in local:
                    gu = x[t] @ w13[e]
                    h = torch.nn.functional.silu(gu[:I]) * gu[I:]
                    want[t] += w[t, k] * (h @ w2[e])
        torch.testing.assert_close(got, want, atol=1e-4, rtol=1e-4)
