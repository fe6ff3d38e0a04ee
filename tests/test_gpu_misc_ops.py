"""GPU parity: permute_cols, awq_dequantize, advance_step_flashattn and the act-order Marlin GEMM path."""
import pytest
import torch

from oracle import marlin as om

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(1, 128), (33, 512), (256, 4096)])
def test_permute_cols_exact(ops, dtype, shape):
    M, K = shape
    g = torch.Generator().manual_seed(M + K)
    a = torch.randn(M, K, generator=g).to(dtype)
    perm = torch.randperm(K, generator=g).int()
    out = ops.permute_cols(a.to(DEV), perm.to(DEV))
    assert torch.equal(out.cpu(), a[:, perm.long()])


@pytest.mark.parametrize("shape", [(128, 64, 128), (512, 256, 128), (256, 1024, 32)])
def test_awq_dequantize_exact(ops, shape):
    K, N, G = shape
    g = torch.Generator().manual_seed(K + N)
    q_w = torch.randint(0, 16, (K, N), generator=g, dtype=torch.int32)
    zp = torch.randint(0, 16, (K // G, N), generator=g, dtype=torch.int32)
    s = (torch.rand(K // G, N, generator=g) * 0.02 + 0.001).to(torch.float16)
    qp, zpp = om.awq_pack(q_w, 4), om.awq_pack(zp, 4)
    out = ops.awq_dequantize(qp.to(DEV), s.to(DEV), zpp.to(DEV), 0, 0, 0)
    ref = om.awq_dequantize(qp, s, zpp)
    assert out.dtype == torch.float16 and tuple(out.shape) == (K, N)
    assert torch.equal(out.cpu(), ref)
    direct = ((q_w - zp.repeat_interleave(G, 0)).to(torch.float16).float() * s.repeat_interleave(G, 0).float()).half()
    assert torch.equal(ref, direct)


@pytest.mark.parametrize("cfg", [(1, 16), (37, 16), (256, 32), (1000, 8)])
def test_advance_step_flashattn_exact(ops, cfg):
    n, bs = cfg
    g = torch.Generator().manual_seed(n)
    max_blocks = 40
    seq_lens = torch.randint(1, max_blocks * bs - 1, (n,), generator=g, dtype=torch.int32)
    bt = torch.randint(0, 10000, (n, max_blocks), generator=g, dtype=torch.int32)
    sampled = torch.randint(0, 32000, (n, 1), generator=g, dtype=torch.long)
    tokens = torch.zeros(n, dtype=torch.long)
    pos = torch.zeros(n, dtype=torch.long)
    slots = torch.zeros(n, dtype=torch.long)
    d = [t.to(DEV) for t in (tokens, sampled, pos, seq_lens, slots, bt)]
    ops.advance_step_flashattn(n, n, bs, d[0], d[1], d[2], d[3], d[4], d[5])
    torch.cuda.synchronize()
    nl = seq_lens + 1
    np_ = (nl - 1).long()
    exp_slots = bt[torch.arange(n), (np_ // bs)].long() * bs + np_ % bs
    assert torch.equal(d[0].cpu(), sampled.flatten()) and torch.equal(d[3].cpu(), nl)
    assert torch.equal(d[2].cpu(), np_) and torch.equal(d[4].cpu(), exp_slots)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("mkn", [(16, 256, 128), (100, 1024, 256), (256, 2048, 512)])
def test_gptq_marlin_gemm_act_order(ops, dtype, mkn):
    from aphrodite_engine_b200.scalar_type import scalar_types
    M, K, N = mkn
    torch.manual_seed(M + K)
    a = torch.randn(M, K).to(dtype)
    w = torch.randn(K, N).to(dtype)
    w_ref, mq, ms, g_idx, sort_idx = om.marlin_quantize_act_order(w, 4, 128, seed=K)
    # the repack op with the sort permutation reproduces the sorted Marlin layout from the checkpoint layout
    _, q_w, _, _ = om.quantize_weights(w, 4, 128, 8)
    ws = torch.zeros((N // 64) * 16, dtype=torch.int32, device=DEV)
    empty = torch.empty(0, dtype=torch.int32, device=DEV)
    out = ops.gptq_marlin_gemm(a.to(DEV), mq.to(DEV), ms.to(DEV), empty, g_idx.to(DEV), sort_idx.to(DEV), ws,
                               scalar_types.uint4b8, M, N, K, True, False, True, False)
    torch.cuda.synchronize()
    ref = om.marlin_gemm(a, w_ref).float()
    o = out.float().cpu()
    assert ((o - ref).abs().max() / ref.abs().max()) < 1e-2
    assert ((o - ref).abs().mean() / ref.abs().mean()) < 2e-3
