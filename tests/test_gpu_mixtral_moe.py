"""The quantised Mixtral MoE block (aphrodite_engine_b200/mixtral_moe.py) against
  * the CPU oracle restating aphrodite/modeling/models/mixtral_quant.py:128-152 on dequantised expert weights,
  * itself with the reference's element-wise expert loop instead of the fused scale/accumulate kernel (bit-exact),
  * itself over the reference's own CUDA kernels (oracle/_ref/_ref_cuda_C.so), when that library is present."""
import pytest
import torch

from oracle import marlin as om

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _build(dtype, tp_size, tp_rank, T=48, H=512, I=1024, E=4, topk=2, seed=0):
    from aphrodite_engine_b200.mixtral_moe import MixtralQuantMoE, MixtralShape
    shape = MixtralShape(name="tiny", hidden=H, intermediate=I, num_experts=E, topk=topk, group_size=128)
    moe = MixtralQuantMoE(shape, DEV, dtype, tp_rank=tp_rank, tp_size=tp_size)
    g = torch.Generator().manual_seed(seed)
    w13_refs, w2_refs = {}, {}
    for e in moe.experts:                       # replace the random Marlin words by quantised REAL weights
        for name, (k, n), refs, store in (("w13", (H, 2 * I), w13_refs, moe.w13), ("w2", (I, H), w2_refs, moe.w2)):
            w = (torch.randn(k, n, generator=g) * 0.05).to(dtype)
            w_ref, mq, ms, mz = om.awq_marlin_quantize(w, 4, 128)
            refs[e] = w_ref
            store[e].update(q=mq.to(DEV), s=ms.to(DEV), z=mz.to(DEV))
    x = (torch.randn(T, H, generator=g) * 0.5).to(dtype)
    return moe, x, w13_refs, w2_refs


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("tp", [(1, 0), (2, 1), (4, 3)])
def test_mixtral_quant_moe_matches_oracle_and_reference_loop(dtype, tp):
    from aphrodite_engine_b200.mixtral_moe import MixtralQuantMoE
    tp_size, tp_rank = tp
    moe, x, w13_refs, w2_refs = _build(dtype, tp_size, tp_rank)
    out = moe.forward(x.to(DEV))
    torch.cuda.synchronize()
    want = om.mixtral_quant_moe(x, moe.gate.cpu(), w13_refs, w2_refs, moe.s.topk, moe.experts)
    err = (out.float().cpu() - want.float()).abs().max() / want.float().abs().max().clamp_min(1e-6)
    assert float(err) < 2e-2, float(err)
    # the fused mask/scale/accumulate kernel reproduces the reference's element-wise loop bit for bit
    loop = MixtralQuantMoE(moe.s, DEV, dtype, tp_rank=tp_rank, tp_size=tp_size, fused_scale_add=False, share_from=moe)
    assert torch.equal(loop.forward(x.to(DEV)), out)


def test_mixtral_quant_moe_vs_reference_cuda_kernels():
    from oracle import ref_cuda_ops as rco
    if not rco.available():
        pytest.skip("oracle/_ref/_ref_cuda_C.so not built")
    from aphrodite_engine_b200.mixtral_moe import MixtralQuantMoE
    moe, x, _, _ = _build(torch.bfloat16, 4, 1, T=128)
    ref = MixtralQuantMoE(moe.s, DEV, torch.bfloat16, tp_rank=1, tp_size=4, op_table=rco.RefCudaOps(), share_from=moe)
    a, b = moe.forward(x.to(DEV)).float(), ref.forward(x.to(DEV)).float()
    torch.cuda.synchronize()
    assert float((a - b).abs().max() / b.abs().max().clamp_min(1e-6)) < 2e-2
