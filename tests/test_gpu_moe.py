"""GPU parity: MoE routing kernels vs the CPU oracle (integer outputs bit-exact)."""
import pytest
import torch

from oracle import paged_ops as po

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("id_dtype", [torch.int32, torch.int64])
@pytest.mark.parametrize("cfg", [(1, 2, 8, 16), (33, 2, 8, 16), (128, 2, 8, 64), (255, 4, 60, 32), (1024, 8, 64, 128),
                                 (7, 1, 3, 4), (4096, 2, 16, 16), (50, 6, 160, 8)])
def test_moe_align_block_size_exact(ops, id_dtype, cfg):
    T, topk, E, bs = cfg
    g = torch.Generator().manual_seed(T + E)
    ids = torch.randint(0, E, (T, topk), generator=g).to(id_dtype)
    if T > 8:
        ids[: T // 2] = ids[: T // 2] % max(1, E // 4)        # skewed routing, some experts empty
    numel = ids.numel()
    max_pad = numel + E * (bs - 1)
    sorted_ids = torch.full((max_pad,), numel, dtype=torch.int32, device=DEV)
    expert_ids = torch.full(((max_pad + bs - 1) // bs,), -1, dtype=torch.int32, device=DEV)
    post = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.moe_align_block_size(ids.to(DEV), E, bs, sorted_ids, expert_ids, post)
    torch.cuda.synchronize()
    rs, re, rp = po.moe_align_block_size(ids, E, bs, max_pad)
    assert torch.equal(post.cpu(), rp)
    assert torch.equal(sorted_ids.cpu(), rs)
    assert torch.equal(expert_ids.cpu(), re)


@pytest.mark.parametrize("cfg", [(1, 8, 2), (77, 8, 2), (256, 64, 6), (128, 60, 4), (33, 160, 8), (5, 3, 3), (512, 256, 8)])
def test_topk_softmax(ops, cfg):
    T, E, k = cfg
    torch.manual_seed(T * E)
    gate = torch.randn(T, E, dtype=torch.float32) * 2
    gate[0, :2] = gate[0, :2].max()                             # an exact tie: lowest index must win
    w = torch.empty(T, k, dtype=torch.float32, device=DEV)
    ids = torch.empty(T, k, dtype=torch.int32, device=DEV)
    src = torch.empty(T, k, dtype=torch.int32, device=DEV)
    ops.topk_softmax(w, ids, src, gate.to(DEV))
    torch.cuda.synchronize()
    rw, ri, rsrc = po.topk_softmax(gate, k)
    assert torch.equal(src.cpu(), rsrc)
    torch.testing.assert_close(w.cpu(), rw, atol=1e-6, rtol=1e-5)
    same = ids.cpu() == ri
    if not same.all():   # only near-ties (expf ulp differences) may swap order
        bad = ~same
        assert (w.cpu()[bad] - rw[bad]).abs().max() < 1e-6
