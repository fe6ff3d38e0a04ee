"""GPU parity: cache ops vs the CPU oracle. Index / byte work is compared bit-exactly (torch.equal),
as the reference's tests/kernels/test_cache.py does for the `auto` cache dtype (:105-109, :205-206)."""
import random

import pytest
import torch

from oracle import paged_ops as po
from tests import tolerances as tol

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("shape", [(42, 8, 128, 16), (7, 3, 80, 8), (130, 8, 256, 32), (1, 1, 64, 16)])
@pytest.mark.parametrize("strided", [False, True])
def test_reshape_and_cache_exact(ops, dtype, shape, strided):
    T, H, D, BS = shape
    torch.manual_seed(0)
    random.seed(0)
    NB = (T + BS - 1) // BS + 3
    slots = random.sample(range(NB * BS), T)
    slots[0] = -1  # padding token must be skipped
    slot_mapping = torch.tensor(slots, dtype=torch.long)
    if strided:  # key/value are views into a fused qkv tensor, as in the model (llama.py:209-214)
        qkv = torch.randn(T, 3, H, D).to(dtype)
        key, value = qkv[:, 1], qkv[:, 2]
    else:
        key, value = torch.randn(T, H, D).to(dtype), torch.randn(T, H, D).to(dtype)
    kc, vc = po.make_kv_cache(NB, BS, H, D, dtype, "auto", 1)
    kcd, vcd = kc.to(DEV), vc.to(DEV)
    if strided:
        qkvd = qkv.to(DEV)
        kd, vd = qkvd[:, 1], qkvd[:, 2]
    else:
        kd, vd = key.to(DEV), value.to(DEV)
    ops.reshape_and_cache(kd, vd, kcd, vcd, slot_mapping.to(DEV), "auto", 1.0, 1.0)
    po.reshape_and_cache(key, value, kc, vc, slot_mapping)
    torch.cuda.synchronize()
    assert torch.equal(kcd.cpu(), kc) and torch.equal(vcd.cpu(), vc)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("kv_dtype", ["fp8", "fp8_e4m3", "fp8_e5m2"])
def test_reshape_and_cache_fp8(ops, dtype, kv_dtype):
    T, H, D, BS, NB = 33, 4, 128, 16, 8
    torch.manual_seed(0)
    slot_mapping = torch.randperm(NB * BS)[:T]
    key, value = torch.randn(T, H, D).to(dtype), torch.randn(T, H, D).to(dtype)
    kc = torch.zeros(NB, H, D // 16, BS, 16, dtype=torch.uint8)
    vc = torch.zeros(NB, H, D, BS, dtype=torch.uint8)
    kcd, vcd = kc.to(DEV), vc.to(DEV)
    for ks, vs in ((1.0, 1.0), (0.5, 2.0)):
        ops.reshape_and_cache(key.to(DEV), value.to(DEV), kcd, vcd, slot_mapping.to(DEV), kv_dtype, ks, vs)
        po.reshape_and_cache(key, value, kc, vc, slot_mapping, kv_dtype, ks, vs)
        torch.cuda.synchronize()
        # the quantiser is deterministic RN-even + satfinite on both sides: bytes must agree
        assert torch.equal(kcd.cpu(), kc) and torch.equal(vcd.cpu(), vc)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_reshape_and_cache_flash_exact(ops, dtype):
    T, H, D, BS, NB = 50, 8, 128, 16, 6
    torch.manual_seed(0)
    slot_mapping = torch.randperm(NB * BS)[:T]
    slot_mapping[3] = -1
    key, value = torch.randn(T, H, D).to(dtype), torch.randn(T, H, D).to(dtype)
    kc = torch.randn(NB, BS, H, D).to(dtype)
    vc = torch.randn(NB, BS, H, D).to(dtype)
    kcd, vcd = kc.to(DEV), vc.to(DEV)
    ops.reshape_and_cache_flash(key.to(DEV), value.to(DEV), kcd, vcd, slot_mapping.to(DEV), "auto", 1.0, 1.0)
    po.reshape_and_cache_flash(key, value, kc, vc, slot_mapping)
    torch.cuda.synchronize()
    assert torch.equal(kcd.cpu(), kc) and torch.equal(vcd.cpu(), vc)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32, torch.uint8])
@pytest.mark.parametrize("num_layers", [1, 5])
def test_copy_blocks_exact(ops, dtype, num_layers):
    NB, H, D, BS = 40, 4, 96, 16
    random.seed(0)
    src = random.sample(range(NB), 8)
    rest = list(set(range(NB)) - set(src))
    dst = random.sample(rest, 16)
    mapping = [(s, dst[2 * i + j]) for i, s in enumerate(src) for j in range(2)]
    bm = torch.tensor(mapping, dtype=torch.long)
    kv_dtype = "fp8" if dtype == torch.uint8 else "auto"
    base = torch.float16 if dtype == torch.uint8 else dtype
    kcs, vcs = zip(*[po.make_kv_cache(NB, BS, H, D, base, kv_dtype, seed=i) for i in range(num_layers)])
    kd, vd = [k.to(DEV) for k in kcs], [v.to(DEV) for v in vcs]
    ops.copy_blocks(kd, vd, bm.to(DEV))
    po.copy_blocks(list(kcs), list(vcs), bm)
    torch.cuda.synchronize()
    for a, b in zip(kd + vd, list(kcs) + list(vcs)):
        assert torch.equal(a.cpu(), b)


@pytest.mark.parametrize("direction", ["d2h", "h2d", "d2d"])
def test_swap_blocks_exact(ops, direction):
    NB, H, D, BS = 20, 4, 128, 16
    kc, _ = po.make_kv_cache(NB, BS, H, D, torch.bfloat16, "auto", 0)
    dstc, _ = po.make_kv_cache(NB, BS, H, D, torch.bfloat16, "auto", 1)
    bm = torch.tensor([[0, 5], [3, 1], [19, 19], [7, 0]], dtype=torch.long)
    sdev = "cpu" if direction == "h2d" else DEV
    ddev = "cpu" if direction == "d2h" else DEV
    s, d = kc.to(sdev), dstc.to(ddev)
    if sdev == "cpu":
        s = s.pin_memory()
    ops.swap_blocks(s, d, bm)
    torch.cuda.synchronize()
    po.swap_blocks(kc, dstc, bm)
    assert torch.equal(d.cpu(), dstc)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_convert_fp8_roundtrip(ops, dtype):
    torch.manual_seed(0)
    x = torch.empty(64, 4, 128, 16).uniform_(-240, 240).to(dtype)
    x.view(-1)[:4] = torch.tensor([1e6, -1e6, 0.0, 448.0]).to(dtype)   # saturation / zero / max
    for scale in (1.0, 0.5):
        q = torch.empty(x.shape, dtype=torch.uint8, device=DEV)
        ops.convert_fp8(q, x.to(DEV), scale, "fp8")
        back = torch.empty(x.shape, dtype=dtype, device=DEV)
        ops.convert_fp8(back, q, scale, "fp8")
        torch.cuda.synchronize()
        assert torch.equal(q.cpu(), po.fp8_quant(x, scale))
        assert torch.equal(back.cpu(), po.fp8_dequant(q.cpu(), scale, dtype))
        torch.testing.assert_close(back.cpu().float()[4:], x.float()[4:], atol=tol.CACHE_FP8_ATOL,
                                   rtol=tol.CACHE_FP8_RTOL)


def test_device_attribute_queries(ops):
    smem = ops.get_max_shared_memory_per_block_device_attribute(0)
    assert smem >= 200 * 1024          # B200: 227 KB opt-in
    assert ops.get_device_attribute(16, 0) == torch.cuda.get_device_properties(0).multi_processor_count
