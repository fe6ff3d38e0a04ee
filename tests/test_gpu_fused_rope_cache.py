"""The fused rotary + cache-write launch must reproduce `rotary_embedding` followed by `reshape_and_cache` BIT-EXACTLY
(both of which are pinned to the reference's kernels in test_gpu_vs_ref_cuda.py): q, k, key_cache, value_cache."""
import pytest
import torch

from oracle import paged_ops as po

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    import aphrodite_engine_b200._custom_ops as o
    return o


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("kv_dtype", ["auto", "fp8", "fp8_e5m2"])
@pytest.mark.parametrize("cfg", [  # (T, Hq, Hkv, D, BS, neox, rot_dim)
    (256, 32, 8, 128, 16, True, 128),
    (7, 4, 1, 128, 16, True, 128),
    (33, 8, 8, 64, 32, True, 64),
    (5, 8, 2, 128, 16, False, 128),      # GPT-J style: the two-kernel path inside the same call
    (9, 6, 2, 96, 16, True, 48),         # partial rotary: idem
])
def test_fused_rope_cache_equals_the_two_ops(ops, dtype, kv_dtype, cfg):
    from aphrodite_engine_b200 import ext_ops
    T, Hq, Hkv, D, BS, neox, rot = cfg
    if dtype == torch.float32 and kv_dtype != "auto":
        pytest.skip("fp8 cache from fp32 activations is not a reference configuration worth a case")
    torch.manual_seed(T * D + Hq)
    NB = T + 3
    qkv = torch.randn(T, (Hq + 2 * Hkv) * D, device=DEV).to(dtype)
    cos_sin = torch.randn(4096, rot, device=DEV).to(dtype)
    pos = torch.randint(0, 4096, (T,), device=DEV)
    slots = torch.randperm(NB * BS, device=DEV)[:T].long()
    slots[T // 2] = -1                                        # a padding token
    kc0, vc0 = po.make_kv_cache(NB, BS, Hkv, D, dtype, kv_dtype, seed=3)
    k_scale, v_scale = (0.5, 2.0) if kv_dtype != "auto" else (1.0, 1.0)

    def run(fused):
        x = qkv.clone()
        q, k, v = x.split([Hq * D, Hkv * D, Hkv * D], dim=-1)
        kc, vc = kc0.to(DEV).clone(), vc0.to(DEV).clone()
        if fused:
            ext_ops.rotary_embedding_and_cache(pos, q, k, v, D, cos_sin, neox, kc, vc, slots, kv_dtype, k_scale, v_scale)
        else:
            ops.rotary_embedding(pos, q, k, D, cos_sin, neox)
            ops.reshape_and_cache(k.view(T, Hkv, D), v.view(T, Hkv, D), kc, vc, slots, kv_dtype, k_scale, v_scale)
        torch.cuda.synchronize()
        return x, kc, vc

    a, b = run(True), run(False)
    for name, u, w in zip(("qkv", "key_cache", "value_cache"), a, b):
        assert torch.equal(u.view(torch.uint8) if u.dtype != torch.uint8 else u,
                           w.view(torch.uint8) if w.dtype != torch.uint8 else w), name
