"""GPU parity: rms_norm / fused_add_rms_norm / rotary_embedding / activations vs the CPU oracle."""
import pytest
import torch

from oracle import paged_ops as po
from tests import tolerances as tol

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DTYPES = [torch.bfloat16, torch.float16, torch.float32]


def _name(dt):
    return str(dt).split(".")[-1]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(7, 768), (83, 4096), (256, 4096), (3, 5120), (2, 8192), (5, 8200), (4, 770), (2, 16384)])
@pytest.mark.parametrize("fused", [False, True])
def test_rms_norm(ops, dtype, shape, fused):
    T, Hd = shape
    torch.manual_seed(0)
    x = torch.randn(T, Hd).to(dtype)
    r = torch.randn(T, Hd).to(dtype)
    w = torch.empty(Hd).normal_(1.0, 0.1).to(dtype)
    eps = 1e-6
    xd, rd, wd = x.to(DEV), r.to(DEV), w.to(DEV)
    if fused:
        ops.fused_add_rms_norm(xd, rd, wd, eps)
        ref_x, ref_r = po.fused_add_rms_norm(x, r, w, eps)
        torch.cuda.synchronize()
        assert torch.equal(rd.cpu(), ref_r)           # the residual is a single rounded add: exact
        got = xd.cpu()
    else:
        out = torch.empty_like(xd)
        ops.rms_norm(out, xd, wd, eps)
        ref_x = po.rms_norm(x, w, eps)
        torch.cuda.synchronize()
        got = out.cpu()
    torch.testing.assert_close(got.float(), ref_x.float(), atol=tol.NORM_ATOL, rtol=tol.NORM_RTOL)
    # and much tighter than the reference's own bound: two roundings (T(x*s), then *w) whose first
    # one may flip by an ulp when the fp32 reduction order differs -> at most ~2 ulp of the output
    ulp = {torch.bfloat16: 2 ** -7, torch.float16: 2 ** -10, torch.float32: 1e-5}[dtype]
    assert ((got.float() - ref_x.float()).abs() <= 2.5 * ulp * ref_x.float().abs() + 1e-6).all()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("is_neox", [True, False])
@pytest.mark.parametrize("cfg", [(32, 8, 128, 128), (32, 8, 128, 64), (12, 12, 64, 64), (7, 1, 80, 32), (16, 4, 96, 20)])
@pytest.mark.parametrize("strided", [False, True])
def test_rotary_embedding(ops, dtype, is_neox, cfg, strided):
    Hq, Hkv, D, rot = cfg
    T, max_pos = 37, 4096
    torch.manual_seed(0)
    inv = 1.0 / (10000 ** (torch.arange(0, rot, 2).float() / rot))
    fr = torch.einsum("i,j->ij", torch.arange(max_pos).float(), inv)
    cache = torch.cat((fr.cos(), fr.sin()), dim=-1).to(dtype)
    pos = torch.randint(0, max_pos, (T,))
    if strided:
        qkv = torch.randn(T, (Hq + 2 * Hkv) * D).to(dtype)
        q, k = qkv[:, : Hq * D], qkv[:, Hq * D: (Hq + Hkv) * D]
        qkvd = qkv.to(DEV)
        qd, kd = qkvd[:, : Hq * D], qkvd[:, Hq * D: (Hq + Hkv) * D]
    else:
        q, k = torch.randn(T, Hq * D).to(dtype), torch.randn(T, Hkv * D).to(dtype)
        qd, kd = q.to(DEV), k.to(DEV)
    rq, rk = po.rotary_embedding(pos, q, k, D, cache, is_neox)
    ops.rotary_embedding(pos.to(DEV), qd, kd, D, cache.to(DEV), is_neox)
    torch.cuda.synchronize()
    n = _name(dtype)
    torch.testing.assert_close(qd.cpu().float(), rq.float(), atol=tol.DEFAULT_ATOL[n], rtol=tol.DEFAULT_RTOL[n])
    torch.testing.assert_close(kd.cpu().float(), rk.float(), atol=tol.DEFAULT_ATOL[n], rtol=tol.DEFAULT_RTOL[n])
    if dtype != torch.float32:   # every step is a single rounded op in 16-bit: bit-exact
        assert torch.equal(qd.cpu(), rq) and torch.equal(kd.cpu(), rk)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_batched_rotary_embedding(ops, dtype):
    Hq, Hkv, D, rot, T, max_pos = 8, 2, 128, 128, 19, 512
    torch.manual_seed(0)
    inv = 1.0 / (10000 ** (torch.arange(0, rot, 2).float() / rot))
    fr = torch.einsum("i,j->ij", torch.arange(2 * max_pos).float(), inv)
    cache = torch.cat((fr.cos(), fr.sin()), dim=-1).to(dtype)
    pos = torch.randint(0, max_pos, (T,))
    offs = torch.randint(0, 2, (T,)) * max_pos
    q, k = torch.randn(T, Hq * D).to(dtype), torch.randn(T, Hkv * D).to(dtype)
    rq, rk = po.rotary_embedding(pos, q, k, D, cache, True, rot, offs)
    qd, kd = q.to(DEV), k.to(DEV)
    ops.batched_rotary_embedding(pos.to(DEV), qd, kd, D, cache.to(DEV), True, rot, offs.to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(qd.cpu(), rq) and torch.equal(kd.cpu(), rk)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(7, 512), (83, 14336), (256, 14336), (5, 13824), (3, 1001)])
@pytest.mark.parametrize("name", ["silu_and_mul", "gelu_and_mul", "gelu_tanh_and_mul"])
def test_act_and_mul(ops, dtype, shape, name):
    T, d = shape
    torch.manual_seed(0)
    x = torch.randn(T, 2 * d).to(dtype)
    out = torch.empty(T, d, dtype=dtype, device=DEV)
    getattr(ops, name)(out, x.to(DEV))
    torch.cuda.synchronize()
    ref = getattr(po, name)(x)
    n = _name(dtype)
    if name == "silu_and_mul" and dtype != torch.float32:
        # the reference test demands exact equality (tests/kernels/test_activation.py:55-58);
        # allow only the 1-ulp of expf libm-vs-device difference
        diff = (out.cpu().float() - ref.float()).abs()
        assert (diff <= 2 ** -7 * ref.float().abs() + 1e-30).all()
        assert (out.cpu() == ref).float().mean() > 0.999
    torch.testing.assert_close(out.cpu().float(), ref.float(), atol=tol.DEFAULT_ATOL[n], rtol=tol.DEFAULT_RTOL[n])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["gelu_new", "gelu_fast", "gelu_quick"])
def test_activation(ops, dtype, name):
    torch.manual_seed(0)
    x = torch.randn(33, 1000).to(dtype)
    out = torch.empty_like(x, device=DEV)
    getattr(ops, name)(out, x.to(DEV))
    torch.cuda.synchronize()
    ref = getattr(po, name)(x)
    n = _name(dtype)
    torch.testing.assert_close(out.cpu().float(), ref.float(), atol=max(tol.DEFAULT_ATOL[n], 4e-3 if dtype != torch.float32 else 0),
                               rtol=tol.DEFAULT_RTOL[n])
