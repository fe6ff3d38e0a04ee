"""GPU parity against the committed golden vectors: outputs of the reference's own CPU kernels
(tests/golden, generated in the build container from /root/reference/kernels/cpu)."""
import pytest
import torch

from tests import tolerances as tol
from tests.golden_io import load

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TAGS = ["f32", "bf16"]


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("alibi", [False, True])
@pytest.mark.parametrize("version", ["v1", "v2"])
def test_attention_vs_reference_cpu_kernel(ops, tag, alibi, version):
    g = load(f"paged_attention_{tag}")
    q, kc, vc = g["q"].to(DEV), g["key_cache"].to(DEV), g["value_cache"].to(DEV)
    bt, sl = g["block_tables"].to(DEV), g["seq_lens"].to(DEV)
    al = g["alibi"].to(DEV) if alibi else None
    max_len = int(g["seq_lens"].max())
    out = torch.empty_like(q)
    if version == "v1":
        ops.paged_attention_v1(out, q, kc, vc, g["num_kv_heads"], g["scale"], bt, sl, g["block_size"],
                               max_len, al, "auto", 1.0, 1.0)
    else:
        S, H, D = q.shape
        P = (max_len + 511) // 512
        tmp = torch.empty(S, H, P, D, dtype=q.dtype, device=DEV)
        es = torch.empty(S, H, P, dtype=torch.float32, device=DEV)
        ml = torch.empty_like(es)
        ops.paged_attention_v2(out, es, ml, tmp, q, kc, vc, g["num_kv_heads"], g["scale"], bt, sl,
                               g["block_size"], max_len, al, "auto", 1.0, 1.0)
    ref = g[f"out_{version}{'_alibi' if alibi else ''}"]
    torch.testing.assert_close(out.cpu().float(), ref.float(), atol=tol.ATTN_ATOL, rtol=tol.ATTN_RTOL)


@pytest.mark.parametrize("tag", TAGS)
def test_cache_ops_vs_reference_cpu_kernel_bit_exact(ops, tag):
    g = load(f"cache_ops_{tag}")
    kc, vc = g["key_cache_in"].to(DEV), g["value_cache_in"].to(DEV)
    ops.reshape_and_cache(g["key"].to(DEV), g["value"].to(DEV), kc, vc, g["slot_mapping"].to(DEV),
                          "auto", 1.0, 1.0)
    assert torch.equal(kc.cpu(), g["key_cache_out"]) and torch.equal(vc.cpu(), g["value_cache_out"])
    ops.copy_blocks([kc], [vc], g["block_mapping"].to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(kc.cpu(), g["key_cache_copied"]) and torch.equal(vc.cpu(), g["value_cache_copied"])


@pytest.mark.parametrize("tag", TAGS)
def test_rms_norm_vs_reference_cpu_kernel(ops, tag):
    g = load(f"rms_norm_{tag}")
    x, r, w = g["x"].to(DEV), g["residual"].to(DEV), g["weight"].to(DEV)
    out = torch.empty_like(x)
    ops.rms_norm(out, x, w, g["eps"])
    torch.testing.assert_close(out.cpu().float(), g["out"].float(), atol=tol.NORM_ATOL, rtol=tol.NORM_RTOL)
    ops.fused_add_rms_norm(x, r, w, g["eps"])
    torch.testing.assert_close(x.cpu().float(), g["fused_out"].float(), atol=tol.NORM_ATOL, rtol=tol.NORM_RTOL)
    torch.testing.assert_close(r.cpu().float(), g["fused_residual"].float(), atol=tol.NORM_ATOL, rtol=tol.NORM_RTOL)


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("style", ["neox", "gptj"])
def test_rotary_vs_reference_cpu_kernel(ops, tag, style):
    g = load(f"rotary_{tag}")
    q, k = g["q"].to(DEV), g["k"].to(DEV)
    ops.rotary_embedding(g["positions"].to(DEV), q, k, g["head_size"], g["cos_sin_cache"].to(DEV),
                         style == "neox")
    atol = 1e-5 if tag == "f32" else 2e-2
    rtol = 1.3e-6 if tag == "f32" else 1.6e-2
    torch.testing.assert_close(q.cpu().float(), g[f"q_{style}"].float(), atol=atol, rtol=rtol)
    torch.testing.assert_close(k.cpu().float(), g[f"k_{style}"].float(), atol=atol, rtol=rtol)


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("name", ["silu_and_mul", "gelu_and_mul", "gelu_tanh_and_mul", "gelu_new",
                                  "gelu_fast", "gelu_quick"])
def test_activations_vs_reference_cpu_kernel(ops, tag, name):
    g = load(f"activations_{tag}")
    x = g["x"].to(DEV)
    out = torch.empty_like(g[name], device=DEV)
    getattr(ops, name)(out, x)
    atol = 2e-5 if tag == "f32" else 2e-2
    rtol = 1.3e-6 if tag == "f32" else 1.6e-2
    torch.testing.assert_close(out.cpu().float(), g[name].float(), atol=atol, rtol=rtol)
