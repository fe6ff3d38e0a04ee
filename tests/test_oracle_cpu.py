"""CPU-only: pins the oracle (oracle/paged_ops.py restatement) against
  (a) the committed golden vectors produced by the reference's own CPU kernels (tests/golden), and
  (b) those kernels live, when oracle/_ref/*.so loads on this host (build container / AVX-512 box).
Index / byte ops are bit-exact; floating-point ops use the reference tests' tolerances."""
import pytest
import torch

from oracle import paged_ops as po
from oracle import ref_lib
from tests import tolerances as tol
from tests.golden_io import load

TAGS = ["f32", "bf16"]
ATOL = {"f32": 1e-5, "bf16": tol.ATTN_ATOL}


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("alibi", [False, True])
def test_attention_restatement_matches_golden(tag, alibi):
    g = load(f"paged_attention_{tag}")
    al = g["alibi"] if alibi else None
    out = po.paged_attention(g["q"], g["key_cache"], g["value_cache"], g["block_tables"],
                             g["seq_lens"], g["scale"], al)
    sfx = "_alibi" if alibi else ""
    for ver in ("v1", "v2"):
        torch.testing.assert_close(out.float(), g[f"out_{ver}{sfx}"].float(), atol=ATOL[tag],
                                   rtol=tol.ATTN_RTOL)


@pytest.mark.parametrize("tag", TAGS)
def test_cache_ops_restatement_matches_golden_bit_exact(tag):
    g = load(f"cache_ops_{tag}")
    kc, vc = g["key_cache_in"].clone(), g["value_cache_in"].clone()
    po.reshape_and_cache(g["key"], g["value"], kc, vc, g["slot_mapping"])
    assert torch.equal(kc, g["key_cache_out"]) and torch.equal(vc, g["value_cache_out"])
    po.copy_blocks([kc], [vc], g["block_mapping"])
    assert torch.equal(kc, g["key_cache_copied"]) and torch.equal(vc, g["value_cache_copied"])


@pytest.mark.parametrize("tag", TAGS)
def test_rms_norm_restatement_matches_golden(tag):
    g = load(f"rms_norm_{tag}")
    torch.testing.assert_close(po.rms_norm(g["x"], g["weight"], g["eps"]).float(), g["out"].float(),
                               atol=tol.NORM_ATOL, rtol=tol.NORM_RTOL)
    fx, fr = po.fused_add_rms_norm(g["x"], g["residual"], g["weight"], g["eps"])
    torch.testing.assert_close(fx.float(), g["fused_out"].float(), atol=tol.NORM_ATOL, rtol=tol.NORM_RTOL)
    torch.testing.assert_close(fr.float(), g["fused_residual"].float(), atol=tol.NORM_ATOL, rtol=tol.NORM_RTOL)


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("style", ["neox", "gptj"])
def test_rotary_restatement_matches_golden(tag, style):
    g = load(f"rotary_{tag}")
    q, k = po.rotary_embedding(g["positions"], g["q"], g["k"], g["head_size"], g["cos_sin_cache"],
                               style == "neox")
    n = "float32" if tag == "f32" else "bfloat16"
    # the CPU reference computes the rotation in fp32 and rounds once; the CUDA reference (restated
    # by the oracle) rounds every product: 16-bit results may differ by an ulp or two.
    atol = 1e-5 if tag == "f32" else 2e-2
    torch.testing.assert_close(q.float(), g[f"q_{style}"].float(), atol=atol, rtol=tol.DEFAULT_RTOL[n])
    torch.testing.assert_close(k.float(), g[f"k_{style}"].float(), atol=atol, rtol=tol.DEFAULT_RTOL[n])


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("name", ["silu_and_mul", "gelu_and_mul", "gelu_tanh_and_mul", "gelu_new",
                                  "gelu_fast", "gelu_quick"])
def test_activation_restatement_matches_golden(tag, name):
    g = load(f"activations_{tag}")
    n = "float32" if tag == "f32" else "bfloat16"
    atol = 2e-5 if tag == "f32" else 2e-2
    torch.testing.assert_close(getattr(po, name)(g["x"]).float(), g[name].float(), atol=atol,
                               rtol=tol.DEFAULT_RTOL[n])


def test_fp8_helpers_roundtrip_and_saturation():
    x = torch.tensor([0.0, 1.0, -1.5, 448.0, 1e6, -1e6, 0.0019], dtype=torch.float32)
    q = po.fp8_quant(x, 1.0)
    back = po.fp8_dequant(q, 1.0, torch.float32)
    assert back[3] == 448.0 and back[4] == 448.0 and back[5] == -448.0      # satfinite
    assert back[0] == 0.0 and back[1] == 1.0 and back[2] == -1.5
    q2 = po.fp8_quant(x, 2.0)
    assert torch.equal(po.fp8_dequant(q2, 2.0, torch.float32)[1:3], torch.tensor([1.0, -1.5]))
    e5 = po.fp8_quant(torch.tensor([1e9]), 1.0, "fp8_e5m2")
    assert po.fp8_dequant(e5, 1.0, torch.float32, "fp8_e5m2")[0] == 57344.0


def test_blocksparse_mask_semantics():
    # local window of 2 blocks + every 3rd block (offset by head) — attention_kernels.cu:210-257
    m = po._blocksparse_mask(seq_len=64 * 6, block_size=16, head=0, kv_head=0, num_heads=4,
                             num_kv_heads=2, tp_rank=0, local_blocks=2, vert_stride=3,
                             bs_block_size=64, head_sliding_step=0)
    kb = (torch.arange(64 * 6) // 64)
    expect = ((kb + 1) % 3 == 0) | (kb > 5 - 2)
    assert torch.equal(m, expect)


needs_ref = pytest.mark.skipif(ref_lib.load() is None, reason="oracle/_ref not loadable on this host")


@needs_ref
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(3, 8, 8, 64), (4, 16, 4, 128), (2, 8, 1, 256), (3, 6, 2, 80)])
def test_attention_restatement_matches_live_reference(dtype, shape):
    S, Hq, Hkv, D = shape
    torch.manual_seed(S * 1000 + D)
    BS, NB = 16, 96
    scale = D ** -0.5
    kc, vc = po.make_kv_cache(NB, BS, Hkv, D, dtype, "auto", D)
    q = torch.empty(S, Hq, D).uniform_(-scale, scale).to(dtype)
    sl = torch.randint(1, 900, (S,), dtype=torch.int32)
    bt = torch.randint(0, NB, (S, (int(sl.max()) + BS - 1) // BS), dtype=torch.int32)
    out = torch.empty_like(q)
    ref_lib.paged_attention_v1(out, q, kc, vc, Hkv, scale, bt, sl, BS, int(sl.max()))
    mine = po.paged_attention(q, kc, vc, bt, sl, scale)
    atol = 1e-5 if dtype == torch.float32 else tol.ATTN_ATOL
    torch.testing.assert_close(mine.float(), out.float(), atol=atol, rtol=tol.ATTN_RTOL)


@needs_ref
def test_reshape_and_cache_matches_live_reference_exact():
    _, rcache, _ = ref_lib.load()
    torch.manual_seed(5)
    T, H, D, BS, NB = 40, 8, 128, 16, 5
    for dtype in (torch.float32, torch.bfloat16):
        key, value = torch.randn(T, H, D).to(dtype), torch.randn(T, H, D).to(dtype)
        kc, vc = po.make_kv_cache(NB, BS, H, D, dtype, "auto", 9)
        kc2, vc2 = kc.clone(), vc.clone()
        slots = torch.randperm(NB * BS)[:T]
        rcache.reshape_and_cache(key, value, kc, vc, slots, "auto", 1.0, 1.0)
        po.reshape_and_cache(key, value, kc2, vc2, slots)
        assert torch.equal(kc, kc2) and torch.equal(vc, vc2)


@needs_ref
def test_baseline_config0_opt125m_fp32_decode_loop_matches_live_reference():
    """BASELINE.json configs[0] — facebook/opt-125m, fp32, CPU backend, bs = 1, seq = 128 — at the attention layer's
    shape (12 heads, MHA, head 64, block 16): the reference's own CPU kernels are driven token by token (cache write,
    then paged attention over the tokens so far), and the restatement must follow every step."""
    _, rcache, _ = ref_lib.load()
    torch.manual_seed(125)
    H, D, BS, SEQ = 12, 64, 16, 128
    NB = SEQ // BS + 2
    scale = D ** -0.5
    kc = torch.zeros(NB, H, D // 4, BS, 4)                  # fp32: x = 16 B / 4 B = 4 elements per run
    vc = torch.zeros(NB, H, D, BS)
    kc2, vc2 = kc.clone(), vc.clone()
    table = torch.randperm(NB)[: SEQ // BS].to(torch.int32).view(1, -1)
    worst = 0.0
    for t in range(SEQ):
        k_new, v_new = torch.randn(1, H, D) * 0.3, torch.randn(1, H, D) * 0.3
        q = torch.empty(1, H, D).uniform_(-scale, scale)
        slot = torch.tensor([int(table[0, t // BS]) * BS + t % BS])
        rcache.reshape_and_cache(k_new, v_new, kc, vc, slot, "auto", 1.0, 1.0)
        po.reshape_and_cache(k_new, v_new, kc2, vc2, slot)
        sl = torch.tensor([t + 1], dtype=torch.int32)
        out = torch.empty_like(q)
        ref_lib.paged_attention_v1(out, q, kc, vc, H, scale, table, sl, BS, t + 1)
        mine = po.paged_attention(q, kc2, vc2, table, sl, scale)
        worst = max(worst, float((mine - out).abs().max()))
    assert torch.equal(kc, kc2) and torch.equal(vc, vc2)
    assert worst <= 1e-5, worst
