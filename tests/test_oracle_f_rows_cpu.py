"""CPU: the restatements of the SURVEY §8(f) rows (oracle/f_rows.py) are pinned before they are trusted as checkers.

* prefill attention: against golden vectors produced by the REFERENCE'S OWN Triton kernel (prefix_prefill.py) run by the
  Triton CPU interpreter (tests/golden/make_golden_prefill.py) — fp16, GQA / MQA / MHA, head 64 / 96 / 128 / 256, block 16 /
  32, sliding window, ALiBi, fp8-e4m3 / e5m2 caches with scales.
* sampling / renorm / mask: against brute-force definitions (sort-based top-k / top-p sets; the rejection samplers'
  outputs must lie inside the filtered set whenever they report success).
* fp8 quantisation and the scaled GEMM: against plain torch arithmetic of the same definition."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import f_rows
from tests import golden_io

GOLDEN = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(golden_io.GOLDEN_DIR, "prefill_*.npz")))


def test_prefill_golden_files_present():
    assert len(GOLDEN) >= 8


@pytest.mark.parametrize("name", GOLDEN)
def test_prefill_oracle_matches_reference_triton_kernel(name):
    d = golden_io.load(name)
    kvd = str(np.load(os.path.join(golden_io.GOLDEN_DIR, name + ".npz"))["kv_cache_dtype"])
    out = f_rows.context_attention(d["q"], d["k"], d["v"], d["key_cache"], d["value_cache"], d["block_tables"],
                                   d["start_loc"], d["seq_lens"], d["ctx_lens"], kvd, float(d["k_scale"]),
                                   float(d["v_scale"]), d.get("alibi_slopes"), int(d["sliding_window"]))
    assert not torch.isnan(d["out"]).any()
    # two fp16 ulps at the outputs' magnitude (<= 2): the reference renormalises P per tile, the restatement once
    torch.testing.assert_close(out.float(), d["out"].float(), atol=1e-3, rtol=1e-3)


def _probs(B, V, seed, peaky=True):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, V, generator=g) * (3.0 if peaky else 0.5)
    return torch.softmax(logits, -1).numpy().astype(np.float32)


def test_sampling_from_probs_is_the_inverse_cdf():
    p = _probs(16, 1000, 0)
    u = np.random.default_rng(0).random(16).astype(np.float32)
    ids, margins = f_rows.sampling_from_probs(p, u)
    for b in range(16):
        cdf = np.cumsum(p[b].astype(np.float64))
        assert cdf[ids[b]] > u[b] and (ids[b] == 0 or cdf[ids[b] - 1] <= u[b])
    assert (margins >= 0).all()
    # u beyond the total mass: the last index
    ids, _ = f_rows.sampling_from_probs(p[:1] * 0.5, np.array([0.9], dtype=np.float32))
    assert ids[0] == 999


@pytest.mark.parametrize("mode", ["top_k", "top_p", "min_p", "top_k_top_p"])
def test_rejection_samplers_land_inside_the_filter(mode):
    B, V, R = 32, 2000, 32
    p = _probs(B, V, 1)
    u = np.random.default_rng(1).random((R, B)).astype(np.float32)
    k, pp = 20, 0.8
    if mode == "min_p":
        pp = 0.1
    ids, ok, _ = f_rows.rejection_sampling(mode, p, u, k=k, p=pp)
    assert ok.all()                      # 32 rounds are plenty for these filters
    for b in range(B):
        row = p[b]
        order = np.sort(row)[::-1]
        if mode in ("top_k", "top_k_top_p"):
            assert row[ids[b]] >= order[k - 1]
        if mode in ("top_p", "top_k_top_p"):
            # the sampled entry belongs to the nucleus: the mass of strictly larger entries is < p
            assert row[row > row[ids[b]]].sum() < pp + 1e-6
        if mode == "min_p":
            assert row[ids[b]] >= row.max() * pp


def test_renorm_and_mask_match_sort_based_definitions():
    B, V = 8, 1500
    p = _probs(B, V, 2)
    k = np.array([1, 2, 5, 50, 100, 1499, 1500, 3000], dtype=np.int32)
    out = f_rows.top_k_renorm_prob(p, k)
    for b in range(B):
        kept = out[b] > 0
        kk = min(int(k[b]), V)
        assert kept.sum() == kk                       # softmax of gaussians: no ties
        assert set(np.nonzero(kept)[0]) == set(np.argsort(-p[b])[:kk])
        np.testing.assert_allclose(out[b].sum(), 1.0 if kk < V else p[b].sum(), rtol=1e-5)
    logits = np.random.default_rng(3).standard_normal((B, V)).astype(np.float32)
    m = f_rows.top_k_mask_logits(logits, k)
    for b in range(B):
        kk = min(int(k[b]), V)
        assert np.isfinite(m[b]).sum() == kk
        assert (m[b][np.isfinite(m[b])] == logits[b][np.isfinite(m[b])]).all()
    tp = np.array([0.1, 0.5, 0.9, 0.99, 1e-6, 0.3, 0.7, 0.95], dtype=np.float32)
    out = f_rows.top_p_renorm_prob(p, tp)
    for b in range(B):
        order = np.argsort(-p[b])
        cs = np.cumsum(p[b][order].astype(np.float64))
        n = int(np.searchsorted(cs, tp[b], side="left")) + 1          # smallest prefix with mass >= p
        assert set(np.nonzero(out[b] > 0)[0]) == set(order[:n])
        np.testing.assert_allclose(out[b].sum(), 1.0, rtol=1e-5)


def test_fp8_quant_restatement():
    x = torch.randn(7, 96) * 3
    x[0, 0] = 1000.0
    s = torch.tensor([0.5])
    q = f_rows.static_scaled_fp8_quant(x, s)
    assert q.dtype == torch.float8_e4m3fn and float(q[0, 0].float()) == 448.0        # saturates, never NaN / inf
    torch.testing.assert_close(q.float(), torch.clamp(x / 0.5, -448, 448).to(torch.float8_e4m3fn).float())
    q, sc = f_rows.dynamic_scaled_fp8_quant(x)
    assert float(sc) == float(np.float32(1000.0) / np.float32(448.0))
    q, sc = f_rows.dynamic_per_token_scaled_fp8_quant(x)
    assert sc.shape == (7, 1) and float(sc[0]) == float(np.float32(1000.0) / np.float32(448.0))
    assert float(q[0, 0].float()) == 448.0
    q, sc = f_rows.dynamic_per_token_scaled_fp8_quant(torch.zeros(2, 8))
    assert float(sc[0]) == pytest.approx(1.0 / (448.0 * 512.0))                      # the minimum scaling factor
    q, sc = f_rows.dynamic_per_token_scaled_fp8_quant(x, torch.tensor([2.0]))
    assert float(sc[0]) == pytest.approx(2.0 / 448.0)


def test_scaled_mm_restatement():
    g = torch.Generator().manual_seed(0)
    a = (torch.randn(5, 64, generator=g)).to(torch.float8_e4m3fn)
    b = (torch.randn(64, 32, generator=g)).to(torch.float8_e4m3fn)
    sa, sb = torch.rand(5, generator=g) + 0.5, torch.rand(32, generator=g) + 0.5
    bias = torch.randn(32, generator=g).to(torch.bfloat16)
    out = f_rows.scaled_mm(a, b, sa, sb, torch.bfloat16, bias)
    ref = (sa[:, None] * a.float()) @ (b.float() * sb[None, :]) + bias.float()      # the reference test's baseline_scaled_mm
    torch.testing.assert_close(out.float(), ref, atol=2e-2, rtol=1e-2)
    ai = torch.randint(-128, 128, (5, 64), generator=g, dtype=torch.int8)
    bi = torch.randint(-128, 128, (64, 32), generator=g, dtype=torch.int8)
    out = f_rows.scaled_mm(ai, bi, torch.tensor([0.01]), torch.tensor([0.02]), torch.float16)
    torch.testing.assert_close(out.float(), (ai.float() @ bi.float()) * 0.0002, atol=1e-2, rtol=1e-3)
