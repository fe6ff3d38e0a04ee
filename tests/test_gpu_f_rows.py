"""GPU parity of the SURVEY §8(f) rows — every call goes torch op -> C ABI -> sm_100a kernel.

  f1 prefill attention over the paged cache : golden vectors from the reference's own Triton kernel + the oracle on
                                              shapes the golden set does not hold (bf16, blocks 8 / 32 / 64, x = 16,
                                              head 80 / 112 / 192, long contexts, strided q / k / v views)
  f2 sampling                               : the oracle (exact ids wherever the decision margin exceeds fp32 rounding)
                                              and the reference's own CUDA kernels recompiled for sm_100a
  f4 fp8 quantisation / W8A8 scaled GEMM    : bit-exact against the oracle and the reference's kernels (quant);
                                              the GEMM against the exact-sum restatement, every k-split bit-identical
"""
import glob
import math
import os
import random

import numpy as np
import pytest
import torch

from oracle import f_rows
from oracle import paged_ops as po
from tests import golden_io

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "_ref_cuda_C.so")
GOLDEN = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(golden_io.GOLDEN_DIR, "prefill_*.npz")))


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/_ref_cuda_C.so not built (needs /root/reference at build time)")
    torch.ops.load_library(REF_SO)
    r = torch.ops._ref_cuda_C
    if not hasattr(r, "sampling_from_probs"):
        pytest.skip("oracle/_ref/_ref_cuda_C.so predates the sampling / fp8-quant additions")
    return r


# ====================================================================================================== f1 prefill
def _run_prefill(d, kvd, alibi=None, sw=0):
    from aphrodite_engine_b200.attention.prefix_prefill import context_attention_fwd
    q, k, v = d["q"].to(DEV), d["k"].to(DEV), d["v"].to(DEV)
    o = torch.full_like(q, float("nan"))
    context_attention_fwd(q, k, v, o, kvd, d["key_cache"].to(DEV), d["value_cache"].to(DEV), d["block_tables"].to(DEV),
                          d["start_loc"].to(DEV), d["seq_lens"].to(DEV), d["ctx_lens"].to(DEV), int(d["max_query_len"]),
                          float(d["k_scale"]), float(d["v_scale"]), None if alibi is None else alibi.to(DEV), sw)
    torch.cuda.synchronize()
    return o.cpu()


@pytest.mark.parametrize("name", GOLDEN)
def test_prefill_vs_reference_triton_golden(name):
    d = golden_io.load(name)
    kvd = str(np.load(os.path.join(golden_io.GOLDEN_DIR, name + ".npz"))["kv_cache_dtype"])
    out = _run_prefill(d, kvd, d.get("alibi_slopes"), int(d["sliding_window"]))
    assert not torch.isnan(out).any()
    # the reference's own test holds its kernel to atol 1e-3 against xformers for fp8 caches (test_prefix_prefill.py:221)
    torch.testing.assert_close(out.float(), d["out"].float(), atol=2e-3, rtol=2e-3)


def _make_prefill_case(dt, Hq, Hkv, D, BS, ctxs, qls, kvd, x, seed, scale_kv=(1.0, 1.0)):
    g = torch.Generator().manual_seed(seed)
    B, T = len(ctxs), sum(qls)
    q = (torch.randn(T, Hq, D, generator=g) * 0.5).to(dt)
    k = (torch.randn(T, Hkv, D, generator=g) * 0.5).to(dt)
    v = (torch.randn(T, Hkv, D, generator=g) * 0.5).to(dt)
    max_blocks = max((c + BS - 1) // BS for c in ctxs) + 1
    NB = B * max_blocks + 2
    bt = torch.randperm(NB, generator=g)[: B * max_blocks].view(B, max_blocks).to(torch.int32)
    kc = torch.randn(NB, Hkv, D // x, BS, x, generator=g) * 0.5
    vc = torch.randn(NB, Hkv, D, BS, generator=g) * 0.5
    if kvd == "auto":
        kc, vc = kc.to(dt), vc.to(dt)
    else:
        f8 = torch.float8_e4m3fn if kvd in ("fp8", "fp8_e4m3") else torch.float8_e5m2
        kc = (kc / scale_kv[0]).to(f8).view(torch.uint8)
        vc = (vc / scale_kv[1]).to(f8).view(torch.uint8)
    start = torch.tensor([sum(qls[:i]) for i in range(B)], dtype=torch.int32)
    return dict(q=q, k=k, v=v, key_cache=kc, value_cache=vc, block_tables=bt, start_loc=start,
                seq_lens=torch.tensor([c + n for c, n in zip(ctxs, qls)], dtype=torch.int32),
                ctx_lens=torch.tensor(ctxs, dtype=torch.int32), max_query_len=max(qls), k_scale=scale_kv[0],
                v_scale=scale_kv[1])


@pytest.mark.parametrize("cfg", [
    # dtype, Hq, Hkv, D, BS, ctx lens, query lens, kv dtype, x, sliding window, alibi
    (torch.bfloat16, 8, 2, 128, 16, [300, 0, 65, 1], [130, 64, 1, 77], "auto", 8, 0, False),
    (torch.bfloat16, 4, 4, 64, 8, [23, 64], [70, 9], "auto", 8, 0, False),
    (torch.float16, 6, 2, 80, 32, [100, 31], [33, 129], "auto", 8, 0, False),
    (torch.bfloat16, 4, 1, 112, 64, [200, 64], [64, 65], "auto", 8, 0, True),
    (torch.float16, 4, 2, 192, 16, [90], [100], "auto", 8, 32, False),
    (torch.bfloat16, 8, 2, 128, 16, [250, 17], [40, 200], "fp8", 16, 0, False),
    (torch.float16, 4, 4, 128, 32, [64, 128], [128, 3], "fp8_e5m2", 16, 128, False),
    (torch.bfloat16, 2, 1, 256, 16, [513], [190], "auto", 8, 0, False),
    (torch.bfloat16, 32, 8, 128, 16, [1024, 700], [256, 320], "auto", 8, 0, False),
])
def test_prefill_vs_oracle(cfg):
    dt, Hq, Hkv, D, BS, ctxs, qls, kvd, x, sw, use_alibi = cfg
    scales = (1.0, 1.0) if kvd == "auto" else (0.75, 1.5)
    d = _make_prefill_case(dt, Hq, Hkv, D, BS, ctxs, qls, kvd, x, seed=D + BS + len(ctxs), scale_kv=scales)
    alibi = (torch.rand(Hq) * 0.2 + 0.01).float() if use_alibi else None
    out = _run_prefill(d, kvd, alibi, sw)
    ref = f_rows.context_attention(d["q"], d["k"], d["v"], d["key_cache"], d["value_cache"], d["block_tables"],
                                   d["start_loc"], d["seq_lens"], d["ctx_lens"], kvd, scales[0], scales[1], alibi, sw)
    assert not torch.isnan(out).any()
    tol = 2e-3 if dt == torch.float16 else 1e-2          # one bf16 ulp at |out| ~ 1 is 8e-3
    torch.testing.assert_close(out.float(), ref.float(), atol=tol, rtol=tol)


def test_prefill_strided_qkv_views_and_forward_prefix():
    """q / k / v as views of one fused qkv buffer (the caller's layout) through PagedAttention.forward_prefix; garbage
    (NaN) in the cache slots beyond the context must not leak into the output."""
    from aphrodite_engine_b200.attention.paged_attn import PagedAttention
    dt, Hq, Hkv, D, BS = torch.bfloat16, 8, 2, 128, 16
    d = _make_prefill_case(dt, Hq, Hkv, D, BS, [37, 90], [50, 20], "auto", 8, seed=5)
    T = d["q"].shape[0]
    qkv = torch.zeros(T, (Hq + 2 * Hkv) * D, dtype=dt)
    qkv[:, : Hq * D] = d["q"].reshape(T, -1)
    qkv[:, Hq * D: (Hq + Hkv) * D] = d["k"].reshape(T, -1)
    qkv[:, (Hq + Hkv) * D:] = d["v"].reshape(T, -1)
    qkv = qkv.to(DEV)
    q, k, v = qkv.split([Hq * D, Hkv * D, Hkv * D], dim=-1)
    kc, vc = d["key_cache"].clone(), d["value_cache"].clone()
    for b, ctx in enumerate([37, 90]):                    # poison every slot past the context in the last used block
        blk, off = int(d["block_tables"][b, ctx // BS]), ctx % BS
        if off:
            kc[blk, :, :, off:, :] = float("nan")
            vc[blk, :, :, off:] = float("nan")
    qsl = torch.tensor([0, 50, 70], dtype=torch.int32, device=DEV)
    out = PagedAttention.forward_prefix(q.view(T, Hq, D), k.view(T, Hkv, D), v.view(T, Hkv, D), "auto", kc.to(DEV),
                                        vc.to(DEV), d["block_tables"].to(DEV), qsl, d["seq_lens"].to(DEV),
                                        d["ctx_lens"].to(DEV), 50, None, None, 1.0, 1.0)
    ref = f_rows.context_attention(d["q"], d["k"], d["v"], d["key_cache"], d["value_cache"], d["block_tables"],
                                   d["start_loc"], d["seq_lens"], d["ctx_lens"])
    assert not torch.isnan(out).any()
    torch.testing.assert_close(out.cpu().float(), ref.float(), atol=1e-2, rtol=1e-2)


def test_prefill_rejects_unsupported_shapes():
    from aphrodite_engine_b200.attention.prefix_prefill import context_attention_fwd
    d = _make_prefill_case(torch.float16, 2, 2, 24, 16, [8], [8], "auto", 8, seed=1)     # head size 24
    q = d["q"].to(DEV)
    with pytest.raises(RuntimeError, match="unsupported head size"):
        context_attention_fwd(q, d["k"].to(DEV), d["v"].to(DEV), torch.empty_like(q), "auto", d["key_cache"].to(DEV),
                              d["value_cache"].to(DEV), d["block_tables"].to(DEV), d["start_loc"].to(DEV),
                              d["seq_lens"].to(DEV), d["ctx_lens"].to(DEV), 8)


# ====================================================================================================== f4 fp8 quant
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("shape", [(1, 16), (7, 100), (33, 4096), (256, 14336), (5, 1030)])
def test_fp8_quant_bit_exact_vs_oracle(ops, dtype, shape):
    torch.manual_seed(shape[0] * 131 + shape[1])
    x = (torch.randn(*shape) * 4).to(dtype)
    x[0, 0] = 900.0                                           # saturates a static scale of 1
    xd = x.to(DEV)
    s = torch.tensor([0.37], dtype=torch.float32)
    out, _ = ops.scaled_fp8_quant(xd, s.to(DEV))
    assert torch.equal(out.cpu().view(torch.uint8), f_rows.static_scaled_fp8_quant(x, s).view(torch.uint8))
    out, sc = ops.scaled_fp8_quant(xd)
    ref, rsc = f_rows.dynamic_scaled_fp8_quant(x)
    assert torch.equal(sc.cpu(), rsc) and torch.equal(out.cpu().view(torch.uint8), ref.view(torch.uint8))
    for ub in (None, torch.tensor([3.0])):
        out, sc = ops.scaled_fp8_quant(xd, use_per_token_if_dynamic=True, scale_ub=None if ub is None else ub.to(DEV))
        ref, rsc = f_rows.dynamic_per_token_scaled_fp8_quant(x, ub)
        assert torch.equal(sc.cpu(), rsc), "per-token scales"
        assert torch.equal(out.cpu().view(torch.uint8), ref.view(torch.uint8))


def test_fp8_quant_padding_and_graph_capture(ops):
    x = torch.randn(5, 512, device=DEV, dtype=torch.bfloat16)
    out, sc = ops.scaled_fp8_quant(x, num_token_padding=17)
    assert out.shape == (17, 512) and sc.shape == (1,)
    s = torch.zeros(1, device=DEV)
    o = torch.empty(5, 512, device=DEV, dtype=torch.float8_e4m3fn)
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        torch.ops._C.dynamic_scaled_fp8_quant(o, x, s)
        st.synchronize()
        s.zero_()
        with torch.cuda.graph(g, stream=st):
            torch.ops._C.dynamic_scaled_fp8_quant(o, x, s)
    s.zero_()
    g.replay()
    torch.cuda.synchronize()
    ref, rsc = f_rows.dynamic_scaled_fp8_quant(x.cpu())
    assert torch.equal(s.cpu(), rsc) and torch.equal(o.cpu().view(torch.uint8), ref.view(torch.uint8))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_fp8_quant_vs_reference_kernels(ops, ref, dtype):
    torch.manual_seed(3)
    x = (torch.randn(65, 4096, device=DEV) * 5).to(dtype)
    s = torch.tensor([0.21], device=DEV)
    mine, _ = ops.scaled_fp8_quant(x, s)
    theirs = torch.empty_like(mine)
    ref.static_scaled_fp8_quant(theirs, x, s)
    assert torch.equal(mine.view(torch.uint8), theirs.view(torch.uint8))
    mine, ms = ops.scaled_fp8_quant(x)
    ts = torch.zeros(1, device=DEV)
    ref.dynamic_scaled_fp8_quant(theirs, x, ts)
    assert torch.equal(ms, ts) and torch.equal(mine.view(torch.uint8), theirs.view(torch.uint8))
    mine, ms = ops.scaled_fp8_quant(x, use_per_token_if_dynamic=True)
    ts = torch.empty(65, 1, device=DEV)
    ref.dynamic_per_token_scaled_fp8_quant(theirs, x, ts, None)
    assert torch.equal(ms, ts) and torch.equal(mine.view(torch.uint8), theirs.view(torch.uint8))


# ====================================================================================================== f4 scaled GEMM
def _mm_inputs(M, N, K, kind, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "fp8":
        a = torch.randn(M, K, generator=g).to(torch.float8_e4m3fn)
        w = torch.randn(N, K, generator=g).to(torch.float8_e4m3fn)          # the checkpoint's [N, K] weight
    else:
        a = torch.randint(-128, 128, (M, K), generator=g, dtype=torch.int8)
        w = torch.randint(-128, 128, (N, K), generator=g, dtype=torch.int8)
    return a, w


@pytest.mark.parametrize("kind", ["fp8", "int8"])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("mnk", [(1, 128, 128), (16, 256, 512), (33, 4096, 4096), (256, 6144, 4096), (256, 4096, 14336),
                                 (300, 1024, 1040), (512, 64, 256), (64, 28672, 4096)])
@pytest.mark.parametrize("scales", ["tensor", "token_channel"])
def test_scaled_mm_vs_oracle(ops, kind, out_dtype, mnk, scales):
    M, N, K = mnk
    a, w = _mm_inputs(M, N, K, kind, seed=M + N + K)
    g = torch.Generator().manual_seed(7)
    mag = (1.0 / math.sqrt(K)) if kind == "fp8" else (1.0 / (128.0 * math.sqrt(K)))
    if scales == "tensor":
        sa, sb = torch.tensor([0.9]), torch.tensor([mag])
    else:
        sa = torch.rand(M, 1, generator=g) + 0.5
        sb = (torch.rand(N, 1, generator=g) + 0.5) * mag
    bias = (torch.randn(N, generator=g)).to(out_dtype) if scales == "token_channel" else None
    out = ops.cutlass_scaled_mm(a.to(DEV), w.to(DEV).t(), sa.to(DEV), sb.to(DEV), out_dtype,
                                None if bias is None else bias.to(DEV))
    torch.cuda.synchronize()
    ref = f_rows.scaled_mm(a, w.t(), sa, sb, out_dtype, bias)
    # fp32 tensor-core accumulation over K terms against the exact sum, then ONE rounding to the 16-bit output
    eps = 2.0 ** -8 if out_dtype == torch.bfloat16 else 2.0 ** -11
    torch.testing.assert_close(out.cpu().float(), ref.float(), atol=4 * eps, rtol=2 * eps)


def test_scaled_mm_k_splits_are_bit_identical(cabi):
    """Every split plan through the C ABI (cluster of S CTAs, partial tiles in the fp32 scratch) gives the same bits for
    int8 (integer partial sums) and fp8 results within fp32 re-association of the unsplit result; repeated launches of
    one plan are bit-identical (deterministic reduction order)."""
    import ctypes
    M, N, K = 48, 512, 4096
    for kind, code in (("int8", 1), ("fp8", 0)):
        a, w = _mm_inputs(M, N, K, kind, seed=11)
        ad, wd = a.to(DEV), w.to(DEV)
        sa = torch.tensor([1.0], device=DEV)
        sb = torch.tensor([1.0 / 4096 if kind == "fp8" else 1.0 / (4096 * 64)], device=DEV)
        outs = {}
        scratch = torch.empty(8, M, N, dtype=torch.float32, device=DEV)
        for split in (1, 2, 3, 4, 8):
            o = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
            for rep in range(2):
                rc = cabi.b200_cutlass_scaled_mm(o.data_ptr(), ad.data_ptr(), wd.data_ptr(), sa.data_ptr(), sb.data_ptr(),
                                                 None, M, N, K, K, K, N, 1, 1, code, 2, split, scratch.data_ptr(),
                                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
                assert rc == 0, cabi.b200_last_error()
                torch.cuda.synchronize()
                if rep == 0:
                    first = o.clone()
                else:
                    assert torch.equal(first, o), f"{kind} split {split}: not deterministic"
            outs[split] = o.clone()
        ref = f_rows.scaled_mm(a, w.t(), sa.cpu(), sb.cpu(), torch.bfloat16)
        for split, o in outs.items():
            if kind == "int8":
                assert torch.equal(o, outs[1]), f"int8 split {split} differs from the unsplit result"
            torch.testing.assert_close(o.cpu().float(), ref.float(), atol=2 ** -6, rtol=2 ** -7)


@pytest.mark.parametrize("mnk", [(256, 6144, 4096), (48, 384, 1024), (300, 4096, 2048), (16, 128, 512)])
def test_scaled_mm_tile_shapes_agree(ops, cabi, mnk):
    """128- and 256-channel CTA tiles (b200_scaled_mm_set_tile) are two schedules of the same arithmetic: int8 results are
    bit-identical, fp8 results agree to fp32 re-association of the k-split partial sums."""
    M, N, K = mnk
    try:
        for kind in ("int8", "fp8"):
            a, w = _mm_inputs(M, N, K, kind, seed=5)
            sa = torch.rand(M, 1, generator=torch.Generator().manual_seed(1)) + 0.5
            sb = (torch.rand(N, 1, generator=torch.Generator().manual_seed(2)) + 0.5) * (1.0 / K if kind == "fp8" else 1.0 / (K * 64))
            outs = []
            for tile in (1, 2):
                cabi.b200_scaled_mm_set_tile(tile)
                outs.append(ops.cutlass_scaled_mm(a.to(DEV), w.to(DEV).t(), sa.to(DEV), sb.to(DEV), torch.bfloat16).cpu())
            if kind == "int8":
                assert torch.equal(outs[0], outs[1])
            ref = f_rows.scaled_mm(a, w.t(), sa, sb, torch.bfloat16)
            for o in outs:
                torch.testing.assert_close(o.float(), ref.float(), atol=2 ** -6, rtol=2 ** -7)
    finally:
        cabi.b200_scaled_mm_set_tile(0)


def test_scaled_mm_strided_operands_and_errors(ops):
    M, N, K = 40, 256, 512
    a, w = _mm_inputs(M, N, K + 128, "fp8", seed=2)
    ad, wd = a.to(DEV)[:, :K], w.to(DEV)[:, :K]                 # row strides K + 128: non-contiguous views
    sa, sb = torch.tensor([1.0], device=DEV), torch.tensor([0.05], device=DEV)
    out = ops.cutlass_scaled_mm(ad, wd.t(), sa, sb, torch.bfloat16)
    ref = f_rows.scaled_mm(a[:, :K], w[:, :K].t(), sa.cpu(), sb.cpu(), torch.bfloat16)
    torch.testing.assert_close(out.cpu().float(), ref.float(), atol=2 ** -6, rtol=2 ** -7)
    assert torch.ops._C.cutlass_scaled_mm_supports_fp8(100) is True
    with pytest.raises(RuntimeError, match="not implemented"):
        torch.ops._C.cutlass_scaled_mm_azp(out, ad, wd.t(), sa, sb, torch.zeros(N, dtype=torch.int32, device=DEV), None, None)
    with pytest.raises(RuntimeError):
        ops.cutlass_scaled_mm(ad, wd.t().contiguous(), sa, sb, torch.bfloat16)       # b must be column-major


# ====================================================================================================== f2 sampling
def _gpu_probs(B, V, seed, temp=3.0):
    g = torch.Generator().manual_seed(seed)
    return torch.softmax(torch.randn(B, V, generator=g) * temp, -1).float()


MARGIN = 2e-6        # decisions whose cdf-vs-u gap is below this may legitimately differ by fp32 summation order


@pytest.mark.parametrize("V", [1000, 32000, 128256, 50257])
def test_sampling_from_probs_vs_oracle(ops, V):
    B = 24
    p = _gpu_probs(B, V, V)
    u = torch.rand(B, generator=torch.Generator().manual_seed(1))
    u[0], u[1] = 0.0, 0.99999994
    ids = ops.sampling_from_probs(p.to(DEV), u.to(DEV)).cpu().numpy()
    ref, margins = f_rows.sampling_from_probs(p.numpy(), u.numpy())
    sure = margins > MARGIN
    assert sure.sum() >= B - 4
    assert (ids[sure] == ref[sure]).all(), (ids, ref, margins)
    cdf = np.cumsum(p.numpy().astype(np.float64), axis=1)
    for b in np.nonzero(~sure)[0]:                              # borderline rows: still the crossing, up to fp32 rounding
        i = int(ids[b])
        lo = cdf[b, i - 1] if i > 0 else 0.0
        assert (lo - 1e-5 <= u[b].item() < cdf[b, i] + 1e-5) or i == V - 1


@pytest.mark.parametrize("mode", ["top_k", "top_p", "min_p", "top_k_top_p"])
@pytest.mark.parametrize("V", [2000, 128256])
@pytest.mark.parametrize("per_row", [False, True])
def test_rejection_samplers_vs_oracle(ops, mode, V, per_row):
    B, R = 16, 32
    p = _gpu_probs(B, V, V + 1)
    u = torch.rand(R, B, generator=torch.Generator().manual_seed(2))
    rng = np.random.default_rng(3)
    k_arr = rng.integers(1, 60, B).astype(np.int32)
    p_arr = (rng.random(B) * 0.6 + 0.3).astype(np.float32)
    if mode == "min_p":
        p_arr = (rng.random(B) * 0.3 + 0.01).astype(np.float32)
    k_val, p_val = int(k_arr[0]), float(p_arr[0])
    kt = torch.from_numpy(k_arr).to(DEV) if per_row else None
    pt = torch.from_numpy(p_arr).to(DEV) if per_row else None
    pd, ud = p.to(DEV), u.to(DEV)
    if mode == "top_k":
        ids, ok = ops.top_k_sampling_from_probs(pd, ud, kt, k_val)
    elif mode == "top_p":
        ids, ok = ops.top_p_sampling_from_probs(pd, ud, pt, p_val)
    elif mode == "min_p":
        ids, ok = ops.min_p_sampling_from_probs(pd, ud, pt, p_val)
    else:
        ids, ok = ops.top_k_top_p_sampling_from_probs(pd, ud, kt, k_val, pt, p_val)
    rid, rok, margins = f_rows.rejection_sampling(mode, p.numpy(), u.numpy(), k=k_arr if per_row else k_val,
                                                  p=p_arr if per_row else p_val)
    assert ids.dtype == torch.int32 and ok.dtype == torch.bool
    sure = margins > MARGIN
    assert sure.sum() >= B - 3
    assert (ids.cpu().numpy()[sure] == rid[sure]).all(), (ids, rid, margins)
    assert (ok.cpu().numpy()[sure] == rok[sure]).all()


def test_rejection_sampler_reports_failure_when_rounds_run_out(ops):
    V = 4096
    p = torch.full((4, V), 0.9 / (V - 1))
    p[:, 0] = 0.1                                                            # the only top-1 entry
    u = torch.full((1, 4), 0.5001)                                           # one round, lands in the flat tail (margin 1e-4)
    ids, ok = ops.top_k_sampling_from_probs(p.to(DEV), u.to(DEV), None, 1)
    rid, rok, _ = f_rows.rejection_sampling("top_k", p.numpy(), u.numpy(), k=1)
    assert not rok.any() and not ok.any()                                    # entry 0 is still above the pivot
    assert (ids.cpu().numpy() == rid).all()
    u2 = torch.cat([u, torch.full((1, 4), 0.05)])                            # a second round draws inside the top-1 mass
    ids, ok = ops.top_k_sampling_from_probs(p.to(DEV), u2.to(DEV), None, 1)
    assert ok.all() and (ids == 0).all()


@pytest.mark.parametrize("V", [1500, 128256])
def test_renorm_and_mask_vs_oracle(ops, V):
    B = 8
    p = _gpu_probs(B, V, 9)
    k = np.array([1, 2, 5, 50, 1000, V - 1, V, V + 5], dtype=np.int32)
    out = ops.top_k_renorm_prob(p.to(DEV), torch.from_numpy(k).to(DEV), 0).cpu().numpy()
    ref = f_rows.top_k_renorm_prob(p.numpy(), k)
    assert ((out > 0) == (ref > 0)).all()
    np.testing.assert_allclose(out, ref, rtol=2e-5, atol=1e-9)
    out = ops.top_k_renorm_prob(p.to(DEV), None, 40).cpu().numpy()
    np.testing.assert_allclose(out, f_rows.top_k_renorm_prob(p.numpy(), 40), rtol=2e-5, atol=1e-9)
    logits = torch.randn(B, V, generator=torch.Generator().manual_seed(4))
    out = ops.top_k_mask_logits(logits.to(DEV), torch.from_numpy(k).to(DEV), 0).cpu().numpy()
    assert np.array_equal(out, f_rows.top_k_mask_logits(logits.numpy(), k))
    tp = np.array([0.1, 0.5, 0.9, 0.99, 1e-6, 0.3, 0.7, 0.95], dtype=np.float32)
    out = ops.top_p_renorm_prob(p.to(DEV), torch.from_numpy(tp).to(DEV), 0.0).cpu().numpy()
    ref = f_rows.top_p_renorm_prob(p.numpy(), tp)
    same_support = ((out > 0) == (ref > 0)).all(axis=1)
    assert same_support.sum() >= B - 1                           # a mass within rounding of p may move the cut by one entry
    np.testing.assert_allclose(out[same_support], ref[same_support], rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(out.sum(1), 1.0, rtol=1e-4)


def test_sampling_vs_reference_kernels(ops, ref):
    """Same probabilities and the same uniforms through the reference's own kernels: identical tokens / masks wherever
    the reference itself is reproducible (its deterministic mode)."""
    B, V, R = 32, 128256, 32
    p = _gpu_probs(B, V, 21).to(DEV)
    u1 = torch.rand(B, generator=torch.Generator().manual_seed(5)).to(DEV)
    u = torch.rand(R, B, generator=torch.Generator().manual_seed(6)).to(DEV)
    mism = (ops.sampling_from_probs(p, u1) != ref.sampling_from_probs(p, u1, True)).sum().item()
    assert mism <= 1
    for mine, theirs in (
        (ops.top_k_sampling_from_probs(p, u, None, 50), ref.top_k_sampling_from_probs(p, u, None, 50, True)),
        (ops.top_p_sampling_from_probs(p, u, None, 0.9), ref.top_p_sampling_from_probs(p, u, None, 0.9, True)),
        (ops.min_p_sampling_from_probs(p, u, None, 0.05), ref.min_p_sampling_from_probs(p, u, None, 0.05, True)),
        (ops.top_k_top_p_sampling_from_probs(p, u, None, 50, None, 0.9),
         ref.top_k_top_p_sampling_from_probs(p, u, None, 50.0, None, 0.9, True)),
    ):
        assert (mine[0] != theirs[0]).sum().item() <= 1 and (mine[1] != theirs[1]).sum().item() <= 1
    a, b = ops.top_k_renorm_prob(p, None, 64), ref.top_k_renorm_prob(p, None, 64)
    assert ((a > 0) == (b > 0)).all()
    torch.testing.assert_close(a, b, rtol=2e-5, atol=1e-9)
    a, b = ops.top_p_renorm_prob(p, None, 0.92), ref.top_p_renorm_prob(p, None, 0.92)
    rows = ((a > 0) == (b > 0)).all(dim=1)
    assert rows.sum().item() >= B - 1
    torch.testing.assert_close(a[rows], b[rows], rtol=2e-5, atol=1e-9)
    logits = torch.randn(B, V, device=DEV)
    assert torch.equal(ops.top_k_mask_logits(logits, None, 100), ref.top_k_mask_logits(logits, None, 100))


# ====================================================================================================== CUDA graphs
def test_f_rows_are_cuda_graph_capturable(ops):
    """Decode and prefill run under torch.cuda.graph in the reference (worker/model_runner.py:1682+): the W8A8 GEMM (k-split
    path: scratch from the caching allocator, cluster launch), the quantiser, a rejection sampler and the prefill attention
    are captured once and replayed on new inputs in the same buffers."""
    from aphrodite_engine_b200.attention.prefix_prefill import context_attention_fwd
    g = torch.Generator().manual_seed(0)
    M, N, K = 64, 512, 4096                                       # 4 tiles -> k-split cluster
    x = torch.randn(M, K, generator=g).to(torch.bfloat16).to(DEV)
    w8, sw = ops.scaled_fp8_quant((torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(DEV), use_per_token_if_dynamic=True)
    probs = torch.softmax(torch.randn(8, 4000, generator=g) * 3, -1).to(DEV)
    uni = torch.rand(32, 8, generator=g).to(DEV)
    d = _make_prefill_case(torch.bfloat16, 4, 2, 128, 16, [40, 7], [50, 64], "auto", 8, seed=3)
    dd = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    o = torch.empty_like(dd["q"])
    out = {}

    def step():
        a8, sa = ops.scaled_fp8_quant(x, use_per_token_if_dynamic=True)
        out["mm"] = ops.cutlass_scaled_mm(a8, w8.t(), sa, sw, torch.bfloat16)
        out["ids"], out["ok"] = ops.top_k_sampling_from_probs(probs, uni, None, 10)
        context_attention_fwd(dd["q"], dd["k"], dd["v"], o, "auto", dd["key_cache"], dd["value_cache"], dd["block_tables"],
                              dd["start_loc"], dd["seq_lens"], dd["ctx_lens"], 64)

    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        step()
        st.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            step()
    # new inputs in the captured buffers
    x.copy_(torch.randn(M, K, generator=g).to(torch.bfloat16))
    probs.copy_(torch.softmax(torch.randn(8, 4000, generator=g) * 3, -1))
    dd["q"].copy_((torch.randn(dd["q"].shape, generator=g) * 0.5).to(torch.bfloat16))
    graph.replay()
    torch.cuda.synchronize()
    a8, sa = f_rows.dynamic_per_token_scaled_fp8_quant(x.cpu())
    ref = f_rows.scaled_mm(a8, w8.cpu().t(), sa, sw.cpu(), torch.bfloat16)
    torch.testing.assert_close(out["mm"].cpu().float(), ref.float(), atol=2 ** -6, rtol=2 ** -7)
    rid, rok, marg = f_rows.rejection_sampling("top_k", probs.cpu().numpy(), uni.cpu().numpy(), k=10)
    sure = marg > MARGIN
    assert (out["ids"].cpu().numpy()[sure] == rid[sure]).all()
    refo = f_rows.context_attention(dd["q"].cpu(), d["k"], d["v"], d["key_cache"], d["value_cache"], d["block_tables"],
                                    d["start_loc"], d["seq_lens"], d["ctx_lens"])
    torch.testing.assert_close(o.cpu().float(), refo.float(), atol=1e-2, rtol=1e-2)
