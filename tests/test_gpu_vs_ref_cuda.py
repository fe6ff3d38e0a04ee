"""GPU parity against the reference's OWN CUDA kernels, recompiled for sm_100a (oracle/build_ref_cuda.py ->
oracle/_ref/_ref_cuda_C.so, registered as torch.ops._ref_cuda_C.*): same process, same device tensors, both
implementations called through their torch ops. Integer / byte / index work and 16-bit rotary must be bit-exact;
floating-point ops are held to the tolerances of tests/tolerances.py (tighter than the reference's own tests).
Skipped when the prebuilt library is absent (it is built where /root/reference exists and shipped to the GPU box)."""
import os
import random

import pytest
import torch

from oracle import marlin as om
from oracle import paged_ops as po
from tests import tolerances as tol

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REF_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "_ref_cuda_C.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/_ref_cuda_C.so not built (needs /root/reference at build time)")
    torch.ops.load_library(REF_SO)
    return torch.ops._ref_cuda_C


def _types():
    from aphrodite_engine_b200.scalar_type import scalar_types
    return scalar_types


# ---------------------------------------------------------------------------------------------- attention
@pytest.mark.parametrize("version", ["v1", "v2"])
@pytest.mark.parametrize("cfg", [  # (seqs, Hq, Hkv, D, BS, dtype, kv_dtype, alibi, max_len)
    (7, 32, 8, 128, 16, torch.bfloat16, "auto", False, 700),
    (5, 8, 8, 64, 16, torch.float16, "auto", True, 333),
    (3, 16, 2, 256, 32, torch.bfloat16, "auto", False, 1300),
    (4, 12, 4, 120, 16, torch.float16, "auto", False, 100),
    (4, 8, 2, 80, 8, torch.float32, "auto", False, 90),
    (6, 32, 8, 128, 16, torch.bfloat16, "fp8", False, 600),
    (6, 4, 1, 128, 16, torch.float16, "fp8_e5m2", False, 1111),
    (2, 40, 40, 128, 16, torch.float16, "auto", True, 2049),
])
def test_paged_attention_vs_reference_kernel(ops, ref, version, cfg):
    S, H, Hkv, D, BS, dtype, kv_dtype, use_alibi, max_len = cfg
    torch.manual_seed(S * H + D)
    random.seed(D)
    lens = [random.randint(1, max_len) for _ in range(S)]
    lens[0] = max_len
    nb_per = (max_len + BS - 1) // BS
    NB = S * nb_per + 3
    scale = D ** -0.5
    q = torch.empty(S, H, D).uniform_(-scale, scale).to(dtype).to(DEV)
    kc, vc = po.make_kv_cache(NB, BS, Hkv, D, dtype, kv_dtype, seed=1)
    kc, vc = kc.to(DEV), vc.to(DEV)
    bt = torch.randperm(NB)[: S * nb_per].view(S, nb_per).to(torch.int32).to(DEV)
    sl = torch.tensor(lens, dtype=torch.int32, device=DEV)
    alibi = torch.randn(H, dtype=torch.float32, device=DEV) if use_alibi else None
    ks, vs = (0.75, 1.5) if kv_dtype != "auto" else (1.0, 1.0)
    outs = []
    for impl in ("mine", "ref"):
        out = torch.full_like(q, float("nan"))
        if version == "v1":
            args = (out, q, kc, vc, Hkv, scale, bt, sl, BS, max_len, alibi, kv_dtype, ks, vs, 0, 0, 0, 64, 0)
            (ops.paged_attention_v1 if impl == "mine" else ref.paged_attention_v1)(*args)
        else:
            P = (max_len + 511) // 512
            tmp = torch.empty(S, H, P, D, dtype=dtype, device=DEV)
            es = torch.empty(S, H, P, dtype=torch.float32, device=DEV)
            ml = torch.empty(S, H, P, dtype=torch.float32, device=DEV)
            args = (out, es, ml, tmp, q, kc, vc, Hkv, scale, bt, sl, BS, max_len, alibi, kv_dtype, ks, vs, 0, 0, 0, 64, 0)
            (ops.paged_attention_v2 if impl == "mine" else ref.paged_attention_v2)(*args)
        torch.cuda.synchronize()
        outs.append(out.float().cpu())
    atol = tol.ATTN_FP8_ATOL if kv_dtype != "auto" else tol.ATTN_ATOL
    assert torch.isfinite(outs[0]).all()
    torch.testing.assert_close(outs[0], outs[1], atol=atol, rtol=tol.ATTN_RTOL)


# ---------------------------------------------------------------------------------------------- cache ops
@pytest.mark.parametrize("kv_dtype", ["auto", "fp8", "fp8_e5m2"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_reshape_and_cache_vs_reference_kernel(ops, ref, dtype, kv_dtype):
    T, Hkv, D, BS, NB = 83, 8, 128, 16, 64
    torch.manual_seed(3)
    qkv = torch.randn(T, 3, Hkv, D, dtype=dtype, device=DEV)
    key, value = qkv[:, 1], qkv[:, 2]                       # strided views, as out of the fused qkv GEMM
    slots = torch.randperm(NB * BS)[:T].to(torch.int64).to(DEV)
    slots[5] = -1                                           # padding token: skipped
    cdt = dtype if kv_dtype == "auto" else torch.uint8
    x = 16 // torch.tensor([], dtype=cdt).element_size()
    caches = []
    for fn in (ops.reshape_and_cache, ref.reshape_and_cache):
        kc = torch.zeros(NB, Hkv, D // x, BS, x, dtype=cdt, device=DEV)
        vc = torch.zeros(NB, Hkv, D, BS, dtype=cdt, device=DEV)
        fn(key, value, kc, vc, slots, kv_dtype, 0.5, 2.0)
        caches.append((kc, vc))
    torch.cuda.synchronize()
    assert torch.equal(caches[0][0], caches[1][0]) and torch.equal(caches[0][1], caches[1][1])
    flash = []
    for fn in (ops.reshape_and_cache_flash, ref.reshape_and_cache_flash):
        kc = torch.zeros(NB, BS, Hkv, D, dtype=cdt, device=DEV)
        vc = torch.zeros_like(kc)
        fn(key, value, kc, vc, slots, kv_dtype, 0.5, 2.0)
        flash.append((kc, vc))
    torch.cuda.synchronize()
    assert torch.equal(flash[0][0], flash[1][0]) and torch.equal(flash[0][1], flash[1][1])


def test_copy_blocks_and_convert_fp8_vs_reference_kernel(ops, ref):
    torch.manual_seed(4)
    L, NB = 3, 40
    mapping = torch.tensor([[0, 7], [0, 9], [3, 11], [38, 1]], dtype=torch.int64, device=DEV)
    base = [(torch.randn(NB, 8, 16, 16, 8, dtype=torch.bfloat16, device=DEV),
             torch.randn(NB, 8, 128, 16, dtype=torch.bfloat16, device=DEV)) for _ in range(L)]
    res = []
    for fn in (ops.copy_blocks, ref.copy_blocks):
        kcs = [k.clone() for k, _ in base]
        vcs = [v.clone() for _, v in base]
        fn(kcs, vcs, mapping)
        torch.cuda.synchronize()
        res.append((kcs, vcs))
    for l in range(L):
        assert torch.equal(res[0][0][l], res[1][0][l]) and torch.equal(res[0][1][l], res[1][1][l])
    src = torch.randn(4, 8, 128, 16, dtype=torch.float16, device=DEV) * 3
    for kvd in ("fp8", "fp8_e4m3"):        # the reference's convert_fp8 has no e5m2 branch (cache_kernels.cu:372-408)
        a, b = (torch.zeros(src.shape, dtype=torch.uint8, device=DEV) for _ in range(2))
        ops.convert_fp8(a, src, 0.5, kvd)
        ref.convert_fp8(b, src, 0.5, kvd)
        assert torch.equal(a, b)
        back_a, back_b = torch.zeros_like(src), torch.zeros_like(src)
        ops.convert_fp8(back_a, a, 0.5, kvd)
        ref.convert_fp8(back_b, b, 0.5, kvd)
        assert torch.equal(back_a, back_b)


# ---------------------------------------------------------------------------------------------- norm / rope / act
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_norm_rope_act_vs_reference_kernel(ops, ref, dtype):
    torch.manual_seed(6)
    T, H = 67, 4096
    x = torch.randn(T, H, dtype=dtype, device=DEV)
    w = (torch.randn(H, device=DEV) * 0.1 + 1).to(dtype)
    ulp = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10
    o1, o2 = torch.empty_like(x), torch.empty_like(x)
    ops.rms_norm(o1, x, w, 1e-5)
    ref.rms_norm(o2, x, w, 1e-5)
    torch.testing.assert_close(o1.float(), o2.float(), atol=1e-6, rtol=2.5 * ulp)       # <= 2 ulp (reduction order)
    assert (o1 != o2).float().mean() < 0.02
    xs, rs = [], []
    for fn in (ops.fused_add_rms_norm, ref.fused_add_rms_norm):
        xi, ri = x.clone(), torch.randn(T, H, dtype=dtype, device=DEV, generator=torch.Generator(DEV).manual_seed(1))
        fn(xi, ri, w, 1e-5)
        xs.append(xi)
        rs.append(ri)
    assert torch.equal(rs[0], rs[1])                                                   # residual = rounded sum: exact
    torch.testing.assert_close(xs[0].float(), xs[1].float(), atol=1e-6, rtol=2.5 * ulp)
    # rotary: every product and sum is rounded to the 16-bit type in both implementations -> bit-exact
    Hq, Hkv, D = 32, 8, 128
    pos = torch.randint(0, 4096, (T,), dtype=torch.int64, device=DEV)
    cache = torch.randn(4096, D, dtype=dtype, device=DEV)
    for neox in (True, False):
        got = []
        for fn in (ops.rotary_embedding, ref.rotary_embedding):
            qkv = torch.randn(T, (Hq + 2 * Hkv) * D, dtype=dtype, device=DEV,
                              generator=torch.Generator(DEV).manual_seed(2))
            q, k = qkv[:, : Hq * D], qkv[:, Hq * D: (Hq + Hkv) * D]
            fn(pos, q, k, D, cache, neox)
            got.append(qkv)
        assert torch.equal(got[0], got[1])
    d = 14336
    xin = torch.randn(T, 2 * d, dtype=dtype, device=DEV)
    for name, exact in (("silu_and_mul", True), ("gelu_and_mul", False), ("gelu_tanh_and_mul", False)):
        a, b = torch.empty(T, d, dtype=dtype, device=DEV), torch.empty(T, d, dtype=dtype, device=DEV)
        getattr(ops, name)(a, xin)
        getattr(ref, name)(b, xin)
        if exact:
            assert torch.equal(a, b), name
        else:
            torch.testing.assert_close(a.float(), b.float(), atol=1e-5, rtol=2.5 * ulp)
    for name in ("gelu_new", "gelu_fast", "gelu_quick"):
        a, b = torch.empty(T, 2 * d, dtype=dtype, device=DEV), torch.empty(T, 2 * d, dtype=dtype, device=DEV)
        getattr(ops, name)(a, xin)
        getattr(ref, name)(b, xin)
        torch.testing.assert_close(a.float(), b.float(), atol=1e-5, rtol=2.5 * ulp)


# ---------------------------------------------------------------------------------------------- marlin
def _ref_gemm(ref, a, mq, ms, mz, g_idx, perm, ws, st, M, N, K, has_zp, zp_float=False):
    return ref.gptq_marlin_gemm(a, mq, ms, mz, g_idx, perm, ws, st.exponent, st.mantissa, st.bias, st.signed,
                                M, N, K, True, has_zp, True, zp_float)


def _close(a, b, frac=None):
    # one output ulp of slack on the largest magnitudes: the reference rounds the accumulator to the output type and
    # then applies channel-wise scales in that type (gptq_marlin.cu:1690-1717), this repo scales the weights
    if frac is None:
        frac = 1e-2 if a.dtype == torch.bfloat16 else 5e-3
    a, b = a.float().cpu(), b.float().cpu()
    assert torch.isfinite(a).all()
    scale = b.abs().max().clamp(min=1e-6)
    assert ((a - b).abs().max() / scale) < frac, float((a - b).abs().max() / scale)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("group_size", [-1, 128])
@pytest.mark.parametrize("mkn", [(16, 512, 256), (256, 4096, 1024), (300, 1024, 448)])
def test_marlin_gemm_vs_reference_kernel(ops, ref, dtype, bits, group_size, mkn):
    M, K, N = mkn
    torch.manual_seed(M + K + N + bits)
    a = (torch.randn(M, K) * 0.5).to(dtype).to(DEV)
    w = torch.randn(K, N).to(dtype)
    empty = torch.empty(0, dtype=torch.int32, device=DEV)
    ws = lambda: torch.zeros((N // 64) * 16, dtype=torch.int32, device=DEV)
    st = _types().uint4b8 if bits == 4 else _types().uint8b128
    _, mq, ms = om.marlin_quantize(w, bits, group_size)
    mq, ms = mq.to(DEV), ms.to(DEV)
    mine = ops.gptq_marlin_gemm(a, mq, ms, empty, empty, empty, ws(), st, M, N, K, True, False, True, False)
    theirs = _ref_gemm(ref, a, mq, ms, empty, empty, empty, ws(), st, M, N, K, False)
    torch.cuda.synchronize()
    _close(mine, theirs)
    st = _types().uint4 if bits == 4 else _types().uint8
    _, mq, ms, mz = om.awq_marlin_quantize(w, bits, group_size)
    mq, ms, mz = mq.to(DEV), ms.to(DEV), mz.to(DEV)
    mine = ops.gptq_marlin_gemm(a, mq, ms, mz, empty, empty, ws(), st, M, N, K, True, True, True, False)
    theirs = _ref_gemm(ref, a, mq, ms, mz, empty, empty, ws(), st, M, N, K, True)
    torch.cuda.synchronize()
    _close(mine, theirs)


def test_marlin_act_order_hqq_and_repack_vs_reference_kernel(ops, ref):
    M, K, N, G = 64, 1024, 512, 128
    torch.manual_seed(8)
    a = torch.randn(M, K).half().to(DEV)
    w = torch.randn(K, N).half()
    _, mq, ms, g_idx, sort_idx = om.marlin_quantize_act_order(w, 4, G, seed=5)
    ws = lambda: torch.zeros((N // 64) * 16, dtype=torch.int32, device=DEV)
    args = (a, mq.to(DEV), ms.to(DEV), torch.empty(0, dtype=torch.int32, device=DEV), g_idx.to(DEV), sort_idx.to(DEV))
    mine = ops.gptq_marlin_gemm(*args, ws(), _types().uint4b8, M, N, K, True, False, True, False)
    theirs = _ref_gemm(ref, *args, ws(), _types().uint4b8, M, N, K, False)
    _close(mine, theirs)
    g = torch.Generator().manual_seed(1)
    q = torch.randint(0, 16, (K, N), generator=g)
    s = (torch.rand(K // G, N, generator=g) * 0.02 + 0.005).half()
    zp = (torch.rand(K // G, N, generator=g) * 4 + 6).half()
    w_ref, mq, ms, mz = om.hqq_marlin_quantize(q, s, zp, G)
    empty = torch.empty(0, dtype=torch.int32, device=DEV)
    mine = ops.gptq_marlin_gemm(a, mq.to(DEV), ms.to(DEV), mz.to(DEV), empty, empty, ws(), _types().uint4, M, N, K,
                                True, True, False, True)
    theirs = _ref_gemm(ref, a, mq.to(DEV), ms.to(DEV), mz.to(DEV), empty, empty, ws(), _types().uint4, M, N, K, True, True)
    # float zero points: the documented semantics are W = (q - zp) * s (the upstream test of this path); the
    # reference kernel of this snapshot, recompiled for sm_100a, does not reproduce them (observed on the B200:
    # outputs off by an order of magnitude), so the product is held to the oracle and the reference kernel's
    # deviation is only reported
    expect = om.marlin_gemm(a.cpu(), w_ref)
    _close(mine, expect)
    dev = float((theirs.float().cpu() - expect.float()).abs().max() / expect.float().abs().max())
    if dev > 1e-2:
        import warnings
        warnings.warn(f"reference HQQ (is_zp_float) kernel deviates from (q - zp) * s by {dev:.3f} of the output range")
    # repack: integer re-tiling, bit-exact
    for bits in (4, 8):
        qw = torch.randint(0, 2 ** bits, (K, N), generator=g).int()
        packed = om.pack_rows(qw, bits).to(DEV)
        perm = torch.randperm(K, generator=g).int().to(DEV)
        for p in (torch.empty(0, dtype=torch.int32, device=DEV), perm):
            assert torch.equal(ops.gptq_marlin_repack(packed, p, K, N, bits), ref.gptq_marlin_repack(packed, p, K, N, bits))
        awq = om.awq_pack(qw, bits).to(DEV)
        assert torch.equal(ops.awq_marlin_repack(awq, K, N, bits), ref.awq_marlin_repack(awq, K, N, bits))


# ---------------------------------------------------------------------------------------------- MoE
@pytest.mark.parametrize("cfg", [(33, 8, 2, 16), (128, 8, 2, 64), (200, 60, 6, 16)])
def test_moe_routing_and_grouped_gemm_vs_reference_kernel(ops, ref, cfg):
    from aphrodite_engine_b200 import fused_moe as fm
    M, E, topk, block = cfg
    K, N = 512, 256
    torch.manual_seed(M)
    gate = torch.randn(M, E, dtype=torch.float32, device=DEV) * 2
    outs = []
    for fn in (ops.topk_softmax, ref.topk_softmax):
        w = torch.empty(M, topk, dtype=torch.float32, device=DEV)
        ids = torch.empty(M, topk, dtype=torch.int32, device=DEV)
        src = torch.empty(M, topk, dtype=torch.int32, device=DEV)
        fn(w, ids, src, gate)
        outs.append((w, ids, src))
    torch.cuda.synchronize()
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    torch.testing.assert_close(outs[0][0], outs[1][0], atol=1e-6, rtol=1e-5)
    tw, ids = outs[0][0], outs[0][1]
    numel = ids.numel()
    cap = numel + E * (block - 1)
    al = []
    for fn in (ops.moe_align_block_size, ref.moe_align_block_size):
        sorted_ids = torch.full((cap,), numel, dtype=torch.int32, device=DEV)
        expert_ids = torch.full(((cap + block - 1) // block,), -1, dtype=torch.int32, device=DEV)
        post = torch.zeros(1, dtype=torch.int32, device=DEV)
        fn(ids, E, block, sorted_ids, expert_ids, post)
        al.append((sorted_ids, expert_ids, post))
    torch.cuda.synchronize()
    n = int(al[0][2])
    assert torch.equal(al[0][2], al[1][2]) and torch.equal(al[0][0][:n], al[1][0][:n])
    assert torch.equal(al[0][1][: n // block], al[1][1][: n // block])
    # grouped W4A16 GEMM (the reference kernel is fp16-only)
    g = torch.Generator().manual_seed(E)
    qs, ss = [], []
    for e in range(E):
        _, mq, ms = om.marlin_quantize((torch.randn(K, N, generator=g) * 0.1).half(), 4, 128)
        qs.append(mq)
        ss.append(ms)
    q, s = torch.stack(qs).to(DEV), torch.stack(ss).to(DEV)
    a = torch.randn(M, K, dtype=torch.float16, device=DEV)
    none = torch.empty(E, 0, dtype=torch.int32, device=DEV)
    sorted_ids = al[0][0]
    for replicate, apply_w in ((True, False), (True, True)):
        res = []
        for fn in (torch.ops._moe_C.marlin_gemm_moe, ref.marlin_gemm_moe):
            ws = torch.zeros(((M + 255) // 256) * (N // 64) * 16, dtype=torch.int32, device=DEV)
            res.append(fn(a, q, sorted_ids, tw, ids, s, none, none, ws, M, N, K, True, E, topk, block, replicate, apply_w))
        torch.cuda.synchronize()
        _close(res[0], res[1])


# ---------------------------------------------------------------------------------------------- misc integer ops
def test_misc_ops_vs_reference_kernel(ops, ref):
    torch.manual_seed(10)
    a = torch.randn(37, 512, dtype=torch.float16, device=DEV)
    perm = torch.randperm(512, device=DEV).int()
    assert torch.equal(ops.permute_cols(a, perm), ref.permute_cols(a, perm))
    K, N, G = 256, 512, 128
    qweight = torch.randint(-2 ** 31, 2 ** 31 - 1, (K, N // 8), dtype=torch.int32, device=DEV)
    zeros = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // G, N // 8), dtype=torch.int32, device=DEV)
    scales = (torch.rand(K // G, N, device=DEV) * 0.02).half()
    assert torch.equal(ops.awq_dequantize(qweight, scales, zeros, 0, 0, 0), ref.awq_dequantize(qweight, scales, zeros, 0, 0, 0))
    S, BS = 19, 16
    st = []
    for fn in (ops.advance_step_flashattn, ref.advance_step_flashattn):
        g = torch.Generator(DEV).manual_seed(3)
        tokens = torch.zeros(S, dtype=torch.int64, device=DEV)
        sampled = torch.randint(0, 32000, (S - 4, 1), dtype=torch.int64, device=DEV, generator=g)
        positions = torch.randint(0, 500, (S,), dtype=torch.int64, device=DEV, generator=g)
        seq_lens = (positions + 1).int()
        slots = torch.zeros(S, dtype=torch.int64, device=DEV)
        tables = torch.randint(0, 1000, (S, 40), dtype=torch.int32, device=DEV, generator=g)
        fn(S, S - 4, BS, tokens, sampled, positions, seq_lens, slots, tables)
        st.append((tokens, positions, seq_lens, slots))
    torch.cuda.synchronize()
    for x, y in zip(*st):
        assert torch.equal(x, y)


@pytest.mark.parametrize("kn", [(4096, 1536), (1024, 4096), (4096, 7168), (3584, 4096),      # Llama-3-8B linears at TP=4
                                (4096, 768), (512, 4096), (4096, 3584), (1792, 4096),       # ... at TP=8
                                (2048, 4096), (7168, 4096), (4096, 3072), (4096, 14336)])   # ... at TP=2
@pytest.mark.parametrize("M", [256, 64])
def test_marlin_tp_shard_shapes_vs_reference_kernel(ops, ref, kn, M):
    """The shapes a tensor-parallel GPTQ decode step launches (bench_secondary cfg3 at N = 2 / 4 / 8), with random
    packed words and scales as the benchmark's dummy weights have them: this repo's kernel against the reference's."""
    K, N = kn
    g = torch.Generator(device=DEV).manual_seed(K + N + M)
    a = (torch.randn(M, K, generator=g, device=DEV) * 0.5).to(torch.bfloat16)
    q = torch.randint(-2**31, 2**31 - 1, (K // 16, N * 2), generator=g, device=DEV, dtype=torch.int32)
    s = (torch.rand(K // 128, N, generator=g, device=DEV) * 0.004 + 0.001).to(torch.bfloat16)
    empty = torch.empty(0, dtype=torch.int32, device=DEV)
    ws = lambda: torch.zeros((N // 64) * 16, dtype=torch.int32, device=DEV)
    st = _types().uint4b8
    mine = ops.gptq_marlin_gemm(a, q, s, empty, empty, empty, ws(), st, M, N, K, True, False, True, False)
    theirs = _ref_gemm(ref, a, q, s, empty, empty, empty, ws(), st, M, N, K, False)
    torch.cuda.synchronize()
    _close(mine, theirs)
