"""GPU parity: Marlin repack kernels (bit-exact) and the tcgen05 W4A16 GEMM vs the CPU oracle
(oracle/marlin.py, pinned against the reference's Python by tests/test_oracle_marlin_cpu.py).
Mirrors tests/kernels/test_marlin_gemm.py of the reference: repack equality (:127, :178) and
GEMM error vs `a @ w_ref` (:236-254, mean relative error < 0.04 there; far tighter here because the
dequantised weights are bit-identical to w_ref and only the fp32 accumulation order differs)."""
import pytest
import torch

from oracle import marlin as om

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _types():
    from aphrodite_engine_b200.scalar_type import scalar_types
    return scalar_types


@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("shape", [(128, 64), (256, 128), (1024, 448), (4096, 1024)])
@pytest.mark.parametrize("act_order", [False, True])
def test_gptq_marlin_repack_bit_exact(ops, bits, shape, act_order):
    K, N = shape
    g = torch.Generator().manual_seed(K + N + bits)
    q_w = torch.randint(0, 1 << bits, (K, N), generator=g, dtype=torch.int32)
    packed = om.pack_rows(q_w, bits)
    perm = torch.randperm(K, generator=g).int() if act_order else torch.empty(0, dtype=torch.int32)
    out = ops.gptq_marlin_repack(packed.to(DEV), perm.to(DEV), K, N, bits)
    ref = om.gptq_marlin_repack(packed, perm if act_order else None, K, N, bits)
    assert out.dtype == torch.int32 and tuple(out.shape) == (K // 16, N * 16 // (32 // bits))
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("shape", [(128, 64), (512, 320), (4096, 1024)])
def test_awq_marlin_repack_bit_exact(ops, bits, shape):
    K, N = shape
    g = torch.Generator().manual_seed(K * 3 + N + bits)
    q_w = torch.randint(0, 1 << bits, (K, N), generator=g, dtype=torch.int32)
    packed = om.awq_pack(q_w, bits)
    out = ops.awq_marlin_repack(packed.to(DEV), K, N, bits)
    assert torch.equal(out.cpu(), om.awq_marlin_repack(packed, K, N, bits))
    assert torch.equal(out.cpu(), om.marlin_weights(q_w, bits))


def _check_gemm(out, ref):
    out, ref = out.float().cpu(), ref.float()
    assert torch.isfinite(out).all()
    scale = ref.abs().max().clamp(min=1e-6)
    assert ((out - ref).abs().max() / scale) < 1e-2, float((out - ref).abs().max() / scale)
    assert ((out - ref).abs().mean() / ref.abs().mean().clamp(min=1e-6)) < 2e-3


def _workspace(N):
    return torch.zeros((N // 64) * 16, dtype=torch.int32, device=DEV)


GEMM_SHAPES = [  # (M, K, N)
    (1, 128, 64), (7, 256, 128), (16, 512, 192), (33, 1024, 256), (100, 256, 448), (128, 2048, 512),
    (256, 4096, 1024), (257, 512, 128), (300, 1024, 256), (513, 256, 64),
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("group_size", [-1, 32, 64, 128])
@pytest.mark.parametrize("mkn", GEMM_SHAPES)
def test_gptq_marlin_gemm_uint4b8(ops, dtype, group_size, mkn):
    M, K, N = mkn
    torch.manual_seed(M * 7 + K + N)
    a = torch.randn(M, K).to(dtype)
    w = torch.randn(K, N).to(dtype)
    w_ref, mq, ms = om.marlin_quantize(w, 4, group_size)
    empty = torch.empty(0, dtype=torch.int32, device=DEV)
    out = ops.gptq_marlin_gemm(a.to(DEV), mq.to(DEV), ms.to(DEV), empty, empty, empty, _workspace(N),
                               _types().uint4b8, M, N, K, True, False, True, False)
    torch.cuda.synchronize()
    assert out.shape == (M, N) and out.dtype == dtype
    _check_gemm(out, om.marlin_gemm(a, w_ref))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("group_size", [32, 128])
@pytest.mark.parametrize("mkn", [(1, 128, 64), (48, 512, 256), (256, 1024, 512), (260, 256, 192)])
def test_awq_marlin_gemm_uint4_zp(ops, dtype, group_size, mkn):
    M, K, N = mkn
    torch.manual_seed(M + K + N)
    a = torch.randn(M, K).to(dtype)
    w = torch.randn(K, N).to(dtype)
    w_ref, mq, ms, mzp = om.awq_marlin_quantize(w, 4, group_size)
    empty = torch.empty(0, dtype=torch.int32, device=DEV)
    out = ops.gptq_marlin_gemm(a.to(DEV), mq.to(DEV), ms.to(DEV), mzp.to(DEV), empty, empty, _workspace(N),
                               _types().uint4, M, N, K, True, True, True, False)
    torch.cuda.synchronize()
    _check_gemm(out, om.marlin_gemm(a, w_ref))


def test_gemm_llama_shapes_split_k(ops, cabi):
    """The four Llama-3-8B projections at M = 256 (BASELINE configs[2]); exercises the split-k plan."""
    torch.manual_seed(0)
    M = 256
    for K, N in ((4096, 6144), (4096, 4096), (14336, 4096)):
        a = (torch.randn(M, K) * 0.5).to(torch.bfloat16)
        w = torch.randn(K, N).to(torch.bfloat16)
        w_ref, mq, ms = om.marlin_quantize(w, 4, 128)
        empty = torch.empty(0, dtype=torch.int32, device=DEV)
        ws = _workspace(N)
        outs = [ops.gptq_marlin_gemm(a.to(DEV), mq.to(DEV), ms.to(DEV), empty, empty, empty, ws,
                                     _types().uint4b8, M, N, K, True, False, True, False) for _ in range(3)]
        torch.cuda.synchronize()
        _check_gemm(outs[0], om.marlin_gemm(a, w_ref))
        assert (ws == 0).all(), "the lock workspace must be returned to zero"
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), "split-k reduce must be deterministic"
        assert cabi.b200_marlin_gemm_plan(M, N, K, K // 128) >= 1


def test_marlin_argument_errors(ops):
    empty = torch.empty(0, dtype=torch.int32, device=DEV)
    a = torch.zeros(4, 128, dtype=torch.bfloat16, device=DEV)
    mq = torch.zeros(8, 128, dtype=torch.int32, device=DEV)
    ms = torch.ones(1, 64, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError, match="b_q_type must be uint4b8 or uint8b128"):
        ops.gptq_marlin_gemm(a, mq, ms, empty, empty, empty, _workspace(64), _types().uint4, 4, 64, 128,
                             True, False, True, False)
    with pytest.raises(RuntimeError, match="Shape mismatch: a.size\\(1\\)"):
        ops.gptq_marlin_gemm(a, mq, ms, empty, empty, empty, _workspace(64), _types().uint4b8, 4, 64, 256,
                             True, False, True, False)
    with pytest.raises(RuntimeError, match="workspace.numel"):
        ops.gptq_marlin_gemm(a, mq, ms, empty, empty, empty, torch.zeros(1, dtype=torch.int32, device=DEV),
                             _types().uint4b8, 4, 64, 128, True, False, True, False)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("group_size", [-1, 32, 128])
@pytest.mark.parametrize("mkn", [(1, 128, 64), (16, 512, 192), (100, 256, 448), (256, 2048, 512), (300, 1024, 256)])
def test_gptq_marlin_gemm_uint8b128(ops, dtype, group_size, mkn):
    M, K, N = mkn
    torch.manual_seed(M * 3 + K + N)
    a = torch.randn(M, K).to(dtype)
    w = torch.randn(K, N).to(dtype)
    w_ref, mq, ms = om.marlin_quantize(w, 8, group_size)
    empty = torch.empty(0, dtype=torch.int32, device=DEV)
    out = ops.gptq_marlin_gemm(a.to(DEV), mq.to(DEV), ms.to(DEV), empty, empty, empty, _workspace(N),
                               _types().uint8b128, M, N, K, True, False, True, False)
    torch.cuda.synchronize()
    assert out.shape == (M, N) and out.dtype == dtype
    _check_gemm(out, om.marlin_gemm(a, w_ref))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("group_size", [-1, 32, 128])
@pytest.mark.parametrize("mkn", [(1, 128, 64), (48, 512, 256), (256, 1024, 512), (260, 256, 192)])
def test_awq_marlin_gemm_uint8_zp(ops, dtype, group_size, mkn):
    M, K, N = mkn
    torch.manual_seed(M + K + N + 1)
    a = torch.randn(M, K).to(dtype)
    w = torch.randn(K, N).to(dtype)
    w_ref, mq, ms, mzp = om.awq_marlin_quantize(w, 8, group_size)
    empty = torch.empty(0, dtype=torch.int32, device=DEV)
    out = ops.gptq_marlin_gemm(a.to(DEV), mq.to(DEV), ms.to(DEV), mzp.to(DEV), empty, empty, _workspace(N),
                               _types().uint8, M, N, K, True, True, True, False)
    torch.cuda.synchronize()
    _check_gemm(out, om.marlin_gemm(a, w_ref))


def test_awq_marlin_gemm_uint4_zp_channelwise(ops):
    M, K, N = 40, 512, 320
    torch.manual_seed(5)
    a = torch.randn(M, K).to(torch.bfloat16)
    w = torch.randn(K, N).to(torch.bfloat16)
    w_ref, mq, ms, mzp = om.awq_marlin_quantize(w, 4, -1)
    empty = torch.empty(0, dtype=torch.int32, device=DEV)
    out = ops.gptq_marlin_gemm(a.to(DEV), mq.to(DEV), ms.to(DEV), mzp.to(DEV), empty, empty, _workspace(N),
                               _types().uint4, M, N, K, True, True, True, False)
    torch.cuda.synchronize()
    _check_gemm(out, om.marlin_gemm(a, w_ref))


@pytest.mark.parametrize("group_size", [64, 128])
@pytest.mark.parametrize("mkn", [(1, 128, 64), (33, 512, 192), (256, 1024, 512)])
def test_hqq_marlin_gemm_float_zero_points(ops, group_size, mkn):
    """is_zp_float = True (HQQ): fp16 zero points laid out like the scales; W = fp16(fp16(q - zp) * s)."""
    M, K, N = mkn
    g = torch.Generator().manual_seed(M + K + N)
    a = torch.randn(M, K, generator=g).half()
    q = torch.randint(0, 16, (K, N), generator=g)
    s = (torch.rand(K // group_size, N, generator=g) * 0.02 + 0.005).half()
    zp = (torch.rand(K // group_size, N, generator=g) * 4 + 6).half()
    w_ref, mq, ms, mz = om.hqq_marlin_quantize(q, s, zp, group_size)
    empty = torch.empty(0, dtype=torch.int32, device=DEV)
    out = ops.gptq_marlin_gemm(a.to(DEV), mq.to(DEV), ms.to(DEV), mz.to(DEV), empty, empty, _workspace(N),
                               _types().uint4, M, N, K, True, True, False, True)
    torch.cuda.synchronize()
    _check_gemm(out, om.marlin_gemm(a, w_ref))
    with pytest.raises(RuntimeError, match="Computation type must be float16"):
        ops.gptq_marlin_gemm(a.to(DEV).bfloat16(), mq.to(DEV), ms.to(DEV).bfloat16(), mz.to(DEV), empty, empty,
                             _workspace(N), _types().uint4, M, N, K, True, True, False, True)


@pytest.mark.parametrize("mkn", [(16, 4096, 6144), (32, 14336, 4096), (5, 1024, 28672)])
def test_small_batch_kernel_stream_k_is_deterministic_and_resets_locks(ops, cabi, mkn):
    """M <= 32 takes the mma.sync streaming kernel: tiles shared by several CTAs are reduced through fp32 slabs and
    a ticket on the lock workspace — the result must not depend on CTA arrival order and the workspace must return
    to zero (the reference's workspace contract, kernels/torch_bindings.cpp:167-176)."""
    M, K, N = mkn
    torch.manual_seed(K + N)
    a = (torch.randn(M, K) * 0.5).to(torch.bfloat16)
    w = torch.randn(K, N).to(torch.bfloat16)
    w_ref, mq, ms = om.marlin_quantize(w, 4, 128)
    empty = torch.empty(0, dtype=torch.int32, device=DEV)
    ws = _workspace(N)
    ad, mqd, msd = a.to(DEV), mq.to(DEV), ms.to(DEV)
    outs = [ops.gptq_marlin_gemm(ad, mqd, msd, empty, empty, empty, ws, _types().uint4b8, M, N, K, True, False, True, False)
            for _ in range(4)]
    torch.cuda.synchronize()
    _check_gemm(outs[0], om.marlin_gemm(a, w_ref))
    assert (ws == 0).all(), "the lock workspace must be returned to zero"
    for o in outs[1:]:
        assert torch.equal(outs[0], o), "slab reduce must be deterministic"
    assert cabi.b200_marlin_gemm_plan(M, N, K, K // 128) >= 1


def test_marlin_ops_are_cuda_graph_capturable(ops):
    """Decode runs under torch.cuda.graph (worker/model_runner.py:1682+): the quantised GEMMs (both kernels) and the
    grouped MoE GEMM must not synchronise or allocate outside torch's graph-aware allocator."""
    from aphrodite_engine_b200 import fused_moe as fm
    torch.manual_seed(1)
    K, N, E, topk = 512, 256, 4, 2
    w = torch.randn(K, N).half()
    w_ref, mq, ms = om.marlin_quantize(w, 4, 128)
    mqd, msd = mq.to(DEV), ms.to(DEV)
    empty = torch.empty(0, dtype=torch.int32, device=DEV)
    ws = _workspace(N)
    a_small = torch.randn(8, K, dtype=torch.float16, device=DEV)
    a_big = torch.randn(200, K, dtype=torch.float16, device=DEV)
    qs, ss, refs = [], [], []
    for e in range(E):
        r, q, s = om.marlin_quantize((torch.randn(K, N) * 0.1).half(), 4, 128)
        refs.append(r); qs.append(q); ss.append(s)
    q3, s3 = torch.stack(qs).to(DEV), torch.stack(ss).to(DEV)
    ids = torch.randint(0, E, (8, topk), dtype=torch.int32)
    ids[:, 1] = (ids[:, 0] + 1) % E
    tw = torch.rand(8, topk)
    idsd, twd = ids.to(DEV), tw.to(DEV)
    none = torch.empty(E, 0, dtype=torch.int32, device=DEV)
    sorted_ids, _, _ = fm.moe_align_block_size(idsd, 16, E)

    def run():
        o1 = ops.gptq_marlin_gemm(a_small, mqd, msd, empty, empty, empty, ws, _types().uint4b8, 8, N, K, True, False, True, False)
        o2 = ops.gptq_marlin_gemm(a_big, mqd, msd, empty, empty, empty, ws, _types().uint4b8, 200, N, K, True, False, True, False)
        o3 = torch.ops._moe_C.marlin_gemm_moe(a_small, q3, sorted_ids, twd, idsd, s3, none, none, ws, 8, N, K, True, E, topk,
                                              16, True, True)
        return o1, o2, o3

    eager = [t.clone() for t in run()]
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs = run()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    for e, o in zip(eager, outs):
        assert torch.equal(e, o)
    _check_gemm(outs[0], om.marlin_gemm(a_small.cpu(), w_ref))
    _check_gemm(outs[2], om.marlin_gemm_moe(a_small.cpu(), refs, ids, tw, True, True))
