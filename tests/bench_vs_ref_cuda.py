"""Per-kernel timing of this repo's ops beside the reference's OWN CUDA kernels recompiled for sm_100a
(oracle/_ref/_ref_cuda_C.so, see oracle/build_ref_cuda.py) — the "reference kernel recompiled" column of
SURVEY §8(d). Same device tensors, same call arguments, L2 flushed between timed launches, CUDA events,
median of --iters launches. One JSON line per op; never a bench.py value.

TEST INFRASTRUCTURE (lives under tests/ because it loads oracle/_ref): run as `python tests/bench_vs_ref_cuda.py`.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aphrodite_engine_b200._custom_ops as ops  # noqa: E402
from aphrodite_engine_b200 import fused_moe as fm  # noqa: E402
from aphrodite_engine_b200.scalar_type import scalar_types  # noqa: E402

DEV = "cuda:0"


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def emit(name, shape, t_mine, t_ref, bytes_=None, flops=None):
    rec = {"op": name, "shape": shape, "b200_ms": round(t_mine, 5), "ref_sm100a_ms": round(t_ref, 5),
           "speedup": round(t_ref / t_mine, 2)}
    if bytes_:
        rec["b200_GBps"] = round(bytes_ / t_mine / 1e6, 1)
        rec["ref_GBps"] = round(bytes_ / t_ref / 1e6, 1)
    if flops:
        rec["b200_TFLOPs"] = round(flops / t_mine / 1e9, 1)
        rec["ref_TFLOPs"] = round(flops / t_ref / 1e9, 1)
    print(json.dumps(rec), flush=True)


def bench_attention(ref, it, flush, S, H, KV, D, CTX, kv_dtype):
    BS = 16
    nb_per = CTX // BS
    NB = S * nb_per
    scale = D ** -0.5
    esz = 2 if kv_dtype == "auto" else 1
    x = 16 // esz
    if esz == 2:
        kv = torch.empty(2, NB, BS * KV * D, dtype=torch.bfloat16, device=DEV).uniform_(-scale, scale)
    else:
        kv = torch.randint(0, 120, (2, NB, BS * KV * D), dtype=torch.uint8, device=DEV)
    kc, vc = kv[0].view(NB, KV, D // x, BS, x), kv[1].view(NB, KV, D, BS)
    q = torch.empty(S, H, D, dtype=torch.bfloat16, device=DEV).uniform_(-scale, scale)
    bt = torch.randperm(NB, device=DEV).view(S, nb_per).to(torch.int32)
    sl = torch.full((S,), CTX, dtype=torch.int32, device=DEV)
    out = torch.empty_like(q)
    args = (out, q, kc, vc, KV, scale, bt, sl, BS, CTX, None, kv_dtype, 1.0, 1.0, 0, 0, 0, 64, 0)
    t1 = timeit(lambda: ops.paged_attention_v1(*args), it, flush)
    t2 = timeit(lambda: ref.paged_attention_v1(*args), max(3, it // 4), flush)
    algo = S * CTX * 2 * KV * D * esz + 2 * S * H * D * 2 + S * nb_per * 4
    emit("paged_attention_v1", f"seqs={S} ctx={CTX} Hq={H} Hkv={KV} D={D} kv={kv_dtype}", t1, t2, bytes_=algo)
    del kv


def bench_small_ops(ref, it, flush):
    T, Hd, Hq, Hkv, D, inter = 256, 4096, 32, 8, 128, 14336
    dt = torch.bfloat16
    x = torch.randn(T, Hd, dtype=dt, device=DEV)
    w = torch.ones(Hd, dtype=dt, device=DEV)
    o = torch.empty_like(x)
    emit("rms_norm", f"[{T},{Hd}] bf16", timeit(lambda: ops.rms_norm(o, x, w, 1e-5), it, flush),
         timeit(lambda: ref.rms_norm(o, x, w, 1e-5), it, flush), bytes_=T * Hd * 2 * 2)
    r = torch.randn_like(x)
    emit("fused_add_rms_norm", f"[{T},{Hd}] bf16", timeit(lambda: ops.fused_add_rms_norm(x, r, w, 1e-5), it, flush),
         timeit(lambda: ref.fused_add_rms_norm(x, r, w, 1e-5), it, flush), bytes_=T * Hd * 2 * 4)
    qkv = torch.randn(T, (Hq + 2 * Hkv) * D, dtype=dt, device=DEV)
    qv, kv_ = qkv[:, : Hq * D], qkv[:, Hq * D: (Hq + Hkv) * D]
    pos = torch.randint(0, 4096, (T,), dtype=torch.int64, device=DEV)
    cache = torch.randn(8192, D, dtype=dt, device=DEV)
    emit("rotary_embedding", f"T={T} Hq={Hq} Hkv={Hkv} D={D}",
         timeit(lambda: ops.rotary_embedding(pos, qv, kv_, D, cache, True), it, flush),
         timeit(lambda: ref.rotary_embedding(pos, qv, kv_, D, cache, True), it, flush), bytes_=T * (Hq + Hkv) * D * 2 * 2)
    gu = torch.randn(T, 2 * inter, dtype=dt, device=DEV)
    ao = torch.empty(T, inter, dtype=dt, device=DEV)
    emit("silu_and_mul", f"[{T},{2 * inter}] bf16", timeit(lambda: ops.silu_and_mul(ao, gu), it, flush),
         timeit(lambda: ref.silu_and_mul(ao, gu), it, flush), bytes_=T * inter * 2 * 3)
    NB, BS = 4096, 16
    kc = torch.zeros(NB, Hkv, D // 8, BS, 8, dtype=dt, device=DEV)
    vc = torch.zeros(NB, Hkv, D, BS, dtype=dt, device=DEV)
    key, val = qkv[:, Hq * D: (Hq + Hkv) * D].view(T, Hkv, D), qkv[:, (Hq + Hkv) * D:].view(T, Hkv, D)
    slots = torch.randperm(NB * BS, device=DEV)[:T].to(torch.int64)
    emit("reshape_and_cache", f"T={T} Hkv={Hkv} D={D}",
         timeit(lambda: ops.reshape_and_cache(key, val, kc, vc, slots, "auto", 1.0, 1.0), it, flush),
         timeit(lambda: ref.reshape_and_cache(key, val, kc, vc, slots, "auto", 1.0, 1.0), it, flush),
         bytes_=T * Hkv * D * 2 * 4)


def bench_marlin(ref, it, flush, Ms):
    st = scalar_types.uint4b8
    empty = torch.empty(0, dtype=torch.int32, device=DEV)
    for M in Ms:
        for K, N in ((4096, 6144), (4096, 4096), (4096, 28672), (14336, 4096)):
            x = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
            q = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 16, N * 2), device=DEV, dtype=torch.int32)
            s = (torch.rand(K // 128, N, device=DEV) * 0.01).to(torch.bfloat16)
            ws = torch.zeros((N // 64) * 16, dtype=torch.int32, device=DEV)
            t1 = timeit(lambda: ops.gptq_marlin_gemm(x, q, s, empty, empty, empty, ws, st, M, N, K, True, False, True,
                                                     False), it, flush)
            t2 = timeit(lambda: ref.gptq_marlin_gemm(x, q, s, empty, empty, empty, ws, st.exponent, st.mantissa,
                                                     st.bias, st.signed, M, N, K, True, False, True, False), it, flush)
            emit("gptq_marlin_gemm", f"M={M} K={K} N={N} g=128 u4b8 bf16", t1, t2, bytes_=K * N / 2 + K // 128 * N * 2,
                 flops=2.0 * M * K * N)


def bench_moe(ref, it, flush):
    # BASELINE configs[4]-like: Mixtral expert shapes, T = 128 tokens, 8 experts, top-2 (reference kernel: fp16 only)
    T, E, topk = 128, 8, 2
    for K, N in ((4096, 28672), (14336, 4096)):
        gate = torch.randn(T, E, dtype=torch.float32, device=DEV)
        tw = torch.empty(T, topk, dtype=torch.float32, device=DEV)
        ids = torch.empty(T, topk, dtype=torch.int32, device=DEV)
        src = torch.empty(T, topk, dtype=torch.int32, device=DEV)
        ops.topk_softmax(tw, ids, src, gate)
        block = fm.marlin_moe_block_size(T, E)
        sorted_ids, _, _ = fm.moe_align_block_size(ids, block, E)
        a = torch.randn(T, K, dtype=torch.float16, device=DEV)
        q = torch.randint(-2 ** 31, 2 ** 31 - 1, (E, K // 16, N * 2), device=DEV, dtype=torch.int32)
        s = (torch.rand(E, K // 128, N, device=DEV) * 0.01).half()
        none = torch.empty(E, 0, dtype=torch.int32, device=DEV)
        ws = torch.zeros(((T + 255) // 256) * (N // 64) * 16, dtype=torch.int32, device=DEV)
        call = lambda fn: fn(a, q, sorted_ids, tw, ids, s, none, none, ws, T, N, K, True, E, topk, block, True, False)
        t1 = timeit(lambda: call(torch.ops._moe_C.marlin_gemm_moe), it, flush)
        t2 = timeit(lambda: call(ref.marlin_gemm_moe), it, flush)
        emit("marlin_gemm_moe", f"T={T} E={E} top{topk} K={K} N={N} g=128 fp16", t1, t2,
             bytes_=E * (K * N / 2 + K // 128 * N * 2), flops=2.0 * T * topk * K * N)
        del q, s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--only", default="")
    ap.add_argument("--ms", type=int, nargs="+", default=[256, 16], help="token counts of the marlin GEMM legs")
    a = ap.parse_args()
    so = os.path.join(ROOT, "oracle", "_ref", "_ref_cuda_C.so")
    if not os.path.exists(so):
        print(json.dumps({"unavailable": "oracle/_ref/_ref_cuda_C.so not built"}))
        return
    torch.ops.load_library(so)
    ref = torch.ops._ref_cuda_C
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    want = lambda k: not a.only or k in a.only.split(",")
    if want("small"):
        bench_small_ops(ref, a.iters, flush)
    if want("marlin"):
        bench_marlin(ref, a.iters, flush, a.ms)
    if want("moe"):
        bench_moe(ref, a.iters, flush)
    if want("attn"):
        bench_attention(ref, a.iters, flush, 256, 32, 8, 128, 4096, "auto")
        bench_attention(ref, a.iters, flush, 1024, 4, 1, 128, 8192, "fp8")


if __name__ == "__main__":
    main()
