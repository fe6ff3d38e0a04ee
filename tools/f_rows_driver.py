#!/usr/bin/env python3
"""Launches one SURVEY §8(f) kernel a few times at its bench shape, for `ncu -k regex:<kernel>` captures
(tools/gpu_check_f_rows.sh ncu). Usage: python tools/f_rows_driver.py mm|prefill|sampling|quant"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aphrodite_engine_b200._custom_ops as ops  # noqa: E402

dev = "cuda:0"
which = sys.argv[1]
g = torch.Generator(device=dev).manual_seed(0)
if which == "mm":
    M, K, N = 256, 4096, 28672
    a8, sa = ops.scaled_fp8_quant(torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16), use_per_token_if_dynamic=True)
    w8, sw = ops.scaled_fp8_quant((torch.randn(N, K, device=dev, generator=g) * 0.05).to(torch.bfloat16), use_per_token_if_dynamic=True)
    for _ in range(6):
        ops.cutlass_scaled_mm(a8, w8.t(), sa, sw, torch.bfloat16)
elif which == "quant":
    x = torch.randn(8192, 4096, device=dev, generator=g).to(torch.bfloat16)
    for _ in range(3):
        ops.scaled_fp8_quant(x, use_per_token_if_dynamic=True)
elif which == "sampling":
    B, V = 256, 128256
    p = torch.softmax(torch.randn(B, V, device=dev, generator=g) * 3, -1)
    u = torch.rand(32, B, device=dev, generator=g)
    for _ in range(3):
        ops.top_k_sampling_from_probs(p, u, None, 50)
elif which == "prefill":
    from aphrodite_engine_b200.attention.prefix_prefill import context_attention_fwd
    dt, Hq, Hkv, D, BS, B, ctx, ql = torch.bfloat16, 32, 8, 128, 16, 4, 2048, 2048
    T = B * ql
    q = (torch.randn(T, Hq, D, device=dev, generator=g) * 0.5).to(dt)
    k = (torch.randn(T, Hkv, D, device=dev, generator=g) * 0.5).to(dt)
    v = (torch.randn(T, Hkv, D, device=dev, generator=g) * 0.5).to(dt)
    nblk = ctx // BS + 1
    kc = (torch.randn(B * nblk, Hkv, D // 8, BS, 8, device=dev, generator=g) * 0.5).to(dt)
    vc = (torch.randn(B * nblk, Hkv, D, BS, device=dev, generator=g) * 0.5).to(dt)
    bt = torch.randperm(B * nblk, device=dev, generator=g).view(B, nblk).to(torch.int32)
    start = (torch.arange(B, device=dev) * ql).to(torch.int32)
    seq = torch.full((B,), ctx + ql, dtype=torch.int32, device=dev)
    cl = torch.full((B,), ctx, dtype=torch.int32, device=dev)
    o = torch.empty_like(q)
    for _ in range(3):
        context_attention_fwd(q, k, v, o, "auto", kc, vc, bt, start, seq, cl, ql)
torch.cuda.synchronize()
print("done", which)
