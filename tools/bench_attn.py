"""Micro-benchmark of paged_attention_v1/v2 at BASELINE config 2's per-layer shape (device-resident
inputs, CUDA events, distinct blocks per sequence so the 4.3 GB working set cannot sit in L2)."""
import argparse
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aphrodite_engine_b200._custom_ops as ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bs", type=int, default=256)
    ap.add_argument("--ctx", type=int, default=4096)
    ap.add_argument("--heads", type=int, default=32)
    ap.add_argument("--kv-heads", type=int, default=8)
    ap.add_argument("--head-size", type=int, default=128)
    ap.add_argument("--block-size", type=int, default=16)
    ap.add_argument("--kv-dtype", default="auto")
    ap.add_argument("--version", default="v1")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--layers", type=int, default=2, help="distinct KV buffers rotated between iterations")
    a = ap.parse_args()
    dev = "cuda:0"
    S, H, KV, D, BS, CTX = a.bs, a.heads, a.kv_heads, a.head_size, a.block_size, a.ctx
    nb_per = (CTX + BS - 1) // BS
    NB = S * nb_per
    scale = D ** -0.5
    esz = 2 if a.kv_dtype == "auto" else 1
    x = 16 // esz
    caches = []
    for l in range(a.layers):
        if esz == 2:
            kv = torch.empty(2, NB, BS * KV * D, dtype=torch.bfloat16, device=dev).uniform_(-scale, scale)
        else:
            kv = torch.randint(0, 120, (2, NB, BS * KV * D), dtype=torch.uint8, device=dev)
        caches.append((kv[0].view(NB, KV, D // x, BS, x), kv[1].view(NB, KV, D, BS)))
    q = torch.empty(S, H, D, dtype=torch.bfloat16, device=dev).uniform_(-scale, scale)
    bt = torch.randperm(NB, device=dev).view(S, nb_per).to(torch.int32)
    sl = torch.full((S,), CTX, dtype=torch.int32, device=dev)
    out = torch.empty_like(q)
    P = (CTX + 511) // 512
    tmp = torch.empty(S, H, P, D, dtype=q.dtype, device=dev)
    es = torch.empty(S, H, P, dtype=torch.float32, device=dev)
    ml = torch.empty_like(es)

    def call(i):
        kc, vc = caches[i % a.layers]
        if a.version == "v1":
            ops.paged_attention_v1(out, q, kc, vc, KV, scale, bt, sl, BS, CTX, None, a.kv_dtype, 1.0, 1.0)
        else:
            ops.paged_attention_v2(out, es, ml, tmp, q, kc, vc, KV, scale, bt, sl, BS, CTX, None, a.kv_dtype, 1.0, 1.0)

    for i in range(3):
        call(i)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.iters)]
    for i, (s, e) in enumerate(evs):
        s.record()
        call(i)
        e.record()
    torch.cuda.synchronize()
    ms = sorted(s.elapsed_time(e) for s, e in evs)
    med = ms[len(ms) // 2]
    algo = S * CTX * 2 * KV * D * esz + 2 * S * H * D * 2 + S * nb_per * 4
    print(json.dumps({"version": a.version, "kv_dtype": a.kv_dtype, "bs": S, "ctx": CTX,
                      "ms_median": med, "ms_min": ms[0], "algo_bytes": algo,
                      "GBps_median": algo / med / 1e6, "GBps_best": algo / ms[0] / 1e6}))


if __name__ == "__main__":
    main()
