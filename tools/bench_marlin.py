"""Micro-benchmark of the tcgen05 W4A16 Marlin-format GEMM at the Llama-3-8B projection shapes
(BASELINE configs[2]: M = 256), with cuBLAS bf16 (F.linear) of the same shape beside it."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aphrodite_engine_b200._custom_ops as ops  # noqa: E402
from aphrodite_engine_b200.scalar_type import scalar_types  # noqa: E402


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()                       # > L2: evict weights between timed launches
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, nargs="+", default=[256])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--group", type=int, default=128)
    ap.add_argument("--kn", type=int, nargs=2, action="append", help="K N pair(s) instead of the four Llama-3-8B shapes")
    ap.add_argument("--sustained", action="store_true",
                    help="also time 200 back-to-back launches rotating over weight copies larger than L2 (boosted clocks, no idle gaps)")
    a = ap.parse_args()
    dev = "cuda:0"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shapes = [tuple(kn) for kn in a.kn] if a.kn else [(4096, 6144), (4096, 4096), (4096, 28672), (14336, 4096)]
    for M in a.m:
        for K, N in shapes:
            x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
            q = torch.randint(-2**31, 2**31 - 1, (K // 16, N * 2), device=dev, dtype=torch.int32)
            groups = K // a.group if a.group > 0 else 1
            s = (torch.rand(groups, N, device=dev) * 0.01).to(torch.bfloat16)
            empty = torch.empty(0, dtype=torch.int32, device=dev)
            ws = torch.zeros((N // 64) * 16, dtype=torch.int32, device=dev)
            w16 = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
            t_q = timeit(lambda: ops.gptq_marlin_gemm(x, q, s, empty, empty, empty, ws, scalar_types.uint4b8,
                                                      M, N, K, True, False, True, False), a.iters, flush)
            t_b = timeit(lambda: torch.nn.functional.linear(x, w16), a.iters, flush)
            fl = 2.0 * M * K * N
            wbytes = K * N / 2 + groups * N * 2
            sus = None
            if a.sustained:
                nrot = max(2, int(300e6 // (K * N / 2)) + 1)
                qs = [q.clone() for _ in range(nrot)]
                for i in range(20):
                    ops.gptq_marlin_gemm(x, qs[i % nrot], s, empty, empty, empty, ws, scalar_types.uint4b8, M, N, K, True, False, True, False)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(200):
                    ops.gptq_marlin_gemm(x, qs[i % nrot], s, empty, empty, empty, ws, scalar_types.uint4b8, M, N, K, True, False, True, False)
                e1.record()
                torch.cuda.synchronize()
                sus = e0.elapsed_time(e1) / 200
                del qs
            print(json.dumps({"M": M, "K": K, "N": N, "group": a.group, "w4a16_ms": t_q, "w4a16_tflops": fl / t_q / 1e9,
                              "w4a16_weight_GBps": wbytes / t_q / 1e6, "cublas_bf16_ms": t_b,
                              "cublas_bf16_tflops": fl / t_b / 1e9,
                              **({"w4a16_sustained_ms": sus, "w4a16_sustained_tflops": fl / sus / 1e9,
                                  "w4a16_sustained_weight_GBps": wbytes / sus / 1e6} if sus else {})}))


if __name__ == "__main__":
    main()
