"""Debug: in-kernel cycle attribution of the marlin GEMM roles (CTA 0), B200_MARLIN_DEBUG=16."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B200_MARLIN_DEBUG"] = str(16 | int(os.environ.get("EXTRA_DEBUG", "0")))
import aphrodite_engine_b200._custom_ops as ops
from aphrodite_engine_b200 import _native
from aphrodite_engine_b200.scalar_type import scalar_types
lib = _native.load_c_abi()
dev = "cuda:0"
for M, K, N in ((16, 4096, 28672), (256, 4096, 28672), (16, 14336, 4096), (256, 14336, 4096)):
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    q = torch.randint(-2**31, 2**31 - 1, (K // 16, N * 2), device=dev, dtype=torch.int32)
    s = (torch.rand(K // 128, N, device=dev) * 0.01).to(torch.bfloat16)
    e = torch.empty(0, dtype=torch.int32, device=dev)
    ws = torch.zeros((N // 64) * 16, dtype=torch.int32, device=dev)
    for _ in range(3):
        ops.gptq_marlin_gemm(x, q, s, e, e, e, ws, scalar_types.uint4b8, M, N, K, True, False, True, False)
    buf = (ctypes.c_ulonglong * 32)()
    lib.b200_debug_marlin_prof(ctypes.cast(buf, ctypes.c_void_p))
    v = list(buf)
    n = max(v[2], 1)
    print(f"M={M} K={K} N={N} chunks={v[2]}  per-chunk cycles:")
    print(f"  act producer : total {v[0]/n:7.0f}  wait_empty {v[1]/n:7.0f}")
    print(f"  mma issuer   : total {v[4]/n:7.0f}  wait_full_act {v[5]/n:7.0f}  wait_full_w {v[6]/n:7.0f}  issue4mma {v[7]/n:7.0f}  commit {v[16]/n:7.0f}")
    print(f"  pk producer  : total {v[8]/n:7.0f}  wait_pk_empty {v[9]/n:7.0f}")
    print(f"  dequant w3   : total {v[12]/n:7.0f}  wait_pk_full {v[13]/n:7.0f}  wait_empty {v[14]/n:7.0f}   (+epilogue: {v[15]/n:7.0f})")
