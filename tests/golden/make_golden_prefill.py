#!/usr/bin/env python3
"""Generates tests/golden/prefill_*.npz from the REFERENCE'S OWN prefill kernel: the Triton `context_attention_fwd`
of /root/reference/aphrodite/attention/ops/prefix_prefill.py, executed by Triton's CPU interpreter
(TRITON_INTERPRET=1) on CPU tensors — no GPU, no modification of the reference: the module is loaded from its file
with a two-line stand-in for `aphrodite.platforms.current_platform` (the package itself does not import here,
SURVEY.md §8c). Run in the build container only:

    python tests/golden/make_golden_prefill.py

Each file stores the seeded inputs and the reference's output. The cases span what the reference's own test sweeps
(tests/kernels/test_prefix_prefill.py: head sizes, queries-per-kv, sliding windows, kv-cache dtypes, ALiBi) at sizes the
interpreter finishes in seconds. fp8 caches are stored as their uint8 bit patterns. fp16 only: the interpreter computes
through numpy, which has no bfloat16 (a bf16 run returns garbage), so bf16 is held to the restatement that these
vectors pin.
"""
import importlib.util
import os
import sys
import types

os.environ["TRITON_INTERPRET"] = "1"

import numpy as np  # noqa: E402
import torch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_FILE = "/root/reference/aphrodite/attention/ops/prefix_prefill.py"


def load_reference():
    ap = types.ModuleType("aphrodite")
    ap.__path__ = []
    pl = types.ModuleType("aphrodite.platforms")

    class _Platform:
        @staticmethod
        def get_device_capability():
            return (8, 0)

    pl.current_platform = _Platform()
    sys.modules.setdefault("aphrodite", ap)
    sys.modules.setdefault("aphrodite.platforms", pl)
    spec = importlib.util.spec_from_file_location("ref_prefix_prefill", REF_FILE)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def pack(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            if v.dtype == torch.bfloat16:
                out[k + "__bf16"] = v.contiguous().view(torch.int16).numpy().view(np.uint16)
            else:
                out[k] = v.contiguous().numpy()
        elif v is None:
            continue
        else:
            out[k] = np.asarray(v)
    return out


CASES = [
    # name, dtype, Hq, Hkv, D, BS, ctx lens, query lens, kv dtype, sliding window, alibi
    ("f16_d128_gqa4", torch.float16, 8, 2, 128, 16, [37, 0, 64], [70, 33, 5], "auto", 0, False),
    ("f16_d64_mha_sw16", torch.float16, 4, 4, 64, 32, [50, 9], [40, 66], "auto", 16, False),
    ("f16_d96_gqa2_sw64", torch.float16, 4, 2, 96, 16, [100, 3], [17, 80], "auto", 64, False),
    ("f16_d128_gqa4_b", torch.float16, 8, 2, 128, 16, [129, 16], [64, 65], "auto", 0, False),
    ("f16_d128_fp8", torch.float16, 4, 1, 128, 16, [45, 80], [35, 20], "fp8", 0, False),
    ("f16_d64_fp8_e5m2", torch.float16, 4, 2, 64, 16, [33, 0], [31, 64], "fp8_e5m2", 0, False),
    ("f16_d128_alibi", torch.float16, 4, 2, 128, 16, [70, 20], [40, 90], "auto", 0, True),
    ("f16_d256_mqa", torch.float16, 4, 1, 256, 16, [20], [70], "auto", 0, False),
]


def make_case(ref, name, dt, Hq, Hkv, D, BS, ctxs, qls, kvd, sw, alibi, seed):
    g = torch.Generator().manual_seed(seed)
    B = len(ctxs)
    T = sum(qls)
    q = (torch.randn(T, Hq, D, generator=g) * 0.5).to(dt)
    k = (torch.randn(T, Hkv, D, generator=g) * 0.5).to(dt)
    v = (torch.randn(T, Hkv, D, generator=g) * 0.5).to(dt)
    max_blocks = max((c + BS - 1) // BS for c in ctxs) + 1
    NB = B * max_blocks + 3
    perm = torch.randperm(NB, generator=g)
    bt = perm[: B * max_blocks].view(B, max_blocks).to(torch.int32)
    x = 8
    kc = (torch.randn(NB, Hkv, D // x, BS, x, generator=g) * 0.5)
    vc = (torch.randn(NB, Hkv, D, BS, generator=g) * 0.5)
    k_scale, v_scale = (1.0, 1.0) if kvd == "auto" else (0.75, 1.5)
    if kvd == "auto":
        kc, vc = kc.to(dt), vc.to(dt)
    else:
        f8 = torch.float8_e4m3fn if kvd == "fp8" else torch.float8_e5m2
        kc = (kc / k_scale).to(f8).view(torch.uint8)
        vc = (vc / v_scale).to(f8).view(torch.uint8)
    start = torch.tensor([sum(qls[:i]) for i in range(B)], dtype=torch.int32)
    seq = torch.tensor([c + q_ for c, q_ in zip(ctxs, qls)], dtype=torch.int32)
    ctx = torch.tensor(ctxs, dtype=torch.int32)
    slopes = (torch.rand(Hq, generator=g) * 0.2 + 0.01).float() if alibi else None
    o = torch.zeros_like(q)
    ref.context_attention_fwd(q, k, v, o, kvd, kc, vc, bt, start, seq, ctx, max(qls), k_scale, v_scale, slopes,
                              sw if sw > 0 else None)
    d = dict(q=q, k=k, v=v, key_cache=kc, value_cache=vc, block_tables=bt, start_loc=start, seq_lens=seq, ctx_lens=ctx,
             out=o, sliding_window=sw, k_scale=k_scale, v_scale=v_scale, kv_cache_dtype=kvd, alibi_slopes=slopes,
             max_query_len=max(qls))
    np.savez_compressed(os.path.join(HERE, "prefill_" + name + ".npz"), **pack(d))
    print("wrote prefill_" + name, "| mean |out| =", float(o.float().abs().mean()))


if __name__ == "__main__":
    ref = load_reference()
    for i, c in enumerate(CASES):
        make_case(ref, *c, seed=100 + i)
