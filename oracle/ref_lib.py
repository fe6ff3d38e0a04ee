"""Loader for the reference's own CPU kernels (oracle/_ref/_ref_cpu_C_<isa>.so) — TEST INFRASTRUCTURE.

The .so files are built by oracle/build_ref.py straight from /root/reference/kernels/cpu/*.cpp and
register the reference's op names under torch.ops._ref_cpu_C / torch.ops._ref_cpu_C_cache_ops
(CPU dispatch key). Limits of that code (not of this loader): fp32/bf16 only
(kernels/cpu/cpu_types_x86.hpp:14-17), block_size 16 only (kernels/cpu/attention.cpp:410-418),
kv_cache_dtype "auto" with k_scale == v_scale == 1 (:430), no head size 120 (:377-402).
"""
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_loaded = None


def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def available_isa():
    flags = _cpu_flags()
    need = {"avx512f", "avx512vl", "avx512bw", "avx512dq"}
    if not need <= flags:
        return None
    return "avx512bf16" if "avx512_bf16" in flags else "avx512"


def load():
    """Returns (ops, cache_ops, isa) or None when the host cannot run / does not have the build."""
    global _loaded
    if _loaded is not None:
        return _loaded or None
    isa = available_isa()
    path = os.path.join(_HERE, "_ref", f"_ref_cpu_C_{isa}.so") if isa else None
    if not path or not os.path.exists(path):
        _loaded = False
        return None
    torch.ops.load_library(path)
    _loaded = (torch.ops._ref_cpu_C, torch.ops._ref_cpu_C_cache_ops, isa)
    return _loaded


def paged_attention_v1(out, query, key_cache, value_cache, num_kv_heads, scale, block_tables,
                       seq_lens, block_size, max_seq_len, alibi_slopes=None):
    ops, _, _ = load()
    ops.paged_attention_v1(out, query, key_cache, value_cache, num_kv_heads, scale, block_tables,
                           seq_lens, block_size, max_seq_len, alibi_slopes, "auto", 1.0, 1.0, 0, 0,
                           0, 64, 0)


def paged_attention_v2(out, exp_sums, max_logits, tmp_out, query, key_cache, value_cache,
                       num_kv_heads, scale, block_tables, seq_lens, block_size, max_seq_len,
                       alibi_slopes=None):
    ops, _, _ = load()
    ops.paged_attention_v2(out, exp_sums, max_logits, tmp_out, query, key_cache, value_cache,
                           num_kv_heads, scale, block_tables, seq_lens, block_size, max_seq_len,
                           alibi_slopes, "auto", 1.0, 1.0, 0, 0, 0, 64, 0)
