#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the REFERENCE'S OWN CPU kernels (oracle/_ref, compiled from
/root/reference/kernels/cpu/*.cpp by oracle/build_ref.py). Run in the build container only:

    python tests/golden/make_golden.py

Each file stores the seeded inputs and the reference outputs; bf16 tensors are stored as their
uint16 bit patterns (key suffix `__bf16`). The reference ships no golden vectors of its own for this
path (its kernel tests are generative), so these are "outputs of the reference itself
run here". Limits of the reference CPU kernels: fp32/bf16, block_size 16, kv_cache_dtype auto.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_lib, paged_ops as po  # noqa: E402


def pack(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            if v.dtype == torch.bfloat16:
                out[k + "__bf16"] = v.contiguous().view(torch.int16).numpy().view(np.uint16)
            else:
                out[k] = v.contiguous().numpy()
        else:
            out[k] = np.asarray(v)
    return out


def save(name, d):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **pack(d))
    print("wrote", name)


def main():
    r = ref_lib.load()
    assert r is not None, "oracle/_ref is not built or this CPU lacks AVX-512"
    rops, rcache, isa = r
    print("reference CPU kernels:", isa)
    for dt, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        # ---- paged attention v1 / v2 (GQA 4, ragged lengths, multi-partition, ALiBi) ----
        torch.manual_seed(0)
        S, Hq, Hkv, D, BS, NB = 5, 8, 2, 64, 16, 80
        scale = D ** -0.5
        seq_lens = [1, 16, 17, 300, 1100]
        kc, vc = po.make_kv_cache(NB, BS, Hkv, D, dt, "auto", 3)
        q = torch.empty(S, Hq, D).uniform_(-scale, scale).to(dt)
        nbp = (max(seq_lens) + BS - 1) // BS
        bt = torch.randint(0, NB, (S, nbp), dtype=torch.int32)
        sl = torch.tensor(seq_lens, dtype=torch.int32)
        alibi = torch.randn(Hq, dtype=torch.float32)
        d = dict(q=q, key_cache=kc, value_cache=vc, block_tables=bt, seq_lens=sl, alibi=alibi,
                 scale=scale, num_kv_heads=Hkv, block_size=BS)
        for al, atag in ((None, ""), (alibi, "_alibi")):
            out = torch.empty_like(q)
            ref_lib.paged_attention_v1(out, q, kc, vc, Hkv, scale, bt, sl, BS, max(seq_lens), al)
            d["out_v1" + atag] = out
            P = (max(seq_lens) + 511) // 512
            out2 = torch.empty_like(q)
            es, ml = torch.zeros(S, Hq, P), torch.zeros(S, Hq, P)
            tmp = torch.zeros(S, Hq, P, D, dtype=dt)
            ref_lib.paged_attention_v2(out2, es, ml, tmp, q, kc, vc, Hkv, scale, bt, sl, BS,
                                       max(seq_lens), al)
            d["out_v2" + atag] = out2
        save(f"paged_attention_{tag}", d)

        # ---- reshape_and_cache + copy_blocks (exact) ----
        torch.manual_seed(1)
        T, H, D, BS, NB = 21, 4, 64, 16, 6
        key, value = torch.randn(T, H, D).to(dt), torch.randn(T, H, D).to(dt)
        kc, vc = po.make_kv_cache(NB, BS, H, D, dt, "auto", 4)
        slots = torch.randperm(NB * BS)[:T]
        kc0, vc0 = kc.clone(), vc.clone()
        rcache.reshape_and_cache(key, value, kc, vc, slots, "auto", 1.0, 1.0)
        bm = torch.tensor([[0, 4], [1, 5], [2, 3]], dtype=torch.long)
        kc2, vc2 = kc.clone(), vc.clone()
        rcache.copy_blocks([kc2], [vc2], bm)
        save(f"cache_ops_{tag}", dict(key=key, value=value, key_cache_in=kc0, value_cache_in=vc0,
                                      slot_mapping=slots, key_cache_out=kc, value_cache_out=vc,
                                      block_mapping=bm, key_cache_copied=kc2, value_cache_copied=vc2))

        # ---- rms_norm / fused_add_rms_norm ----
        torch.manual_seed(2)
        x, res = torch.randn(9, 1024).to(dt), torch.randn(9, 1024).to(dt)
        w = torch.empty(1024).normal_(1.0, 0.1).to(dt)
        out = torch.empty_like(x)
        rops.rms_norm(out, x, w, 1e-5)
        x2, r2 = x.clone(), res.clone()
        rops.fused_add_rms_norm(x2, r2, w, 1e-5)
        save(f"rms_norm_{tag}", dict(x=x, residual=res, weight=w, eps=1e-5, out=out, fused_out=x2,
                                     fused_residual=r2))

        # ---- rotary_embedding (neox / gpt-j) ----
        torch.manual_seed(3)
        Hq, Hkv, D, T, maxp = 8, 2, 64, 13, 512
        inv = 1.0 / (10000 ** (torch.arange(0, D, 2).float() / D))
        fr = torch.einsum("i,j->ij", torch.arange(maxp).float(), inv)
        cache = torch.cat((fr.cos(), fr.sin()), dim=-1).to(dt)
        pos = torch.randint(0, maxp, (T,))
        q, k = torch.randn(T, Hq * D).to(dt), torch.randn(T, Hkv * D).to(dt)
        d = dict(positions=pos, q=q, k=k, cos_sin_cache=cache, head_size=D)
        for neox in (True, False):
            q2, k2 = q.clone(), k.clone()
            rops.rotary_embedding(pos, q2, k2, D, cache, neox)
            d["q_neox" if neox else "q_gptj"] = q2
            d["k_neox" if neox else "k_gptj"] = k2
        save(f"rotary_{tag}", d)

        # ---- activations ----
        torch.manual_seed(4)
        x = torch.randn(11, 2 * 256).to(dt)
        d = dict(x=x)
        for name in ("silu_and_mul", "gelu_and_mul", "gelu_tanh_and_mul"):
            o = torch.empty(11, 256, dtype=dt)
            getattr(rops, name)(o, x)
            d[name] = o
        for name in ("gelu_new", "gelu_fast", "gelu_quick"):
            o = torch.empty_like(x)
            getattr(rops, name)(o, x)
            d[name] = o
        save(f"activations_{tag}", d)


if __name__ == "__main__":
    main()
