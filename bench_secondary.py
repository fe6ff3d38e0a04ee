"""Secondary legs of bench.py: BASELINE.json configs[2..4] as bounded measurements appended to the headline JSON line
(`"secondary": {...}`), each with its own roofline block and a parity check, run after the headline legs on the same
process group:

  cfg3  configs[2]  Llama-3-8B GPTQ int4 (Marlin W4A16) decode step, bs=256 ctx=4096: the headline step with every
                    linear on `gptq_marlin_gemm` (same KV cache), beside the same step over the reference's Marlin kernel.
  cfg4  configs[3]  fp8-e4m3 KV cache, bs=1024 ctx=8192, TP=8: the whole step when the run has 8 GPUs; otherwise the
                    paged-attention kernel at that config's per-GPU shape (4 q-heads, 1 kv-head) against the HBM roofline.
  cfg5  configs[4]  Mixtral-8x7B AWQ int4, bs=128: one MoE block with the reference's semantics for quantised Mixtral
                    (experts split across ranks, dense per expert, one all-reduce; aphrodite/modeling/models/
                    mixtral_quant.py:128-154), TP = the run's world size (configs[4] is TP=4), beside the reference kernels.
Every leg is wrapped: a failure is reported in its slot and never costs the headline numbers.
"""
import statistics

import torch


def _time(env, fn, iters, warm=3):
    for _ in range(warm):
        fn()
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(env.stream)
    for _ in range(iters):
        fn()
    e1.record(env.stream)
    env.barrier()
    return env.max_over_ranks(e0.elapsed_time(e1)) / iters


def _graph_of(env, fn, ca=None):
    import contextlib
    try:
        g = torch.cuda.CUDAGraph()
        with (ca.capture() if ca is not None else contextlib.nullcontext()):
            with torch.cuda.graph(g, stream=env.stream):
                fn()
        return g
    except Exception as e:
        env.log(f"secondary: CUDA graph capture failed ({e!r}); eager")
        torch.cuda.synchronize()
        return None


def _native_split():
    from aphrodite_engine_b200 import _native
    return int(_native.load_c_abi().b200_last_attention_cluster_split())


def _ref_table():
    from oracle import ref_cuda_ops as rco
    if not rco.available():
        return None, None
    t = rco.RefCudaOps()
    return t, rco.make_attention_cls(t)


# ---------------------------------------------------------------------------------------------------------- cfg3
def cfg3_gptq_step(env, args, model, st, shape, num_blocks, dtype, peaks):
    from aphrodite_engine_b200.llama_decode import LlamaDecoder
    import aphrodite_engine_b200._custom_ops as ops
    from aphrodite_engine_b200.scalar_type import scalar_types
    q = LlamaDecoder(shape, args.batch, args.block_size, num_blocks, env.dev, dtype, args.kv_cache_dtype,
                     tp_rank=env.rank, tp_size=env.world, group=env.group, quant="gptq", custom_ar=model.custom_ar,
                     nvls=model.nvls, share_kv_from=model)
    out = {"workload": f"Llama-3-8B GPTQ int4 (uint4b8, group 128, Marlin W4A16) decode bs={args.batch} ctx={args.ctx} "
                       f"tp{env.world} (BASELINE configs[2])"}
    K = 10
    with torch.cuda.stream(env.stream):
        for _ in range(2):
            q.forward(st)
        h_mine = q.last_hidden.float().clone()
        floor = None
        if env.world > 1 and q.tp_mode != "nccl":
            # how far apart two CORRECT exchanges land on this network: this repo's kernels over NCCL's all-reduce
            # (bf16 ring partial sums beyond 2 ranks) against the same kernels over the fp32 rank-order exchange
            twin = LlamaDecoder(shape, args.batch, args.block_size, num_blocks, env.dev, dtype, args.kv_cache_dtype,
                                tp_rank=env.rank, tp_size=env.world, group=env.group, quant="gptq", share_from=q)
            twin.forward(st)
            floor = float((h_mine - twin.last_hidden.float()).norm() / twin.last_hidden.float().norm().clamp_min(1e-30))
            del twin
        g = _graph_of(env, lambda: q.forward(st), model.custom_ar)
        ms = _time(env, (g.replay if g is not None else (lambda: q.forward(st))), K, 3 if env.world == 1 else 10)
        out.update(value=args.batch / (ms * 1e-3), unit="tok/s", ms_per_step=ms, steps=K, cuda_graph=g is not None)
        # the same step over the reference's Marlin kernel (and its other kernels), same weights
        table, attn_cls = _ref_table()
        if table is not None:
            ref_ca = getattr(env, "ref_ca", None)    # the reference's own custom all-reduce (fp32 rank-order sum), if set up
            r = LlamaDecoder(shape, args.batch, args.block_size, num_blocks, env.dev, dtype, args.kv_cache_dtype,
                             tp_rank=env.rank, tp_size=env.world, group=env.group, quant="gptq", op_table=table,
                             attention_cls=attn_cls, custom_ar=ref_ca, share_from=q)
            for _ in range(2):
                r.forward(st)
            rel = float((h_mine - r.last_hidden.float()).norm() / r.last_hidden.float().norm().clamp_min(1e-30))
            gr = _graph_of(env, lambda: r.forward(st), ref_ca)
            ms_r = _time(env, (gr.replay if gr is not None else (lambda: r.forward(st))), 5, 2 if env.world == 1 else 6)
            out["ref_cuda"] = {"ms_per_step": ms_r, "value": args.batch / (ms_r * 1e-3), "ratio": ms_r / ms,
                               "allreduce": getattr(env, "ref_ar", "nccl") if env.world > 1 else "none",
                               "hidden_rel_fro_err_vs_this_repo": rel,
                               "hidden_rel_fro_err_nccl_vs_fp32_exchange_same_kernels": floor,
                               "parity_ok": env.all_agree(rel <= max(3e-2, 3.0 * (floor or 0.0))),
                               "parity_rule": "rel <= max(3e-2, 3 x the distance between two correct exchanges on this "
                                              "network); kernel-level parity at these shard shapes is bit-level-tested in "
                                              "tests/test_gpu_vs_ref_cuda.py::test_marlin_tp_shard_shapes_vs_reference_kernel"}
            del r, gr
        # the four projection GEMMs alone (tensor-bound at M = 256): TFLOP/s against the measured bf16 tensor peak
        if env.world == 1:
            gemm = {}
            flush = torch.empty(192 << 20, dtype=torch.uint8, device=env.dev)
            x_in = {4096: torch.randn(args.batch, 4096, device=env.dev).to(dtype),
                    shape.intermediate: torch.randn(args.batch, shape.intermediate, device=env.dev).to(dtype)}
            for name in ("qkv", "o", "gate_up", "down"):
                w = q.layers[0][name]
                a = x_in[w["k"]]
                ts = []
                for it in range(8):
                    flush.zero_()
                    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s0.record(env.stream)
                    ops.gptq_marlin_gemm(a, w["q"], w["s"], q._empty, q._empty, q._empty, w["ws"], scalar_types.uint4b8,
                                         args.batch, w["n"], w["k"], True, False, True, False)
                    s1.record(env.stream)
                    env.stream.synchronize()
                    if it >= 2:
                        ts.append(s0.elapsed_time(s1))
                us = statistics.median(ts) * 1e3
                tf = 2.0 * args.batch * w["n"] * w["k"] / (us * 1e-6) / 1e12
                gemm[f"{name} {w['k']}x{w['n']}"] = {"us": us, "tflops": tf}
            peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0))
            best = max(v["tflops"] for v in gemm.values())
            out["roofline"] = {"kernel": "marlin_w4a16_tc5_kernel (gptq_marlin_gemm, M=256)", "bound": "tensor",
                               "achieved": best, "peak": peak_tf, "unit": "TFLOP/s", "frac": best / peak_tf,
                               "per_shape": gemm, "l2": "L2 flushed between launches (192 MiB memset)"}
    del q, g
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------------- cfg4
def cfg4_fp8kv(env, args, model, shape, dtype, peaks):
    from aphrodite_engine_b200.llama_decode import DecodeState, LlamaDecoder, make_synthetic_batch, upload
    from aphrodite_engine_b200.attention.paged_attn import PagedAttention
    B, CTX, BS = 1024, 8192, args.block_size
    out = {}
    if env.world == 8 and args.quant is None:
        from aphrodite_engine_b200.distributed.nvls import NvlsTensorParallel
        nvls = None
        if model.nvls is not None:
            nvls = NvlsTensorParallel(env.group, env.dev, B, shape.hidden, dtype, algo=model.nvls.algo)
        host, nb = make_synthetic_batch(B, CTX, BS)
        st = DecodeState(B, host["block_tables"].shape[1], env.dev)
        upload(st, host)
        m = LlamaDecoder(shape, B, BS, nb, env.dev, dtype, "fp8", tp_rank=env.rank, tp_size=env.world, group=env.group,
                         nvls=nvls, share_weights_from=model)
        out["workload"] = "Llama-3-8B fp8-e4m3 KV cache decode bs=1024 ctx=8192 tp8 (BASELINE configs[3]), whole step"
        with torch.cuda.stream(env.stream):
            for _ in range(2):
                m.forward(st)
            h_a = m.last_hidden.float().clone()
            twin = LlamaDecoder(shape, B, BS, nb, env.dev, dtype, "fp8", tp_rank=env.rank, tp_size=env.world,
                                group=env.group, share_from=m)          # NCCL exchange on the same weights / cache
            twin.forward(st)
            rel = float((h_a - twin.last_hidden.float()).norm() / twin.last_hidden.float().norm().clamp_min(1e-30))
            del twin
            exact = None
            if nvls is not None:
                import bench
                exact = bench.exchange_parity_exact(env, "nvls", nvls, None, B, shape.hidden, dtype)
            g = _graph_of(env, lambda: m.forward(st))
            ms = _time(env, (g.replay if g is not None else (lambda: m.forward(st))), 10, 10)
            evs = []
            m.attn_hook = lambda li, begin: (evs.append(torch.cuda.Event(enable_timing=True)), evs[-1].record(env.stream))
            m.forward(st)
            env.stream.synchronize()
            m.attn_hook = None
        attn_ms = statistics.mean(evs[i].elapsed_time(evs[i + 1]) for i in range(0, len(evs), 2))
        heads, kvh = m.heads, m.kv_heads
        out.update(value=B / (ms * 1e-3), unit="tok/s", ms_per_step=ms, steps=10, cuda_graph=g is not None,
                   parity={"hidden_rel_fro_err_vs_nccl_exchange": rel,
                           "exchange_exact_small_int_vs_nccl_plus_norm_at_1024_tokens": exact,
                           "ok": env.all_agree(rel <= 1e-1 and exact is not False),
                           "note": "NCCL's ring rounds partial sums to bf16 beyond 2 ranks; the distance to it is informational"})
    else:
        # per-GPU shape of configs[3]: 32/8 q-heads and 8/8 kv-heads per rank, fp8 cache; two layers' caches alternate
        heads, kvh, D = 4, 1, shape.head_size
        out["workload"] = ("paged_attention_v1 at the per-GPU shape of BASELINE configs[3] (bs=1024 ctx=8192, 4 q-heads, "
                           "1 kv-head, fp8-e4m3 KV); the whole TP=8 step runs when --gpus 8")
        nb_per = CTX // BS
        nb = B * nb_per
        g = torch.Generator(device=env.dev).manual_seed(7)
        caches = []
        import aphrodite_engine_b200._custom_ops as ops
        for _ in range(2):
            kv = torch.empty(PagedAttention.get_kv_cache_shape(nb, BS, kvh, D), dtype=torch.uint8, device=env.dev)
            for plane in range(2):          # fp8 cache = convert_fp8 of U(-s, s), like the reference's benchmark caches
                src = torch.empty(kv[plane].shape, dtype=dtype, device=env.dev).uniform_(-D ** -0.5, D ** -0.5, generator=g)
                ops.convert_fp8(kv[plane], src, 1.0, "fp8")
                del src
            caches.append(PagedAttention.split_kv_cache(kv, kvh, D))
        bt = torch.randperm(nb, device=env.dev, generator=g).view(B, nb_per).to(torch.int32)
        sl = torch.full((B,), CTX, dtype=torch.int32, device=env.dev)
        q = torch.empty(B, heads, D, device=env.dev).uniform_(-D ** -0.5, D ** -0.5, generator=g).to(dtype)
        o = torch.empty_like(q)
        with torch.cuda.stream(env.stream):
            def once(i=[0]):
                kc, vc = caches[i[0] & 1]
                i[0] += 1
                PagedAttention.forward_decode(q, kc, vc, bt, sl, CTX, "fp8", kvh, D ** -0.5, None, 1.0, 1.0, output=o)
            attn_ms = _time(env, once, 20, 4)
            # parity at this shape: a sample of sequences against the oracle
            from oracle import paged_ops as po
            kc, vc = caches[1]
            idx = [0, B // 2, B - 1]
            ref = po.paged_attention(q[idx].cpu(), kc.cpu(), vc.cpu(), bt[idx].cpu(), sl[idx].cpu(), D ** -0.5,
                                     kv_cache_dtype="fp8")
            err = (o[idx].float().cpu() - ref.float()).abs()
            ok = bool((err <= 2e-3 + 1e-3 * ref.float().abs()).all())      # tests/tolerances.py: fp8 KV atol 2e-3, rtol 1e-3
            out["parity"] = {"max_abs_err_vs_oracle_on_3_sequences": float(err.max()), "ok": ok,
                             "cluster_split": _native_split()}
        del caches
    algo_bytes = B * CTX * 2 * kvh * shape.head_size * 1 + 2 * B * heads * shape.head_size * 2 + B * (CTX // BS) * 4
    ach = algo_bytes / (attn_ms * 1e-3) / 1e9
    out["roofline"] = {"kernel": "paged_attention_tc_kernel (fp8 KV)", "bound": "hbm", "achieved": ach,
                       "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                       "algorithmic_bytes_per_launch": algo_bytes, "mean_launch_ms": attn_ms}
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------------- cfg5
def cfg5_mixtral_moe(env, args, model, dtype, peaks):
    from aphrodite_engine_b200.mixtral_moe import MixtralQuantMoE, MixtralShape
    ms_shape = MixtralShape()
    if ms_shape.num_experts % env.world:
        return {"skipped": f"{ms_shape.num_experts} experts do not split over {env.world} ranks"}
    T, NSETS = 128, 4
    layers = [MixtralQuantMoE(ms_shape, env.dev, dtype, tp_rank=env.rank, tp_size=env.world, seed=100 * i + 7)
              for i in range(NSETS)]         # distinct weight sets alternate so that nothing is served from L2
    x = (torch.randn(T, ms_shape.hidden, device=env.dev) * 0.5).to(dtype)
    if env.world > 1:
        env.dist.broadcast(x, 0)
    out = {"workload": f"Mixtral-8x7B AWQ int4 MoE block (mixtral_quant semantics: {len(layers[0].experts)} local "
                       f"experts of 8, dense per expert, one all-reduce), bs={T}, tp{env.world}"
                       + (" = BASELINE configs[4]'s TP" if env.world == 4 else " (configs[4] is TP=4)")}

    def exchange(y):
        if env.world > 1:
            env.dist.all_reduce(y, group=env.group)
        return y

    def run(ls):
        def f():
            for L in ls:
                exchange(L.forward(x))
        return f
    with torch.cuda.stream(env.stream):
        y = exchange(layers[0].forward(x)).float().clone()
        loop = MixtralQuantMoE(ms_shape, env.dev, dtype, tp_rank=env.rank, tp_size=env.world, fused_scale_add=False,
                               share_from=layers[0])
        exact = bool(torch.equal(loop.forward(x), layers[0].forward(x)))
        g = _graph_of(env, run(layers))
        ms = _time(env, (g.replay if g is not None else run(layers)), 10, 3 if env.world == 1 else 10) / NSETS
        out.update(value=T / (ms * 1e-3), unit="tok/s per MoE block", ms_per_layer=ms, cuda_graph=g is not None)
        table, _ = _ref_table()
        par = {"fused_scale_add_equals_reference_loop": env.all_agree(exact)}
        if table is not None:
            refs = [MixtralQuantMoE(ms_shape, env.dev, dtype, tp_rank=env.rank, tp_size=env.world, op_table=table,
                                    share_from=L) for L in layers]
            yr = exchange(refs[0].forward(x)).float()
            rel = float((y - yr).norm() / yr.norm().clamp_min(1e-30))
            par["rel_fro_err_vs_reference_kernels"] = rel
            gr = _graph_of(env, run(refs))
            ms_r = _time(env, (gr.replay if gr is not None else run(refs)), 5, 2 if env.world == 1 else 6) / NSETS
            out["ref_cuda"] = {"ms_per_layer": ms_r, "ratio": ms_r / ms}
            par["ok"] = env.all_agree(exact and rel <= 2e-2)
        else:
            par["ok"] = env.all_agree(exact)
        out["parity"] = par
    H, I, G = ms_shape.hidden, ms_shape.intermediate, ms_shape.group_size
    n_local = len(layers[0].experts)
    per_expert = (H * 2 * I + I * H) // 2 + ((H // G) * 2 * I + (I // G) * H) * 2 + ((H // G) * 2 * I + (I // G) * H) // 2
    flops = 2.0 * T * (H * 2 * I + I * H) * n_local
    peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0))
    tf = flops / (ms * 1e-3) / 1e12
    # dense per expert at 128 tokens: 512 flop per packed-weight byte, above the ~220 flop/B ridge -> tensor-bound
    out["roofline"] = {"kernel": "marlin_w4a16_tc5_kernel (AWQ zero points, M=128), whole MoE block", "bound": "tensor",
                       "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                       "flops_per_layer": flops, "weight_bytes_per_layer": per_expert * n_local,
                       "weight_GBps": per_expert * n_local / (ms * 1e-3) / 1e9}
    del layers
    torch.cuda.empty_cache()
    return out


def run(env, args, model, st, host, shape, num_blocks, dtype, peaks):
    out = {}
    legs = []
    if args.quant is None:
        legs.append(("cfg3_gptq_int4_step", lambda: cfg3_gptq_step(env, args, model, st, shape, num_blocks, dtype, peaks)))
    legs.append(("cfg4_fp8_kv", lambda: cfg4_fp8kv(env, args, model, shape, dtype, peaks)))
    legs.append(("cfg5_mixtral_awq_moe", lambda: cfg5_mixtral_moe(env, args, model, dtype, peaks)))
    for name, fn in legs:
        try:
            out[name] = fn()
        except Exception as e:      # keep the headline numbers whatever happens here
            import traceback
            env.log(f"secondary leg {name} failed: {traceback.format_exc()}")
            out[name] = {"error": repr(e)[:300]}
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
    return out
