#!/usr/bin/env python3
"""bench.py — decode tok/s of the B200 decode hot path (BASELINE.json metric) and its CPU reference.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU kernels

Workload (config.workload): Llama-3-8B bf16 paged-attention decode, bs=256, ctx=4096, block 16,
random-init weights, synthetic prompts — BASELINE.json configs[1]. A "step" is one full decode step:
32 decoder layers + final norm + lm_head + greedy token, 256 tokens. N > 1 = the reference's tensor
parallel split (heads / column-row), two NCCL all-reduces per layer, strong scaling (the batch is fixed).

Legs of the default (CUDA) arm, all in one process per GPU:
  value     CUDA-graph replay of the step, inputs resident in HBM, CUDA events, max over ranks.
  e2e       same step through the public API with HOST inputs: pinned host->device copy of the step's
            inputs (token ids, positions, slot mapping, seq lens, block tables), graph replay,
            device->host read of the sampled tokens, every step, inside the timed region.
  roofline  paged_attention_v1 (the dominant kernel) timed with CUDA events around each of its launches
            inside eagerly-run steps; achieved = algorithmic bytes per launch / mean launch time, against
            MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline (rank 0, N=1 only) the reference's CPU kernels (oracle/_ref; torch.matmul for the GEMMs)
            on a bounded sample: decoder layers at the full shape, extrapolated to the 32-layer step.
The working set of one step (137 GB KV + 16 GB weights) is far larger than L2, so no explicit L2 flush
is needed between timed iterations.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decode tok/s @ bs=256 seq=4k Llama-3-8B"
UNIT = "tok/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--ctx", type=int, default=4096)
    ap.add_argument("--block-size", type=int, default=16)
    ap.add_argument("--layers", type=int, default=32, help="debug only; the metric is quoted on 32")
    ap.add_argument("--kv-cache-dtype", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-layers", type=int, default=2)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--allreduce", default="auto", choices=["auto", "custom", "nccl"],
                    help="TP all-reduce: NVLink peer-memory kernel or NCCL; auto = the peer-memory kernel up to 4 "
                         "ranks (verified bit-exact vs NCCL on 2 and 4 GPUs), NCCL at 8 (same speed there, and the "
                         "8-rank parity run of this round was inconclusive)")
    ap.add_argument("--quant", default=None, choices=[None, "gptq"],
                    help="gptq = BASELINE configs[2] (GPTQ int4 Marlin W4A16 linears); default bf16 = configs[1]")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/b200_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                sm.append(float(c[1])); mx.append(float(c[2]))
                for n, v in zip(names, c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons),
                       samples=len(sm))
        return out


# ================================================================================================
# reference arm / cpu_baseline: the reference's own CPU kernels on host cores
# ================================================================================================
def cpu_reference_sample(args, n_layers, iters, warmup):
    """Times `n_layers` decoder layers + final norm + lm_head at the full bs/ctx shape on the CPU with
    the reference's kernels (oracle/_ref/*.so built from /root/reference/kernels/cpu) and torch.matmul
    for the unquantised GEMMs (the reference's CPU backend does the same through F.linear).
    Returns (tok/s extrapolated to 32 layers, seconds per 32-layer step, kind, cores, description)."""
    from oracle import ref_lib, paged_ops as po
    from aphrodite_engine_b200.llama_decode import LlamaShape
    ref = ref_lib.load()
    kind = "reference" if ref is not None else "port"
    s = LlamaShape()
    B, CTX, BS, D, H, KV = args.batch, args.ctx, args.block_size, s.head_size, s.heads, s.kv_heads
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dt = torch.bfloat16
    g = torch.Generator().manual_seed(0)
    nb_per = (CTX + BS - 1) // BS
    NB = B * nb_per
    scale = D ** -0.5

    def w(*sz):
        return (torch.randn(*sz, generator=g) * 0.02).to(dt)

    tile = torch.empty(1 << 22, dtype=torch.float32).uniform_(-scale, scale, generator=g).to(dt)
    layers = []
    for _ in range(n_layers):
        kv = torch.empty(2, NB, BS * KV * D, dtype=dt)
        flat = kv.view(-1)
        for off in range(0, flat.numel(), tile.numel()):
            n = min(tile.numel(), flat.numel() - off)
            flat[off:off + n] = tile[:n]
        kc = kv[0].view(NB, KV, D // 8, BS, 8)
        vc = kv[1].view(NB, KV, D, BS)
        layers.append(dict(kc=kc, vc=vc, ln1=torch.ones(s.hidden, dtype=dt), ln2=torch.ones(s.hidden, dtype=dt),
                           qkv=w((H + 2 * KV) * D, s.hidden), o=w(s.hidden, H * D),
                           gate_up=w(2 * s.intermediate, s.hidden), down=w(s.hidden, s.intermediate)))
    norm_w, lm_head = torch.ones(s.hidden, dtype=dt), w(s.vocab, s.hidden)
    inv = 1.0 / (s.rope_theta ** (torch.arange(0, D, 2, dtype=torch.float32) / D))
    fr = torch.einsum("i,j->ij", torch.arange(s.max_position, dtype=torch.float32), inv)
    cos_sin = torch.cat((fr.cos(), fr.sin()), dim=-1).to(dt)
    bt = torch.randperm(NB, generator=g).view(B, nb_per).to(torch.int32)
    sl = torch.full((B,), CTX, dtype=torch.int32)
    pos = torch.full((B,), CTX - 1, dtype=torch.long)
    slot = bt[:, (CTX - 1) // BS].long() * BS + (CTX - 1) % BS
    hidden0 = (torch.randn(B, s.hidden, generator=g)).to(dt)

    if ref is not None:
        rops, rcache, _ = ref

        def layer_fwd(L, hidden, residual):
            rops.fused_add_rms_norm(hidden, residual, L["ln1"], s.rms_eps)
            qkv = torch.nn.functional.linear(hidden, L["qkv"])
            q, k, v = qkv.split([H * D, KV * D, KV * D], dim=-1)
            q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
            rops.rotary_embedding(pos, q, k, D, cos_sin, True)
            rcache.reshape_and_cache(k.view(B, KV, D), v.view(B, KV, D), L["kc"], L["vc"], slot, "auto", 1.0, 1.0)
            out = torch.empty(B, H, D, dtype=dt)
            rops.paged_attention_v1(out, q.view(B, H, D), L["kc"], L["vc"], KV, scale, bt, sl, BS, CTX,
                                    None, "auto", 1.0, 1.0, 0, 0, 0, 64, 0)
            hidden = torch.nn.functional.linear(out.view(B, -1), L["o"])
            rops.fused_add_rms_norm(hidden, residual, L["ln2"], s.rms_eps)
            gu = torch.nn.functional.linear(hidden, L["gate_up"])
            act = torch.empty(B, s.intermediate, dtype=dt)
            rops.silu_and_mul(act, gu)
            return torch.nn.functional.linear(act, L["down"]), residual

        def head_fwd(hidden, residual):
            rops.fused_add_rms_norm(hidden, residual, norm_w, s.rms_eps)
            return torch.nn.functional.linear(hidden, lm_head).argmax(dim=-1)
    else:  # host CPU cannot run the AVX-512 build: time the Python/torch restatement instead
        def layer_fwd(L, hidden, residual):
            hidden, residual = po.fused_add_rms_norm(hidden, residual, L["ln1"], s.rms_eps)
            qkv = torch.nn.functional.linear(hidden, L["qkv"])
            q, k, v = qkv.split([H * D, KV * D, KV * D], dim=-1)
            q, k = po.rotary_embedding(pos, q, k, D, cos_sin, True)
            po.reshape_and_cache(k.reshape(B, KV, D), v.reshape(B, KV, D), L["kc"], L["vc"], slot)
            out = po.paged_attention(q.reshape(B, H, D), L["kc"], L["vc"], bt, sl, scale)
            hidden = torch.nn.functional.linear(out.view(B, -1), L["o"])
            hidden, residual = po.fused_add_rms_norm(hidden, residual, L["ln2"], s.rms_eps)
            act = po.silu_and_mul(torch.nn.functional.linear(hidden, L["gate_up"]))
            return torch.nn.functional.linear(act, L["down"]), residual

        def head_fwd(hidden, residual):
            hidden, _ = po.fused_add_rms_norm(hidden, residual, norm_w, s.rms_eps)
            return torch.nn.functional.linear(hidden, lm_head).argmax(dim=-1)

    def one_sample():
        hidden, residual = hidden0.clone(), hidden0.clone()
        t0 = time.perf_counter()
        for L in layers:
            hidden, residual = layer_fwd(L, hidden, residual)
        t1 = time.perf_counter()
        head_fwd(hidden, residual)
        t2 = time.perf_counter()
        return (t1 - t0) / n_layers, t2 - t1

    for _ in range(warmup):
        one_sample()
    per_layer, head = [], []
    for _ in range(iters):
        a, b = one_sample()
        per_layer.append(a); head.append(b)
    step_s = statistics.mean(per_layer) * s.layers + statistics.mean(head)
    desc = (f"{n_layers} decoder layer(s) + final norm + lm_head at bs={B} ctx={CTX} bf16 on the CPU "
            f"({'reference kernels/cpu build' if ref is not None else 'python restatement'}; GEMMs via "
            f"torch.matmul), {iters} timed pass(es), layer time x32 + head")
    return B / step_s, step_s, kind, cores, desc


def run_reference_arm(args, rank):
    if rank != 0:
        return
    iters = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    t0 = time.perf_counter()
    val, step_s, kind, cores, desc = cpu_reference_sample(args, 1, iters, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": iters, "warmup": warm, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"Llama-3-8B bf16 paged-attention decode bs={args.batch} ctx={args.ctx} "
                               f"block={args.block_size} (BASELINE configs[1]); CPU: bounded sample per step",
                   "l2": "working set >> L2"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ================================================================================================
# CUDA arm
# ================================================================================================
def run_b200(args):
    # keep stdout to exactly ONE JSON line: libraries (NCCL prints its version banner) write to fd 1 too
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    import torch.distributed as dist
    from aphrodite_engine_b200.llama_decode import (DecodeState, LlamaDecoder, LlamaShape,
                                                    make_synthetic_batch, upload)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group, ca = None, None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
        if args.allreduce == "custom" or (args.allreduce == "auto" and world <= 4):
            from aphrodite_engine_b200.distributed import CustomAllreduce
            ca = CustomAllreduce(dist.new_group(backend="gloo"), dev)
            if ca.disabled:
                ca = None

    shape = LlamaShape(layers=args.layers)
    host, num_blocks = make_synthetic_batch(args.batch, args.ctx, args.block_size)
    model = LlamaDecoder(shape, args.batch, args.block_size, num_blocks, dev, torch.bfloat16,
                         args.kv_cache_dtype, tp_rank=rank, tp_size=world, group=group, quant=args.quant, custom_ar=ca)
    st = DecodeState(args.batch, host["block_tables"].shape[1], dev)
    h2d_bytes = upload(st, host)
    torch.cuda.synchronize()

    stream = torch.cuda.Stream(device=dev)
    graph = None
    with torch.cuda.stream(stream):
        for _ in range(2):                   # eager warm-up (cuBLAS workspaces, NCCL channels)
            model.forward(st)
        stream.synchronize()
        if not args.no_graph:
            try:
                import contextlib
                graph = torch.cuda.CUDAGraph()
                with (ca.capture() if ca is not None else contextlib.nullcontext()):
                    with torch.cuda.graph(graph, stream=stream):
                        model.forward(st)
            except Exception as e:       # keep measuring, eagerly, and say so
                graph = None
                if rank == 0:
                    print(f"[bench] CUDA graph capture failed ({e}); running eagerly", file=sys.stderr)
                torch.cuda.synchronize()

    def step():
        if graph is not None:
            graph.replay()
        else:
            model.forward(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local)
    with torch.cuda.stream(stream):
        # ---------------- leg 1: device-resident ----------------
        for _ in range(W):
            step()
        barrier()
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(K):
            step()
        e1.record(stream)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms_dev = max_over_ranks(e0.elapsed_time(e1)) / K

        # ---------------- leg 2: end to end from host buffers ----------------
        out_host = torch.empty(args.batch, dtype=torch.long).pin_memory()
        for _ in range(W):
            upload(st, host); step(); out_host.copy_(st.next_tokens, non_blocking=True); stream.synchronize()
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for _ in range(K):
            upload(st, host)
            step()
            out_host.copy_(st.next_tokens, non_blocking=True)
            stream.synchronize()             # the sampled tokens are needed on the host every step
        e3.record(stream)
        barrier()
        ms_e2e = max_over_ranks(e2.elapsed_time(e3)) / K
        d2h_bytes = out_host.numel() * out_host.element_size()

        # ---------------- leg 3: attention launches timed inside eager steps ----------------
        n_prof_steps = 2
        evs = []

        def hook(li, begin):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            evs.append(ev)
        model.attn_hook = hook
        model.forward(st)                    # warm the eager path
        evs.clear()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(n_prof_steps):
            model.forward(st)
        s1.record(stream)
        stream.synchronize()
        model.attn_hook = None
        attn_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(0, len(evs), 2)]
        eager_step_ms = s0.elapsed_time(s1) / n_prof_steps

    peaks, peak_kind = measured_peaks()
    kv_esz = 2 if args.kv_cache_dtype == "auto" else 1
    nb_per = host["block_tables"].shape[1]
    algo_bytes = (args.batch * args.ctx * 2 * model.kv_heads * shape.head_size * kv_esz
                  + 2 * args.batch * model.heads * shape.head_size * 2 + args.batch * nb_per * 4)
    attn_mean_ms = statistics.mean(attn_ms)
    achieved = algo_bytes / (attn_mean_ms * 1e-3) / 1e9
    roofline = {
        "kernel": "paged_attention_tc_kernel (paged_attention_v1)", "bound": "hbm",
        "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
        "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback",
        "traffic": None, "algorithmic_bytes_per_launch": algo_bytes, "launches_timed": len(attn_ms),
        "mean_launch_ms": attn_mean_ms, "share_of_eager_step": attn_mean_ms * shape.layers / eager_step_ms,
        "eager_step_ms": eager_step_ms,
    }
    traffic_file = os.path.join(ROOT, "profiles", "attention_traffic.json")
    if os.path.exists(traffic_file) and world == 1 and (args.batch, args.ctx, args.block_size, args.kv_cache_dtype) == (256, 4096, 16, "auto"):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_launch")
        except Exception:
            pass

    line = {
        "metric": METRIC, "value": args.batch / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"Llama-3-8B bf16 paged-attention decode bs={args.batch} ctx={args.ctx} "
                               f"block={args.block_size} layers={shape.layers} "
                               + ("GPTQ int4 Marlin W4A16 linears (BASELINE configs[2])" if args.quant else "(BASELINE configs[1])"),
                   "parallelism": f"tp{world}", "allreduce": ("nvlink-p2p" if ca is not None else ("nccl" if world > 1 else "none")), "kv_cache_dtype": args.kv_cache_dtype,
                   "cuda_graph": graph is not None, "l2": "working set (KV + weights) >> 126 MB L2, no flush needed"},
        "e2e": {"value": args.batch / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e},
        "gpu_launches": model.my_kernel_launches_per_step * K,
        "roofline": roofline, "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        del model
        torch.cuda.empty_cache()
        try:
            val, step_s, kind, cores, desc = cpu_reference_sample(args, args.cpu_sample_layers, 2, 1)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc,
                                    "s_per_step": step_s}
        except Exception as e:  # never lose the GPU numbers to a host-side problem
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "error": repr(e)}
    if rank == 0:
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        # Tear-down: NCCL communicators captured in a live CUDA graph can block destroy_process_group();
        # everything has been measured and printed, so synchronise and leave without running destructors.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the decode path has no CPU fallback "
                         "(use --impl reference for the CPU reference arm)")
    run_b200(args)


if __name__ == "__main__":
    main()
