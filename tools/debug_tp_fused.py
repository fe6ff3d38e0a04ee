"""Step-by-step bring-up of the fused TP exchange kernel under torchrun (2+ ranks): prints before / after every step."""
import faulthandler, os, sys, time
import torch, torch.distributed as dist
faulthandler.enable()
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def say(*a):
    print(f"[r{rank} {time.time() % 1000:7.2f}]", *a, flush=True)


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aphrodite_engine_b200._custom_ops as ops
from aphrodite_engine_b200.distributed.nvls import NvlsTensorParallel
H = 4096
keep = []
for algo in ("mc_store", "mc_reduce", "p2p"):
    say("construct", algo)
    tp = NvlsTensorParallel(dist.group.WORLD, dev, 256, H, torch.bfloat16, algo=algo)
    keep.append(tp)
    if algo == "mc_reduce":
        # rounding census of the switch's narrowing: sum of the ranks' random bf16 values vs RN / RZ of the fp32 sum
        torch.manual_seed(5 + rank)
        x = torch.randn(256, H, device=dev).to(torch.bfloat16)
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x)
        acc = sum(p_.float() for p_ in parts)
        rn = acc.to(torch.bfloat16)
        rz = (acc.view(torch.int32) & -65536).view(torch.float32).to(torch.bfloat16)
        tp.x(256).copy_(x); torch.cuda.synchronize(); dist.barrier()
        got = tp.all_reduce(256).clone(); torch.cuda.synchronize()
        n = got.numel()
        say(f"multimem.ld_reduce rounding census over {n} sums: == RN-even {int((got == rn).sum())}, == truncation "
            f"{int((got == rz).sum())}, neither {int(((got != rn) & (got != rz)).sum())}, inexact sums {int((rn.float() != acc).sum())}")
    say("constructed: multicast", tp.multicast, "mc_base", hex(tp.mc_base), "peers", [hex(p) for p in tp.peer_bases],
        "block", hex(tp.block.data_ptr()), "flag_bytes", tp.flag_bytes)
    for T in (8, 256):
        torch.manual_seed(T + rank)
        x = torch.randint(-3, 4, (T, H), device=dev).to(torch.bfloat16)
        ref = x.clone(); dist.all_reduce(ref); torch.cuda.synchronize()
        tp.x(T).copy_(x); torch.cuda.synchronize(); dist.barrier()
        say("launch plain all_reduce T", T)
        h = tp.all_reduce(T)
        torch.cuda.synchronize()
        say("done; equal:", bool(torch.equal(h, ref)))
        w = torch.ones(H, dtype=torch.bfloat16, device=dev)
        res0 = torch.randint(-3, 4, (T, H), device=dev).to(torch.bfloat16); dist.broadcast(res0, 0)
        rh, rr = ref.clone(), res0.clone()
        ops.fused_add_rms_norm(rh, rr, w, 1e-5)
        tp.x(T).copy_(x); res = res0.clone(); torch.cuda.synchronize(); dist.barrier()
        say("launch fused T", T)
        h = tp.allreduce_add_rms_norm(T, res, w, 1e-5)
        torch.cuda.synchronize()
        lo, hi = tp.rows_of(rank, T)
        say("done; equal:", bool(torch.equal(h, rh)), bool(torch.equal(res[lo:hi], rr[lo:hi])))
    # timing, eager back-to-back and CUDA graph
    T = 256
    res = torch.zeros(T, H, dtype=torch.bfloat16, device=dev)
    w = torch.ones(H, dtype=torch.bfloat16, device=dev)
    for _ in range(20): tp.allreduce_add_rms_norm(T, res, w, 1e-5)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): tp.allreduce_add_rms_norm(T, res, w, 1e-5)
    e1.record(); torch.cuda.synchronize()
    say(f"eager fused: {e0.elapsed_time(e1) / 200 * 1e3:.2f} us/call")
    # where the time goes inside one call: %globaltimer stamps of CTA 0 (median over 30 eager calls)
    from aphrodite_engine_b200 import _native
    lib = _native.load_c_abi()
    stamps = torch.zeros(5, dtype=torch.int64, device=dev)
    lib.b200_tp_set_stamp_buffer(stamps.data_ptr())
    rows = []
    for _ in range(30):
        tp.allreduce_add_rms_norm(T, res, w, 1e-5)
        torch.cuda.synchronize()
        v = stamps.tolist()
        rows.append([v[i + 1] - v[i] for i in range(4)])
    lib.b200_tp_set_stamp_buffer(None)
    med = [sorted(r[i] for r in rows)[len(rows) // 2] for i in range(4)]
    say(f"{algo} phases (ns, median): start-barrier {med[0]}, loads {med[1]}, norm+stores-issued {med[2]}, end-barrier {med[3]}")
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(50): tp.allreduce_add_rms_norm(T, res, w, 1e-5)
        for _ in range(3): g.replay()
        s.synchronize(); dist.barrier()
        e0.record(s)
        for _ in range(10): g.replay()
        e1.record(s); s.synchronize()
    say(f"graph fused: {e0.elapsed_time(e1) / 500 * 1e3:.2f} us/call")
    # reference: NCCL all-reduce + fused_add_rms_norm in a graph
    y = torch.zeros(T, H, dtype=torch.bfloat16, device=dev)
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        for _ in range(3):
            dist.all_reduce(y); ops.fused_add_rms_norm(y, res, w, 1e-5)
        s.synchronize()
        with torch.cuda.graph(g2, stream=s):
            for _ in range(50):
                dist.all_reduce(y); ops.fused_add_rms_norm(y, res, w, 1e-5)
        for _ in range(3): g2.replay()
        s.synchronize(); dist.barrier()
        e0.record(s)
        for _ in range(10): g2.replay()
        e1.record(s); s.synchronize()
    say(f"graph NCCL all-reduce + fused_add_rms_norm: {e0.elapsed_time(e1) / 500 * 1e3:.2f} us/pair")
dist.barrier(); torch.cuda.synchronize()
say("all done")
os._exit(0)
