"""Multi-GPU parity (>= 2 GPUs): the fused NVSwitch exchange kernel (csrc/tp_fused.cu) against the sequence it replaces —
a sum all-reduce followed by this package's `fused_add_rms_norm` (itself pinned to the reference kernel in
test_gpu_vs_ref_cuda.py).
  * small integers: every partial sum is exact in bf16/fp16 (the trick of the reference's
    tests/distributed/test_custom_all_reduce.py:55-81), so all three algorithms must equal NCCL + norm bit for bit;
  * random normal inputs: the reference's custom all-reduce accumulates in fp32 in rank order and rounds once
    (kernels/all_reduce/custom_all_reduce.cuh:150-168); "p2p" and "mc_store" restate exactly that, so they must equal
    [gather the ranks' inputs -> fp32 rank-order sum -> round -> fused_add_rms_norm] BIT FOR BIT at any world size;
    "mc_reduce" (the switch sums, and does not round to nearest) is only held to a 2-ulp tolerance.
Eagerly and under CUDA-graph replay (two chained exchanges, as in a decoder layer)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        import aphrodite_engine_b200._custom_ops as ops
        from aphrodite_engine_b200.distributed.nvls import NvlsTensorParallel
        msgs, info = [], {}
        def exact_sum(x):
            """fp32 sum of the ranks' tensors in rank order, rounded once to x.dtype."""
            parts = [torch.empty_like(x) for _ in range(world)]
            dist.all_gather(parts, x)
            acc = torch.zeros_like(x, dtype=torch.float32)
            for part in parts:
                acc += part.float()
            return acc.to(x.dtype)

        for dtype in ((torch.bfloat16, torch.float16) if world <= 4 else (torch.bfloat16,)):
            for algo in ("mc_store", "mc_reduce", "p2p"):
                H, TMAX = 4096, 300
                try:
                    tp = NvlsTensorParallel(dist.group.WORLD, dev, TMAX, H, dtype, algo=algo)
                except RuntimeError as e:
                    msgs.append(f"note: {algo} not exercised ({e})")
                    continue
                info[algo] = True
                exact = algo != "mc_reduce"
                tag = f"{dtype} {algo}"
                w = (torch.randint(1, 4, (H,), device=dev)).to(dtype)
                dist.broadcast(w, 0)
                for T in (1, 7, 256, 300):
                    # ---- exact: small integers -----------------------------------------------------------------
                    torch.manual_seed(1000 * T + rank)
                    x = torch.randint(-3, 4, (T, H), device=dev).to(dtype)
                    res0 = torch.randint(-3, 4, (T, H), device=dev).to(dtype)
                    dist.broadcast(res0, 0)
                    ref_sum = x.clone()
                    dist.all_reduce(ref_sum)
                    ref_h, ref_res = ref_sum.clone(), res0.clone()
                    ops.fused_add_rms_norm(ref_h, ref_res, w, 1e-5)
                    tp.x(T).copy_(x)
                    res = res0.clone()
                    h = tp.allreduce_add_rms_norm(T, res, w, 1e-5)
                    torch.cuda.synchronize()
                    lo, hi = tp.rows_of(rank, T)
                    if not torch.equal(h, ref_h):
                        msgs.append(f"{tag} T={T}: normed rows differ (max {float((h.float()-ref_h.float()).abs().max())})")
                    if not torch.equal(res[lo:hi], ref_res[lo:hi]):
                        msgs.append(f"{tag} T={T}: own residual rows differ")
                    if not torch.equal(torch.cat((res[:lo], res[hi:])), torch.cat((res0[:lo], res0[hi:]))):
                        msgs.append(f"{tag} T={T}: foreign residual rows were touched")
                    if not torch.equal(tp.gather_residual(res, T), ref_res):
                        msgs.append(f"{tag} T={T}: gathered residual stream differs")
                    # plain all-reduce variant
                    tp.x(T).copy_(x)
                    h2 = tp.all_reduce(T)
                    torch.cuda.synchronize()
                    if not torch.equal(h2, ref_sum):
                        msgs.append(f"{tag} T={T}: plain all-reduce differs")
                    # ---- random normal inputs: fp32 rank-order sum, one rounding, then the norm kernel ------------
                    xr = torch.randn(T, H, device=dev).to(dtype)
                    rr = torch.randn(T, H, device=dev).to(dtype)
                    dist.broadcast(rr, 0)
                    want_h, want_res = exact_sum(xr), rr.clone()
                    ops.fused_add_rms_norm(want_h, want_res, w, 1e-5)
                    tp.x(T).copy_(xr)
                    res = rr.clone()
                    h = tp.allreduce_add_rms_norm(T, res, w, 1e-5)
                    torch.cuda.synchronize()
                    if exact:
                        if not (torch.equal(h, want_h) and torch.equal(res[lo:hi], want_res[lo:hi])):
                            msgs.append(f"{tag} T={T}: random-normal result not bit-identical to fp32 rank-order sum + norm "
                                        f"(max {float((h.float() - want_h.float()).abs().max())})")
                    else:
                        err = (h.float() - want_h.float()).abs()
                        tol = 5e-2 * want_h.float().abs() + 5e-2     # a few ulp: the switch does not round to nearest
                        if not bool((err <= tol).all()):
                            msgs.append(f"{tag} T={T}: random-normal mismatch max err {float(err.max())}")
                # ---- CUDA graph: two chained exchanges per replay, as in a decoder layer ---------------------------
                T = 256
                xin = torch.zeros(T, H, dtype=dtype, device=dev)
                res = torch.zeros(T, H, dtype=dtype, device=dev)
                dist.barrier(); torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    tp.x(T).copy_(xin)
                    h1 = tp.allreduce_add_rms_norm(T, res, w, 1e-5)
                    tp.x(T).copy_(h1)                 # next "GEMM" consumes H and produces X again
                    h2 = tp.allreduce_add_rms_norm(T, res, w, 1e-5)
                    out = h2.clone()
                for it in range(4):
                    torch.manual_seed(77 * it + rank)
                    xin.copy_(torch.randn(T, H, device=dev).to(dtype))
                    r0 = torch.randn(T, H, device=dev).to(dtype)
                    dist.broadcast(r0, 0)
                    res.copy_(r0)
                    g.replay()
                    torch.cuda.synchronize()
                    a = exact_sum(xin)
                    rres = r0.clone()
                    ops.fused_add_rms_norm(a, rres, w, 1e-5)
                    b = exact_sum(a)
                    ops.fused_add_rms_norm(b, rres, w, 1e-5)
                    lo, hi = tp.rows_of(rank, T)
                    if exact:
                        if not (torch.equal(out, b) and torch.equal(res[lo:hi], rres[lo:hi])):
                            msgs.append(f"{tag}: graph replay {it} differs (max {float((out.float()-b.float()).abs().max())})")
                    elif not bool(((out.float() - b.float()).abs() <= 4e-2 * b.float().abs() + 4e-2).all()):
                        msgs.append(f"{tag}: graph replay {it} outside tolerance")
                dist.barrier(); torch.cuda.synchronize()
        ok = not [m for m in msgs if not m.startswith("note:")]
        q.put((rank, ok, msgs, info))
        dist.barrier()
        torch.cuda.synchronize()
    except Exception as e:  # report instead of hanging the parent
        import traceback
        q.put((rank, False, [repr(e), traceback.format_exc()], {}))
    finally:
        os._exit(0)


@pytest.mark.timeout(500)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_tp_fused_exchange_matches_nccl_plus_norm(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=170 if world <= 4 else 400) for _ in range(world)]
    finally:                      # never leave a rank spinning on a GPU behind a failed or timed-out run
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()
    for rank, ok, msgs, info in res:
        print(f"rank {rank}: {info} {msgs}")
        assert ok, f"rank {rank}: {msgs}"
