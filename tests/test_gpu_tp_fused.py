"""Multi-GPU parity (>= 2 GPUs): the fused NVSwitch exchange kernel (csrc/tp_fused.cu) against the sequence it replaces —
NCCL all-reduce followed by this package's `fused_add_rms_norm` (itself pinned to the reference kernel in
test_gpu_vs_ref_cuda.py). Small integers make every partial sum exact in bf16/fp16 (the trick of the reference's
tests/distributed/test_custom_all_reduce.py:55-81), so the comparison is `torch.equal`; a second pass with random
normal inputs checks the fp32-accumulate / single-rounding numerics within 1 bf16 ulp of an fp32 restatement.
Both the multicast (multimem) and the unicast peer-pointer variants run, eagerly and under CUDA-graph replay."""
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
        for dtype in (torch.bfloat16, torch.float16):
            for use_mc in (True, False):
                H, TMAX = 4096, 300
                tp = NvlsTensorParallel(dist.group.WORLD, dev, TMAX, H, dtype, use_multicast=use_mc)
                info[f"multicast_{use_mc}"] = tp.multicast
                if use_mc and not tp.multicast:
                    msgs.append("note: no multicast mapping on this box; multimem variant not exercised")
                tag = f"{dtype} mc={tp.multicast}"
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
                    # ---- numerics: random normal, fp32 restatement --------------------------------------------
                    xr = torch.randn(T, H, device=dev).to(dtype)
                    rr = torch.randn(T, H, device=dev).to(dtype)
                    dist.broadcast(rr, 0)
                    s32 = xr.float().clone()
                    dist.all_reduce(s32)
                    z = (s32.to(dtype).float() + rr.float()).to(dtype)
                    var = z.float().pow(2).mean(-1, keepdim=True)
                    want = ((z.float() * torch.rsqrt(var + 1e-5)).to(dtype).float() * w.float()).to(dtype)
                    tp.x(T).copy_(xr)
                    res = rr.clone()
                    h = tp.allreduce_add_rms_norm(T, res, w, 1e-5)
                    torch.cuda.synchronize()
                    err = (h.float() - want.float()).abs()
                    tol = 2e-2 * want.float().abs() + 2e-2     # 1-2 ulp of a 16-bit type after a differently-ordered fp32 sum
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
                    xin.copy_(torch.randint(-2, 3, (T, H), device=dev).to(dtype))
                    r0 = torch.randint(-2, 3, (T, H), device=dev).to(dtype)
                    dist.broadcast(r0, 0)
                    res.copy_(r0)
                    g.replay()
                    torch.cuda.synchronize()
                    a = xin.clone(); dist.all_reduce(a)
                    rres = r0.clone()
                    ops.fused_add_rms_norm(a, rres, w, 1e-5)
                    b = a.clone(); dist.all_reduce(b)
                    ops.fused_add_rms_norm(b, rres, w, 1e-5)
                    lo, hi = tp.rows_of(rank, T)
                    if not (torch.equal(out, b) and torch.equal(res[lo:hi], rres[lo:hi])):
                        msgs.append(f"{tag}: graph replay {it} differs (max {float((out.float()-b.float()).abs().max())})")
                dist.barrier(); torch.cuda.synchronize()
                del tp
        ok = not [m for m in msgs if not m.startswith("note:")]
        q.put((rank, ok, msgs, info))
        dist.barrier()
        torch.cuda.synchronize()
    except Exception as e:  # report instead of hanging the parent
        import traceback
        q.put((rank, False, [repr(e), traceback.format_exc()], {}))
    finally:
        os._exit(0)


@pytest.mark.timeout(300)
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
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    for rank, ok, msgs, info in res:
        print(f"rank {rank}: {info} {msgs}")
        assert ok, f"rank {rank}: {msgs}"
