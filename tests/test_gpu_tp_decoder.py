"""Multi-GPU parity of the whole decode step under tensor parallelism (>= 2 GPUs): a small Llama-shaped decoder
(aphrodite/modeling/models/llama.py:234-261 call pattern) run with each TP exchange implementation
  nccl (reference fallback path) | p2p (IPC peer-memory all-reduce kernel) | nvls / nvls-p2p (fused exchange kernel,
  fp32 rank-order sum; multicast resp. unicast stores) | nvls-reduce (fused kernel, the switch sums: tolerance only)
on the same weights and KV cache. With 2 ranks every exchange is one fp32 add rounded once, so the three must agree
BIT-EXACTLY on the final hidden state and on the sampled tokens; with more ranks NCCL's ring rounds intermediate sums
to bf16, so nccl is held to a tolerance and p2p == nvls-p2p stay exact (both sum in rank order in fp32).
Rank 0 also builds the TP=1 model from the same seed (weights are slices of the same full tensors) and checks the
TP=N hidden state against it within bf16 tolerance."""
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
        cpu_group = dist.new_group(backend="gloo")
        from aphrodite_engine_b200.distributed import CustomAllreduce
        from aphrodite_engine_b200.distributed.nvls import NvlsTensorParallel
        from aphrodite_engine_b200.llama_decode import (DecodeState, LlamaDecoder, LlamaShape, make_synthetic_batch,
                                                        upload)
        shape = LlamaShape(name="tiny", hidden=1024, layers=3, heads=8, kv_heads=8, head_size=128, intermediate=2048,
                           vocab=4096)
        B, CTX, BS = 24, 200, 16
        host, NB = make_synthetic_batch(B, CTX, BS, vocab_hi=4096)
        st = DecodeState(B, host["block_tables"].shape[1], dev)
        upload(st, host)
        base = LlamaDecoder(shape, B, BS, NB, dev, torch.bfloat16, "auto", tp_rank=rank, tp_size=world,
                            group=dist.group.WORLD)
        msgs, out = [], {}

        def run(model):
            model.forward(st)
            torch.cuda.synchronize()
            return model.last_hidden.float().clone(), st.next_tokens.clone()
        out["nccl"] = run(base)
        ca = CustomAllreduce(cpu_group, dev)
        if not ca.disabled:
            out["p2p"] = run(LlamaDecoder(shape, B, BS, NB, dev, torch.bfloat16, "auto", tp_rank=rank, tp_size=world,
                                          group=dist.group.WORLD, custom_ar=ca, share_from=base))
        for name, algo in (("nvls", "mc_store"), ("nvls-reduce", "mc_reduce"), ("nvls-p2p", "p2p")):
            try:
                tp = NvlsTensorParallel(dist.group.WORLD, dev, B, shape.hidden, torch.bfloat16, algo=algo)
            except RuntimeError as e:
                msgs.append(f"note: {name} skipped ({e})")
                continue
            m = LlamaDecoder(shape, B, BS, NB, dev, torch.bfloat16, "auto", tp_rank=rank, tp_size=world,
                             group=dist.group.WORLD, nvls=tp, share_from=base)
            out[name] = run(m)
            # CUDA-graph replay of the fused-exchange step must reproduce the eager result
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                m.forward(st)
                s.synchronize()
                with torch.cuda.graph(g, stream=s):
                    m.forward(st)
                for _ in range(3):
                    g.replay()
                s.synchronize()
            if not torch.equal(m.last_hidden.float(), out[name][0]):
                msgs.append(f"{name}: graph replay differs from eager")
        h0, t0 = out["nccl"]
        for name, (h, t) in out.items():
            rel = float((h - h0).norm() / h0.norm())
            exact = torch.equal(h, h0) and torch.equal(t, t0)
            if world == 2 and not exact and name != "nvls-reduce":
                msgs.append(f"{name} vs nccl at 2 ranks: not bit-exact (rel {rel:.3e})")
            if rel > (6e-2 if name == "nvls-reduce" else 2e-2):      # the switch's narrowing is not round-to-nearest
                msgs.append(f"{name} vs nccl: rel fro err {rel:.3e}")
        for a, b in (("p2p", "nvls-p2p"), ("p2p", "nvls"), ("nvls", "nvls-p2p")):
            if a in out and b in out and not torch.equal(out[a][0], out[b][0]):
                msgs.append(f"{a} vs {b}: both sum in rank order in fp32 and round once, must be bit-exact")
        # TP=N against TP=1 (same seed => same model)
        if rank == 0:
            one = LlamaDecoder(shape, B, BS, NB, dev, torch.bfloat16, "auto", tp_rank=0, tp_size=1)
            h1, t1 = run(one)
            for name, (h, t) in out.items():
                rel = float((h - h1).norm() / h1.norm())
                if rel > (6e-2 if name == "nvls-reduce" else 2e-2):
                    msgs.append(f"TP={world} ({name}) vs TP=1: rel fro err {rel:.3e}")
        dist.barrier()
        ok = not [m for m in msgs if not m.startswith("note:")]
        q.put((rank, ok, msgs, sorted(out)))
        dist.barrier()
        torch.cuda.synchronize()
    except Exception as e:
        import traceback
        q.put((rank, False, [repr(e), traceback.format_exc()], []))
    finally:
        os._exit(0)


@pytest.mark.timeout(200)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_tp_decoder_exchange_implementations_agree(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=150) for _ in range(world)]
    finally:                      # never leave a rank spinning on a GPU behind a failed or timed-out run
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()
    for rank, ok, msgs, modes in res:
        print(f"rank {rank}: modes {modes} {msgs}")
        assert ok, f"rank {rank}: {msgs}"
