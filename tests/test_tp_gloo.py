"""CPU-only, world_size 2 over gloo: the host-side logic of the N > 1 path.

The hot path shards by the reference's tensor-parallel split: attention, rotary and
the cache writer are independent per kv-head (rank r owns q-heads [r*Hq/tp, ...) and kv-heads
[r*Hkv/tp, ...), no communication), and every row-parallel GEMM is followed by ONE sum all-reduce of
[T, hidden]. These tests run the oracle restatement per rank and check that
  * head-sharded attention == the matching slice of the unsharded result (no collective needed),
  * column->row parallel MLP with one all-reduce == the unsharded MLP,
  * vocab-sharded greedy sampling (local max/argmax + all-gather) == the unsharded argmax,
which is exactly what llama_decode.LlamaDecoder does on NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import paged_ops as po


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)   # identical "replicated" inputs on every rank
        S, Hq, Hkv, D, BS, NB, hidden, inter, vocab = 3, 8, 4, 64, 16, 40, 256, 512, 1000
        scale = D ** -0.5
        kc, vc = po.make_kv_cache(NB, BS, Hkv, D, torch.float32, "auto", 0)
        qh = torch.empty(S, Hq, D).uniform_(-scale, scale)
        sl = torch.tensor([5, 100, 333], dtype=torch.int32)
        bt = torch.randint(0, NB, (S, 21), dtype=torch.int32)
        full = po.paged_attention(qh, kc, vc, bt, sl, scale)
        hq, hkv = Hq // world, Hkv // world
        mine = po.paged_attention(qh[:, rank * hq:(rank + 1) * hq], kc[:, rank * hkv:(rank + 1) * hkv],
                                  vc[:, rank * hkv:(rank + 1) * hkv], bt, sl, scale)
        ok_attn = torch.allclose(mine, full[:, rank * hq:(rank + 1) * hq], atol=1e-6)

        x = torch.randn(S, hidden)
        w_gu, w_down = torch.randn(2 * inter, hidden) * 0.05, torch.randn(hidden, inter) * 0.05
        ref = po.silu_and_mul(x @ w_gu.t()) @ w_down.t()
        isz = inter // world
        gate = w_gu[:inter][rank * isz:(rank + 1) * isz]
        up = w_gu[inter:][rank * isz:(rank + 1) * isz]
        part = po.silu_and_mul(x @ torch.cat((gate, up)).t()) @ w_down[:, rank * isz:(rank + 1) * isz].t()
        dist.all_reduce(part)                       # the ONE collective after a row-parallel GEMM
        ok_mlp = torch.allclose(part, ref, atol=1e-4, rtol=1e-4)

        lm = torch.randn(vocab, hidden)
        logits = x @ lm.t()
        vs = vocab // world
        loc = x @ lm[rank * vs:(rank + 1) * vs].t()
        mx, idx = loc.max(dim=-1)
        pair = torch.stack((mx, (idx + rank * vs).float()), dim=-1)
        gathered = [torch.empty_like(pair) for _ in range(world)]
        dist.all_gather(gathered, pair)
        allp = torch.stack(gathered)
        best = allp[..., 0].argmax(dim=0)
        tok = allp[best, torch.arange(S), 1].long()
        ok_tok = torch.equal(tok, logits.argmax(dim=-1))
        q.put((rank, ok_attn, ok_mlp, ok_tok))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_tensor_parallel_split_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, a, m, t in res:
        assert a, f"rank {rank}: head-sharded attention differs"
        assert m, f"rank {rank}: row-parallel MLP + all-reduce differs"
        assert t, f"rank {rank}: sharded greedy token differs"


def _car_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import aphrodite_engine_b200.distributed.custom_all_reduce as car
        # rank 1 pretends it cannot reach its peer; rank 0 believes everything is fine: BOTH must stand down, and
        # neither may be left waiting in the IPC exchange that follows the decision
        calls = []
        car.torch = type("T", (), {"__getattr__": lambda self, n: getattr(torch, n)})()
        car.torch.cuda = type("C", (), {"device_count": staticmethod(lambda: 2),
                                        "can_device_access_peer": staticmethod(lambda a, b: rank == 0)})()
        car._is_full_nvlink = lambda devs: True
        ca = car.CustomAllreduce(dist.group.WORLD, f"cuda:{rank}")
        gathered = ca._gather(("hello", rank))
        q.put((rank, ca.disabled, gathered == [("hello", 0), ("hello", 1)], calls))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_custom_allreduce_enable_decision_is_collective_world2_gloo():
    """The peer-memory all-reduce's enable decision over REAL torch.distributed (gloo): an asymmetric peer-access answer
    disables the communicator on every rank and the constructor returns on all of them (the old per-rank
    early return left the other ranks blocked in broadcast_object_list)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_car_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, disabled, gather_ok, _ in res:
        assert disabled, f"rank {rank} enabled the peer path although rank 1 has no peer access"
        assert gather_ok, f"rank {rank}: object all-gather out of rank order"
