"""CPU checks of the all-reduce communicator's host logic (aphrodite_engine_b200/distributed/custom_all_reduce.py) with
the native ops, torch.distributed and the CUDA queries mocked: construction protocol (meta / staging buffer / rank_data,
IPC (handle, offset) pairs gathered in rank order — reference custom_all_reduce.py:105-170), the eligibility rule
(:236-255), and the eager / warm-up / captured dispatch of `custom_all_reduce` (:273-296)."""
import types

import pytest
import torch
import torch.distributed as dist

import aphrodite_engine_b200.distributed.custom_all_reduce as car


class _Ops:
    def __init__(self):
        self.calls = []

    def meta_size(self):
        return 256

    def init_custom_ar(self, meta, rank_data, handles, offsets, rank, full_nvlink):
        self.calls.append(("init", meta.numel(), list(handles), list(offsets), rank, full_nvlink))
        return 77

    def register_buffer(self, ptr, t, handles, offsets):
        self.calls.append(("register_buffer", ptr, t.numel(), list(handles), list(offsets)))

    def get_graph_buffer_ipc_meta(self, ptr):
        return torch.arange(128, dtype=torch.uint8), [16, 48]

    def register_graph_buffers(self, ptr, handles, offsets):
        self.calls.append(("register_graph_buffers", ptr, list(handles), list(offsets)))

    def all_reduce_reg(self, ptr, inp, out):
        self.calls.append(("all_reduce_reg", ptr))

    def all_reduce_unreg(self, ptr, inp, buf, out):
        self.calls.append(("all_reduce_unreg", ptr, buf.numel()))

    def dispose(self, ptr):
        self.calls.append(("dispose", ptr))


class _Torch:
    def __init__(self, peers_ok):
        self.stream_capturing = False
        self.cuda = types.SimpleNamespace(device_count=lambda: 4, can_device_access_peer=lambda a, b: peers_ok,
                                          is_current_stream_capturing=lambda: self.stream_capturing)

    def __getattr__(self, name):
        return getattr(torch, name)

    def zeros(self, *a, device=None, **k):
        return torch.zeros(*a, **k)

    def empty(self, *a, device=None, **k):
        return torch.empty(*a, **k)


def _make(monkeypatch, world, rank, peers_ok=True, peer_devices=None, peer_decision=(True, True), nvlink=True, **kw):
    ops, tp = _Ops(), _Torch(peers_ok)
    monkeypatch.setattr(car, "ops", ops)
    monkeypatch.setattr(car.CustomAllreduce, "_ops", ops)
    monkeypatch.setattr(car, "torch", tp)
    monkeypatch.setattr(car, "_is_full_nvlink", lambda devices: nvlink)
    calls = {"n": 0}

    def bcast(lst, src, group=None, device=None):
        # gathers happen in a fixed order: device indices, (p2p ok, nvlink) decisions, then IPC (handle, offset) pairs
        phase = calls["n"] // world
        calls["n"] += 1
        if lst[0] is None:
            if phase == 0:
                lst[0] = (peer_devices or list(range(world)))[src] if src != rank else 1
            elif phase == 1:
                lst[0] = peer_decision
            else:
                lst[0] = (f"peer{src}".encode(), 1000 + src)

    monkeypatch.setattr(car, "dist", types.SimpleNamespace(
        get_backend=lambda g: "gloo", Backend=dist.Backend, get_rank=lambda group=None: rank,
        get_world_size=lambda group=None: world, get_process_group_ranks=lambda group=None: list(range(world)),
        broadcast_object_list=bcast))
    monkeypatch.setattr(torch.UntypedStorage, "_share_cuda_", lambda self: (0, b"mine", 0, 64, 0, 0, 0, 0),
                        raising=False)
    devs = peer_devices or list(range(world))
    return car.CustomAllreduce("cpu-group", f"cuda:{devs[rank] if rank < len(devs) else 0}", max_size=1 << 16, **kw), ops, tp


@pytest.mark.parametrize("world", [1, 3, 5, 16])
def test_unsupported_world_sizes_stay_disabled(monkeypatch, world):
    ca, ops, _ = _make(monkeypatch, world, 0)
    assert ca.disabled and ops.calls == [] and ca.custom_all_reduce(torch.zeros(8)) is None


def test_missing_peer_access_disables(monkeypatch):
    ca, ops, _ = _make(monkeypatch, 2, 0, peers_ok=False)
    assert ca.disabled and ops.calls == []


def test_decision_is_taken_over_the_groups_devices_and_agreed_by_all_ranks(monkeypatch):
    """the P2P gate looks at the GROUP's devices and every rank takes the same decision."""
    ca, ops, _ = _make(monkeypatch, 2, 0, peer_decision=(False, True))      # the peer cannot reach us: we stand down too
    assert ca.disabled and ops.calls == []
    ca, ops, _ = _make(monkeypatch, 2, 0, peer_devices=[0, 9])              # a peer's device index this process cannot see
    assert ca.disabled and ops.calls == []
    ca, ops, _ = _make(monkeypatch, 4, 1, nvlink=False)                     # 4 PCIe-only GPUs: NCCL, like the reference
    assert ca.disabled and ops.calls == []
    ca, ops, _ = _make(monkeypatch, 2, 1, nvlink=False)                     # 2 ranks without NVLink still take the kernel
    assert not ca.disabled and ops.calls[0][-1] is False
    ca, ops, _ = _make(monkeypatch, 4, 2, peer_decision=(True, False))      # a peer saw no NVLink: nobody assumes it
    assert ca.disabled


def test_physical_id_mapping(monkeypatch):
    monkeypatch.delenv("CUDA_VISIBLE_DEVICES", raising=False)
    assert car._physical_ids([0, 2]) == [0, 2]
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "4,5,6,7")
    assert car._physical_ids([0, 3]) == [4, 7]
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "GPU-abc,1")
    with pytest.raises(ValueError):
        car._physical_ids([0])
    assert car._is_full_nvlink([0]) in (True, False)                        # never raises, with or without NVML / a GPU


def test_construction_exchanges_ipc_pairs_in_rank_order(monkeypatch):
    ca, ops, _ = _make(monkeypatch, 4, 2)
    assert not ca.disabled
    kind, meta_bytes, handles, offsets, rank, nvlink = ops.calls[0]
    assert kind == "init" and meta_bytes == 256 + (1 << 16) and rank == 2 and nvlink is True
    assert handles == [b"peer0", b"peer1", b"mine", b"peer3"] and offsets == [1000, 1001, 64, 1003]
    assert ops.calls[1][:3] == ("register_buffer", 77, 1 << 16)           # the eager staging buffer
    ca.close()
    ca.close()
    assert ops.calls.count(("dispose", 77)) == 1


def test_eligibility_rule(monkeypatch):
    ca, _, _ = _make(monkeypatch, 4, 0)
    assert ca.should_custom_ar(torch.zeros(1 << 14, dtype=torch.float32))              # == max_size: taken
    assert not ca.should_custom_ar(torch.zeros((1 << 14) + 4, dtype=torch.float32))    # larger
    assert not ca.should_custom_ar(torch.zeros(6, dtype=torch.float16))                # 12 bytes: not 16-byte multiple
    base = torch.zeros(64, dtype=torch.float32)
    assert ca.should_custom_ar(base[32:]) and ca.should_custom_ar(base[8:40])      # contiguous slices
    assert not ca.should_custom_ar(base[::2])                                     # strided and not a storage tail
    ca2, _, _ = _make(monkeypatch, 4, 0, full_nvlink=False)
    assert not ca2.should_custom_ar(torch.zeros(64))                                   # > 2 ranks need full NVLink
    ca3, _, _ = _make(monkeypatch, 2, 1, full_nvlink=False)
    assert ca3.should_custom_ar(torch.zeros(64))


def test_dispatch_eager_warmup_and_captured(monkeypatch):
    ca, ops, tp = _make(monkeypatch, 2, 0)
    x = torch.zeros(1024, dtype=torch.bfloat16)
    ops.calls.clear()
    assert ca.custom_all_reduce(x).shape == x.shape and ops.calls == [("all_reduce_unreg", 77, 1 << 16)]
    assert ca.custom_all_reduce(torch.zeros(1 << 20)) is None                           # too large: caller uses NCCL
    ops.calls.clear()
    with ca.capture():
        tp.stream_capturing = False
        warm = ca.custom_all_reduce(x)                                                  # warm-up pass: allocation only
        assert warm.shape == x.shape and ops.calls == []
        tp.stream_capturing = True
        assert ca.custom_all_reduce(x).shape == x.shape and ops.calls == [("all_reduce_reg", 77)]
    kind, ptr, handles, offsets = ops.calls[-1]
    assert kind == "register_graph_buffers" and ptr == 77 and len(handles) == 2 and len(handles[0]) == 128
    assert offsets[0] == [16, 48] and offsets[1] == 1001
