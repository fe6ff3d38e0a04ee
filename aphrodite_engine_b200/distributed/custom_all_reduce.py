"""Tensor-parallel all-reduce communicator over this package's `_C_custom_ar` ops, with the interface of the
reference's `CustomAllreduce` (aphrodite/distributed/device_communicators/custom_all_reduce.py:40-296): constructed
on a NON-NCCL process group; `custom_all_reduce(x)` returns the reduced tensor or None when the caller must fall back
to NCCL; `capture()` wraps CUDA-graph capture; `close()` releases the native object.

Protocol (it is the native side's contract, csrc/custom_all_reduce.cu): every rank owns
  * `meta`      = [signal flags | scratch for the two-shot partial sums], zero-initialised,
  * `buffer`    = a staging area eager-mode inputs are copied into (IPC-registered once),
  * `rank_data` = device table of peer pointers, one entry per registered buffer,
and the ranks exchange (cudaIpcMemHandle of the base allocation, byte offset) pairs over the CPU group. Under graph
capture the kernel records the addresses it was given; `capture()` registers them all at exit."""
from contextlib import contextmanager
from typing import Optional, Union

import torch
import torch.distributed as dist
from torch.distributed import ProcessGroup

from .. import _custom_ops as ops

_WORLD_SIZES = (2, 4, 6, 8)
_RANK_DATA_BYTES = 8 * 1024 * 1024


def _physical_ids(visible_indices):
    """CUDA-visible device indices -> NVML indices (CUDA_VISIBLE_DEVICES may renumber; UUID entries cannot be mapped)."""
    import os
    env = os.environ.get("CUDA_VISIBLE_DEVICES")
    if not env:
        return list(visible_indices)
    table = [int(t) for t in env.split(",") if t.strip().isdigit()]
    if len(table) != len([t for t in env.split(",") if t.strip()]):
        raise ValueError("CUDA_VISIBLE_DEVICES holds non-numeric entries")
    return [table[i] for i in visible_indices]


def _is_full_nvlink(visible_indices) -> bool:
    """True iff NVML reports an NVLink P2P path between EVERY pair of the group's devices (the reference's rule,
    device_communicators/custom_all_reduce.py:_is_full_nvlink). A definite "no" from NVML means False (beyond two ranks
    the caller then keeps NCCL, as the reference does on PCIe-only boxes). When NVML cannot be asked at all (module
    missing, no permission in the container, UUID-style CUDA_VISIBLE_DEVICES) the answer falls back to what the caller
    already established — peer access between every pair — and says so."""
    try:
        import pynvml
        pynvml.nvmlInit()
    except Exception as e:
        import warnings
        warnings.warn(f"NVML unavailable ({e!r}): assuming the peer-accessible devices are NVLink-connected")
        return True
    try:
        handles = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in _physical_ids(visible_indices)]
        for i, a in enumerate(handles):
            for b in handles[i + 1:]:
                if pynvml.nvmlDeviceGetP2PStatus(a, b, pynvml.NVML_P2P_CAPS_INDEX_NVLINK) != pynvml.NVML_P2P_STATUS_OK:
                    return False
        return True
    except Exception as e:
        import warnings
        warnings.warn(f"NVML query failed ({e!r}): assuming the peer-accessible devices are NVLink-connected")
        return True
    finally:
        try:
            pynvml.nvmlShutdown()
        except Exception:
            pass


def is_weak_contiguous(inp: torch.Tensor) -> bool:
    """Contiguous, or a view that covers its storage's tail exactly (what the IPC registration can address)."""
    if inp.is_contiguous():
        return True
    tail_bytes = inp.untyped_storage().nbytes() - inp.storage_offset() * inp.element_size()
    return tail_bytes == inp.numel() * inp.element_size()


class CustomAllreduce:
    _SUPPORTED_WORLD_SIZES = list(_WORLD_SIZES)
    _ops = ops      # op table (a checker may substitute the reference's own kernels behind the same protocol)

    def __init__(self, group: ProcessGroup, device: Union[int, str, torch.device], max_size: int = 8192 * 1024,
                 full_nvlink: Optional[bool] = None) -> None:
        self._IS_CAPTURING = False
        self.disabled = True
        self.group = group
        assert dist.get_backend(group) != dist.Backend.NCCL, "CustomAllreduce should be attached to a non-NCCL group."
        self.rank = dist.get_rank(group=group)
        self.world_size = dist.get_world_size(group=group)
        if self.world_size not in _WORLD_SIZES:          # includes the single-rank case
            return
        self.device = torch.device(f"cuda:{device}") if isinstance(device, int) else torch.device(device)
        # The decision to take the peer-memory path must be the SAME on every rank (a rank that returned early would
        # leave the others blocked in the IPC exchange below), and it concerns the GROUP's devices, not every GPU this
        # process can see: gather the ranks' device indices, test P2P among exactly those, then agree.
        here = self.device.index
        peers = self._gather(here)
        visible = torch.cuda.device_count()
        local_ok = all(0 <= d < visible for d in peers) and len(set(peers)) == len(peers) and all(
            torch.cuda.can_device_access_peer(here, d) for d in peers if d != here)
        if full_nvlink is None:                          # like the reference: ask NVML about every pair of the group
            full_nvlink = _is_full_nvlink(peers) if local_ok else False
        decisions = self._gather((bool(local_ok), bool(full_nvlink)))
        if not all(ok for ok, _ in decisions):
            return                                       # no P2P between some pair: leave it to NCCL (on all ranks)
        full_nvlink = all(nv for _, nv in decisions)
        if self.world_size > 2 and not full_nvlink:
            return                                       # > 2 PCIe-only GPUs: the reference falls back to NCCL as well
        self.max_size = max_size
        self.full_nvlink = full_nvlink
        self.meta = torch.zeros(self._ops.meta_size() + max_size, dtype=torch.uint8, device=self.device)
        self.buffer = torch.empty(max_size, dtype=torch.uint8, device=self.device)
        self.rank_data = torch.empty(_RANK_DATA_BYTES, dtype=torch.uint8, device=self.device)
        self.disabled = False
        handles, offsets = self._exchange(*self._ipc_of(self.meta))
        self._ptr = self._ops.init_custom_ar(self.meta, self.rank_data, handles, offsets, self.rank, self.full_nvlink)
        self.register_buffer(self.buffer)

    # ---- IPC plumbing --------------------------------------------------------------------------------------
    @staticmethod
    def _ipc_of(t: torch.Tensor):
        shared = t.untyped_storage()._share_cuda_()      # (device, handle, size, offset, ...)
        return shared[1], shared[3]

    def _gather(self, obj):
        """All-gather of one picklable object per rank over the CPU group, in rank order."""
        slots = [[None] for _ in range(self.world_size)]
        slots[self.rank][0] = obj
        for i, src in enumerate(sorted(dist.get_process_group_ranks(group=self.group))):
            dist.broadcast_object_list(slots[i], src=src, group=self.group, device="cpu")
        return [slot[0] for slot in slots]

    def _exchange(self, handle, offset):
        """All-gather of one (handle, offset) pair per rank, in rank order."""
        pairs = self._gather((handle, offset))
        return [h for h, _ in pairs], [o for _, o in pairs]

    def register_buffer(self, inp: torch.Tensor):
        handles, offsets = self._exchange(*self._ipc_of(inp))
        self._ops.register_buffer(self._ptr, inp, handles, offsets)

    def register_graph_buffers(self):
        handle_bytes, offset_list = self._ops.get_graph_buffer_ipc_meta(self._ptr)
        handles, offsets = self._exchange(bytes(handle_bytes.numpy().tobytes()), offset_list)
        self._ops.register_graph_buffers(self._ptr, handles, offsets)

    @contextmanager
    def capture(self):
        """While active, `custom_all_reduce` records buffer addresses instead of requiring registered ones; they are
        exchanged and registered when the block exits."""
        self._IS_CAPTURING = True
        try:
            yield
        finally:
            self._IS_CAPTURING = False
            if not self.disabled:
                self.register_graph_buffers()

    # ---- dispatch ------------------------------------------------------------------------------------------
    def should_custom_ar(self, inp: torch.Tensor) -> bool:
        if self.disabled:
            return False
        nbytes = inp.numel() * inp.element_size()
        if nbytes % 16 or not is_weak_contiguous(inp):
            return False
        if not (self.world_size == 2 or self.full_nvlink):
            return False
        return nbytes <= self.max_size          # the reference uses `<` (custom_all_reduce.py:250)

    def all_reduce_reg(self, inp: torch.Tensor, out: Optional[torch.Tensor] = None):
        out = torch.empty_like(inp) if out is None else out
        self._ops.all_reduce_reg(self._ptr, inp, out)
        return out

    def all_reduce_unreg(self, inp: torch.Tensor, out: Optional[torch.Tensor] = None):
        out = torch.empty_like(inp) if out is None else out
        self._ops.all_reduce_unreg(self._ptr, inp, self.buffer, out)
        return out

    def custom_all_reduce(self, input: torch.Tensor) -> Optional[torch.Tensor]:
        if self.disabled or not self.should_custom_ar(input):
            return None
        if not self._IS_CAPTURING:
            return self.all_reduce_unreg(input)          # eager: staged through the registered buffer
        if torch.cuda.is_current_stream_capturing():
            return self.all_reduce_reg(input)            # recorded now, registered when capture() exits
        return torch.empty_like(input)                   # warm-up run before capture: mimic the allocation pattern

    def close(self):
        if not self.disabled and getattr(self, "_ptr", 0):
            self._ops.dispose(self._ptr)
            self._ptr = 0

    def __del__(self):
        self.close()
