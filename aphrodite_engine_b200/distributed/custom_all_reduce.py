"""Mirror of the reference's CustomAllreduce communicator
(aphrodite/distributed/device_communicators/custom_all_reduce.py:40-296) over this package's `_C_custom_ar` ops:
same construction protocol (meta = signals + scratch, a pre-registered staging buffer, a rank_data table; IPC handles
gathered over a NON-NCCL process group), same `should_custom_ar` rule, eager (`all_reduce_unreg`) and CUDA-graph
(`capture()` + `all_reduce_reg` + `register_graph_buffers`) paths."""
from contextlib import contextmanager
from typing import Any, List, Optional, Union

import torch
import torch.distributed as dist
from torch.distributed import ProcessGroup

from .. import _custom_ops as ops


def is_weak_contiguous(inp: torch.Tensor) -> bool:
    return inp.is_contiguous() or (inp.untyped_storage().nbytes() - inp.storage_offset() * inp.element_size()
                                   == inp.numel() * inp.element_size())


class CustomAllreduce:
    _SUPPORTED_WORLD_SIZES = [2, 4, 6, 8]

    def __init__(self, group: ProcessGroup, device: Union[int, str, torch.device],
                 max_size: int = 8192 * 1024, full_nvlink: bool = True) -> None:
        self._IS_CAPTURING = False
        self.disabled = True
        self.group = group
        assert dist.get_backend(group) != dist.Backend.NCCL, \
            "CustomAllreduce should be attached to a non-NCCL group."
        rank = dist.get_rank(group=self.group)
        world_size = dist.get_world_size(group=self.group)
        if world_size == 1 or world_size not in CustomAllreduce._SUPPORTED_WORLD_SIZES:
            return
        if isinstance(device, int):
            device = torch.device(f"cuda:{device}")
        elif isinstance(device, str):
            device = torch.device(device)
        self.device = device
        # software / runtime P2P support between every pair of local devices
        for peer in range(torch.cuda.device_count()):
            if peer != device.index and not torch.cuda.can_device_access_peer(device.index, peer):
                return
        self.disabled = False
        # signals + scratch for the two-shot intermediate results (zeroed: flags start at 0)
        self.meta = torch.zeros(ops.meta_size() + max_size, dtype=torch.uint8, device=self.device)
        # pre-registered staging buffer for eager mode
        self.buffer = torch.empty(max_size, dtype=torch.uint8, device=self.device)
        # device table of peer pointers, 8 * world_size bytes per registered buffer
        self.rank_data = torch.empty(8 * 1024 * 1024, dtype=torch.uint8, device=self.device)
        self.max_size = max_size
        self.rank = rank
        self.world_size = world_size
        self.full_nvlink = full_nvlink
        handles, offsets = self._get_ipc_meta(self.meta)
        self._ptr = ops.init_custom_ar(self.meta, self.rank_data, handles, offsets, rank, self.full_nvlink)
        self.register_buffer(self.buffer)

    @contextmanager
    def capture(self):
        """Registers, at exit, every buffer address the all-reduce saw during CUDA-graph capture."""
        try:
            self._IS_CAPTURING = True
            yield
        finally:
            self._IS_CAPTURING = False
            if not self.disabled:
                self.register_graph_buffers()

    def _get_ipc_meta(self, inp: torch.Tensor):
        data = inp.untyped_storage()._share_cuda_()
        return self._gather_ipc_meta((data[1], data[3]))     # (ipc handle of the base allocation, offset)

    def _gather_ipc_meta(self, shard_data):
        all_data: List[Optional[Any]] = [[None] for _ in range(self.world_size)]
        all_data[self.rank][0] = shard_data
        ranks = sorted(dist.get_process_group_ranks(group=self.group))
        for i, r in enumerate(ranks):
            dist.broadcast_object_list(all_data[i], src=r, group=self.group, device="cpu")
        handles = [all_data[i][0][0] for i in range(len(all_data))]
        offsets = [all_data[i][0][1] for i in range(len(all_data))]
        return handles, offsets

    def register_buffer(self, inp: torch.Tensor):
        handles, offsets = self._get_ipc_meta(inp)
        ops.register_buffer(self._ptr, inp, handles, offsets)

    def register_graph_buffers(self):
        handle, offset = ops.get_graph_buffer_ipc_meta(self._ptr)
        handles, offsets = self._gather_ipc_meta((bytes(handle.numpy().tobytes()), offset))
        ops.register_graph_buffers(self._ptr, handles, offsets)

    def should_custom_ar(self, inp: torch.Tensor) -> bool:
        if self.disabled:
            return False
        inp_size = inp.numel() * inp.element_size()
        if inp_size % 16 != 0 or not is_weak_contiguous(inp):
            return False
        if self.world_size == 2 or self.full_nvlink:
            return inp_size <= self.max_size      # the reference uses `<` (custom_all_reduce.py:250)
        return False

    def all_reduce_reg(self, inp: torch.Tensor, out: Optional[torch.Tensor] = None):
        if out is None:
            out = torch.empty_like(inp)
        ops.all_reduce_reg(self._ptr, inp, out)
        return out

    def all_reduce_unreg(self, inp: torch.Tensor, out: Optional[torch.Tensor] = None):
        if out is None:
            out = torch.empty_like(inp)
        ops.all_reduce_unreg(self._ptr, inp, self.buffer, out)
        return out

    def custom_all_reduce(self, input: torch.Tensor) -> Optional[torch.Tensor]:
        if self.disabled:
            return None
        if self._IS_CAPTURING:
            if torch.cuda.is_current_stream_capturing():
                if self.should_custom_ar(input):
                    return self.all_reduce_reg(input)
            else:
                if self.should_custom_ar(input):
                    return torch.empty_like(input)      # warm-up: mimic the allocation pattern
        else:
            if self.should_custom_ar(input):
                return self.all_reduce_unreg(input)
        return None

    def close(self):
        if not self.disabled and getattr(self, "_ptr", 0):
            ops.dispose(self._ptr)
            self._ptr = 0

    def __del__(self):
        self.close()
