"""Symmetric-memory plumbing for the fused tensor-parallel exchange kernel (csrc/tp_fused.cu).

The reference's row-parallel layers end in `tensor_model_parallel_all_reduce` (aphrodite/distributed/communication_op.py,
GroupCoordinator.all_reduce aphrodite/distributed/parallel_state.py:353-379: custom all-reduce kernel or NCCL) and the
decoder layer then calls `fused_add_rms_norm` on the replicated result (aphrodite/modeling/models/llama.py:250-256). On an
NVSwitch box the two become one kernel (`torch.ops._C_b200.tp_allreduce_rows` -> `b200_tp_allreduce_rows`): the switch
sums the ranks' partial rows (`multimem.ld_reduce`), the owning rank adds the residual and normalises, and the switch
replicates the result into every rank's buffer (`multimem.st`).

This module owns only the MEMORY the kernel works on: one symmetric allocation per rank
    [ barrier flags | X: partial sums [max_tokens, hidden] | H: exchanged result [max_tokens, hidden] ]
obtained from torch's symmetric-memory allocator (cuMemCreate + peer mappings + a multicast object when the fabric has
one — the CUDA VMM plumbing; no torch collective runs on the data path). `x(T)` is where the row-parallel GEMM must
write its output, `allreduce_add_rms_norm` / `all_reduce` launch the kernel and return `h(T)`.
"""
from typing import Optional

import torch
import torch.distributed as dist

from .. import _native


# Symmetric allocations are process-lifetime objects: peers hold mappings of them and CUDA graphs bake their addresses
# in, and torch's allocator synchronises the device inside the tensor destructor (observed: std::terminate when a block
# was released while the process group was still alive). They are therefore never freed.
_KEEPALIVE = []


def _ops():
    _native.load_torch_ops()
    return torch.ops._C_b200


class NvlsTensorParallel:
    """One instance per (process group, max_tokens, hidden, dtype). Every rank must issue the same call sequence."""

    ALGOS = {"p2p": 0, "mc_store": 1, "mc_reduce": 2}       # B200_TP_* of include/b200_decode.h

    def __init__(self, group: dist.ProcessGroup, device: torch.device, max_tokens: int, hidden: int,
                 dtype: torch.dtype = torch.bfloat16, algo: str = "auto") -> None:
        """algo: "p2p" (unicast peer pointers only), "mc_store" (unicast fp32 rank-order reduce — bit-identical to the
        reference's all-reduce — + multicast store and flags), "mc_reduce" (the switch reduces; not bit-identical),
        "auto" = "mc_store" when the fabric gives a multicast mapping, else "p2p"."""
        import torch.distributed._symmetric_memory as symm_mem
        assert dtype in (torch.bfloat16, torch.float16), "fused TP exchange: float16 / bfloat16"
        assert hidden % 8 == 0
        self.group, self.device, self.dtype = group, torch.device(device), dtype
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.max_tokens, self.hidden = max_tokens, hidden
        esz = torch.tensor([], dtype=dtype).element_size()
        self.flag_bytes = int(_ops().tp_flag_bytes())
        self.x_off = (self.flag_bytes + 127) // 128 * 128
        self.h_off = self.x_off + (max_tokens * hidden * esz + 127) // 128 * 128
        nbytes = self.h_off + max_tokens * hidden * esz
        self.block = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
        hdl = symm_mem.rendezvous(self.block, group.group_name)
        self._hdl = hdl
        _KEEPALIVE.append((self.block, hdl))
        off = int(getattr(hdl, "offset", 0))
        self.peer_bases = [int(p) + off for p in hdl.buffer_ptrs]
        assert self.peer_bases[self.rank] == self.block.data_ptr(), "symmetric block is not at its advertised address"
        mc = int(hdl.multicast_ptr)
        self.has_multicast = mc != 0
        if algo == "auto":
            algo = "mc_store" if self.has_multicast else "p2p"
        assert algo in self.ALGOS, algo
        if algo != "p2p" and not self.has_multicast:
            raise RuntimeError(f"NvlsTensorParallel: algo {algo!r} needs a multicast mapping, this box has none")
        self.algo = algo
        self.mc_base = (mc + off) if self.has_multicast else 0
        self.multicast = algo != "p2p"
        # barrier counters start at zero on every rank before anybody signals
        self.block.zero_()
        torch.cuda.synchronize(self.device)
        dist.barrier(group)
        self._x = self.block[self.x_off:self.x_off + max_tokens * hidden * esz].view(dtype).view(max_tokens, hidden)
        self._h = self.block[self.h_off:self.h_off + max_tokens * hidden * esz].view(dtype).view(max_tokens, hidden)

    # ---- buffers ---------------------------------------------------------------------------------------------
    def x(self, num_tokens: int) -> torch.Tensor:
        """[num_tokens, hidden] view the row-parallel GEMM writes its partial sums into."""
        return self._x[:num_tokens]

    def h(self, num_tokens: int) -> torch.Tensor:
        """[num_tokens, hidden] view holding the exchanged result (identical on all ranks after a call)."""
        return self._h[:num_tokens]

    def rows_of(self, rank: int, num_tokens: int):
        """The residual rows rank `rank` owns: [lo, hi)."""
        per = (num_tokens + self.world - 1) // self.world
        return min(num_tokens, rank * per), min(num_tokens, (rank + 1) * per)

    # ---- the exchange ------------------------------------------------------------------------------------------
    def _launch(self, num_tokens: int, residual: Optional[torch.Tensor], weight: Optional[torch.Tensor], eps: float):
        _ops().tp_allreduce_rows(self.mc_base, self.block, self.peer_bases, self.x(num_tokens), self.h(num_tokens),
                                 residual, weight, eps, 0, self.rank, self.world, self.ALGOS[self.algo])
        return self.h(num_tokens)

    def allreduce_add_rms_norm(self, num_tokens: int, residual: torch.Tensor, weight: torch.Tensor,
                               eps: float) -> torch.Tensor:
        """H = rms_norm(sum_r X_r + residual) * weight; residual[own rows] <- sum + residual. Returns h(num_tokens)."""
        assert residual.dtype == self.dtype and residual.is_contiguous() and residual.shape[-1] == self.hidden
        return self._launch(num_tokens, residual, weight, eps)

    def all_reduce(self, num_tokens: int) -> torch.Tensor:
        """H = sum_r X_r. Returns h(num_tokens)."""
        return self._launch(num_tokens, None, None, 0.0)

    def gather_residual(self, residual: torch.Tensor, num_tokens: int) -> torch.Tensor:
        """Test / debug helper: the full residual stream assembled from the ranks' own rows (NCCL all-gather)."""
        per = (num_tokens + self.world - 1) // self.world
        mine = torch.zeros(per, self.hidden, dtype=self.dtype, device=self.device)
        lo, hi = self.rows_of(self.rank, num_tokens)
        mine[:hi - lo] = residual[lo:hi]
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        return torch.cat(parts, dim=0)[:num_tokens]
