"""ZeRO-1 gradient reduce-scatter / parameter all-gather on the peer-memory kernels (``csrc/zero1_comm.cu``).

The optimizer's flat parameter and gradient buffers are carved out of ONE symmetric allocation per sharding group, so
every rank can read its shard of every peer's gradients (pull reduce-scatter, fused with the fp32 accumulate and the 1/dp
scale) and write its updated shard into every peer's parameter buffer (push all-gather, fused with the fp32→bf16 cast).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _ext, symm

_ENABLED = os.environ.get("NXD_ZERO1_FUSED", "1") == "1"


def available(pg) -> bool:
    e = _ext.ext()
    return (_ENABLED and e is not None and hasattr(e, "zero1_reduce_scatter") and symm.available()
            and dist.get_backend(pg) == "nccl" and 1 < dist.get_world_size(pg) <= 8)


class Zero1Symm:
    """Symmetric arena for one optimizer: ``alloc`` hands out torch views; ``offset_of`` gives the byte offset that the
    kernels add to every peer's base pointer."""

    _count = 0

    def __init__(self, pg, nbytes: int):
        Zero1Symm._count += 1
        self.pg = pg
        self.rank, self.world = dist.get_rank(pg), dist.get_world_size(pg)
        self.ws = symm.get_workspace(pg, f"zero1_{Zero1Symm._count}", nbytes + (1 << 20), nflags=1024)
        self.cursor = 0
        self.epoch = 0
        self.done = torch.zeros(8, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))

    def alloc(self, numel: int, dtype: torch.dtype) -> Tuple[torch.Tensor, int]:
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        off = (self.cursor + 255) & ~255
        self.cursor = off + nbytes
        assert self.cursor <= self.ws.nbytes
        t = self.ws.local_tensor(off, (numel,), dtype)
        t.zero_()
        return t, off

    def reduce_scatter(self, grad_off: int, grad_dtype: torch.dtype, shard_numel: int, scale: float, out: torch.Tensor,
                       group_idx: int, sub_begin: int = 0, sub_len: Optional[int] = None, max_ctas: int = 0) -> torch.Tensor:
        """Reduce ``[sub_begin, sub_begin+sub_len)`` of this rank's shard (default: all of it).  Every rank of the group
        must make the same sequence of calls (a rank outside the bucket passes ``sub_len=0``)."""
        self.epoch += 1
        _ext.count_launch()
        _ext.ext().zero1_reduce_scatter(self.ws.ptrs, grad_off, self.ws.flag_ptrs, 32 * group_idx, self.epoch, self.rank,
                                        self.world, shard_numel, scale, out, self.done, grad_dtype == torch.float32,
                                        sub_begin, shard_numel if sub_len is None else sub_len, max_ctas)
        return out

    def all_gather(self, master: torch.Tensor, param_off: int, param_dtype: torch.dtype, shard_numel: int, group_idx: int) -> None:
        self.epoch += 1
        _ext.count_launch()
        _ext.ext().zero1_all_gather(master, self.ws.ptrs, param_off, self.ws.flag_ptrs, 32 * group_idx, self.epoch, self.rank,
                                    self.world, shard_numel, self.done, param_dtype == torch.bfloat16)
