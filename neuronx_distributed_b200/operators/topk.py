"""Distributed top-k over a dimension sharded across TP (reference ``operators/topk.py:31-148``):
local top-k → all-gather of (values, global indices) → top-k of the ``k·tp`` candidates."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from ..parallel_layers import comm
from ..parallel_layers import parallel_state as ps


def topk(tensor: torch.Tensor, k: int, dim: int, gather_dim: Optional[int] = None, process_group=None, stages: int = 1,
         rank_id: Optional[torch.Tensor] = None, use_topk_rotated_kernel: bool = False, lnc: int = 2
         ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Positional order of the reference (topk.py:31-42).  ``use_topk_rotated_kernel`` / ``lnc`` select between two device
    kernels there; here one selection kernel serves every case, so they are accepted and ignored."""
    group = process_group if process_group is not None else ps.get_tensor_model_parallel_group()
    n = dist.get_world_size(group)
    dim = dim % tensor.dim()
    kk = min(k, tensor.shape[dim])
    if dim == tensor.dim() - 1 and tensor.is_cuda:
        from ..ops import select as _select           # smem-staged k-pass selection kernel (csrc/select.cu, role of nkilib topk)

        vals, idx = _select.row_topk(tensor, kk)
        vals = vals.to(tensor.dtype)
    else:
        vals, idx = torch.topk(tensor, kk, dim=dim)
    if n == 1:
        return vals, idx
    r = dist.get_rank(group) if rank_id is None else rank_id.reshape(-1)[0].to(idx.device)
    gidx = idx + r * tensor.shape[dim]
    allv = comm.all_gather(vals, dim=dim, group=group)
    alli = comm.all_gather(gidx, dim=dim, group=group)
    fv, pos = torch.topk(allv, k, dim=dim)
    return fv, torch.gather(alli, dim, pos)


def get_topk_implementation(use_topk_rotated_kernel: bool = False, lnc: int = 1, stages: int = 1):
    """``(topk_unsorted, topk_sorted, stages)`` — the local top-k used inside the distributed one (reference :14-28).
    ``torch.topk`` on CUDA is a radix-select kernel and is used for both; ``use_topk_rotated_kernel`` / ``lnc`` select an
    NKI kernel variant in the reference and are accepted for compatibility."""
    if use_topk_rotated_kernel:
        assert stages == 1, "stages other than 1 is not supported when using topk_rotated kernel"

    def topk_unsorted(t: torch.Tensor, k: int, dim: Optional[int] = None):
        return torch.topk(t, k, dim=-1 if dim is None else dim, sorted=False)

    def topk_sorted(t: torch.Tensor, k: int, dim: Optional[int] = None):
        return torch.topk(t, k, dim=-1 if dim is None else dim, sorted=True)

    return topk_unsorted, topk_sorted, stages
