"""Distributed argmax over a dimension sharded across TP (reference ``operators/argmax.py:55-152``).

local (max, argmax) → ONE all-gather of the packed ``[..., 2]`` (value, global index) pairs over the group →
final argmax of the ``tp`` candidates.  Ties resolve to the smallest global index, like ``torch.argmax`` on
the gathered tensor.  The reference's multi-stage "cascaded max" NKI kernel (K10) is ``csrc/select.cu::row_argmax_kernel``
here: one CTA per row, 16-byte loads, value + index in one pass (``torch.max`` on CPU / other dims)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from ..parallel_layers import comm
from ..parallel_layers import parallel_state as ps


def argmax(tensor: torch.Tensor, dim: int, gather_dim: Optional[int] = None, keepdim: bool = False,
           process_group=None, disable_argmax_kernel: bool = False, rank_id: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Positional order of the reference (argmax.py:106-113); ``disable_argmax_kernel`` forces ``torch.max`` for the local
    reduction instead of the one-pass row arg-max kernel."""
    group = process_group if process_group is not None else ps.get_tensor_model_parallel_group()
    n = dist.get_world_size(group)
    dim = dim % tensor.dim()
    if dim == tensor.dim() - 1 and tensor.is_cuda and not disable_argmax_kernel:
        from ..ops import select as _select           # one-pass row arg-max kernel (csrc/select.cu, role of NKI cascaded_max)

        val, idx = _select.row_max(tensor)
        val, idx = val.unsqueeze(dim).to(tensor.dtype), idx.unsqueeze(dim)
    else:
        val, idx = torch.max(tensor, dim=dim, keepdim=True)
    if n == 1:
        return idx if keepdim else idx.squeeze(dim)
    r = dist.get_rank(group) if rank_id is None else rank_id.reshape(-1)[0].to(idx.device)
    gidx = idx + r * tensor.shape[dim]
    packed = torch.stack([val.float(), gidx.float()], dim=-1)           # one collective for both
    allp = comm.all_gather(packed.unsqueeze(0), dim=0, group=group)     # [n, ..., 1, ..., 2]
    vals, gids = allp[..., 0], allp[..., 1].long()
    # max value; among equal values the lowest global index
    best = vals.max(dim=0, keepdim=True).values
    cand = torch.where(vals == best, gids, torch.full_like(gids, torch.iinfo(torch.long).max))
    out = cand.min(dim=0).values
    return out if keepdim else out.squeeze(dim)
