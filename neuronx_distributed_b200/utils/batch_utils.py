"""Context-parallel batch slicing (reference ``utils/batch_utils.py:4-44``): labels are shifted on the FULL
sequence first (so the token that crosses a slice boundary keeps its target), then every tensor with a sequence
dim is cut contiguously into ``cp`` slices and this rank keeps slice ``cp_rank``."""
from __future__ import annotations

from typing import Any, Dict

import torch

from ..parallel_layers import parallel_state as ps


def context_parallel_slice(t: torch.Tensor, rank: int, cp: int, seq_dim: int = 1, layout: str = "contiguous") -> torch.Tensor:
    """This rank's part of a full-sequence tensor.  ``contiguous``: chunk ``rank`` of ``cp``.  ``zigzag``: chunks ``rank`` and
    ``2·cp − 1 − rank`` of ``2·cp`` (balanced causal attention, see ``modules/attention/ring.py``)."""
    if layout == "contiguous":
        return t.chunk(cp, dim=seq_dim)[rank].contiguous()
    if layout == "zigzag":
        parts = t.chunk(2 * cp, dim=seq_dim)
        return torch.cat([parts[rank], parts[2 * cp - 1 - rank]], dim=seq_dim).contiguous()
    raise ValueError(f"unknown context-parallel layout {layout!r}")


def get_batch_on_this_context_parallel_rank(batch: Dict[str, Any], seq_dim: int = 1, shift_labels: bool = True,
                                            ignore_index: int = -100, layout: str = "contiguous") -> Dict[str, Any]:
    cp = ps.get_context_model_parallel_size()
    out = dict(batch)
    if shift_labels and "labels" in out and out["labels"] is not None:
        lab = out["labels"]
        shifted = torch.full_like(lab, ignore_index)
        idx = [slice(None)] * lab.dim()
        src = list(idx); dst = list(idx)
        src[seq_dim] = slice(1, None); dst[seq_dim] = slice(0, -1)
        shifted[tuple(dst)] = lab[tuple(src)]
        out["labels"] = shifted
    if cp == 1:
        return out
    r = ps.get_context_model_parallel_rank()
    div = cp if layout == "contiguous" else 2 * cp
    for k, v in list(out.items()):
        if isinstance(v, torch.Tensor) and v.dim() > seq_dim and v.shape[seq_dim] % div == 0:
            out[k] = context_parallel_slice(v, r, cp, seq_dim, layout)
    return out


def shift_labels(batch: Dict[str, Any], label_pad_token_id: int = -100) -> Dict[str, Any]:
    """``labels[:, t] ← labels[:, t+1]`` with ``-100`` in the last column (reference :4-16): done BEFORE a sequence is split
    over context-parallel ranks, so the model must not shift again."""
    out = dict(batch)
    if "labels" in out:
        lab = out["labels"]
        out["labels"] = torch.cat([lab[:, 1:], lab.new_full((lab.shape[0], 1), label_pad_token_id)], dim=1)
    return out
