"""Context-parallel batch slicing (reference ``utils/batch_utils.py:4-44``): labels are shifted on the FULL
sequence first (so the token that crosses a slice boundary keeps its target), then every tensor with a sequence
dim is cut contiguously into ``cp`` slices and this rank keeps slice ``cp_rank``."""
from __future__ import annotations

from typing import Any, Dict

import torch

from ..parallel_layers import parallel_state as ps


def get_batch_on_this_context_parallel_rank(batch: Dict[str, Any], seq_dim: int = 1, shift_labels: bool = True,
                                            ignore_index: int = -100) -> Dict[str, Any]:
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
    for k, v in list(out.items()):
        if isinstance(v, torch.Tensor) and v.dim() > seq_dim and v.shape[seq_dim] % cp == 0:
            out[k] = v.chunk(cp, dim=seq_dim)[r].contiguous()
    return out


def shift_labels(batch: Dict[str, Any], label_pad_token_id: int = -100) -> Dict[str, Any]:
    """``labels[:, t] ← labels[:, t+1]`` with ``-100`` in the last column (reference :4-16): done BEFORE a sequence is split
    over context-parallel ranks, so the model must not shift again."""
    out = dict(batch)
    if "labels" in out:
        lab = out["labels"]
        out["labels"] = torch.cat([lab[:, 1:], lab.new_full((lab.shape[0], 1), label_pad_token_id)], dim=1)
    return out
