"""Sequence-length buckets and the routers that pick one per request (role of the reference's
``examples/inference/modules/autobucketing.py``: ``generate_buckets``, ``context_encoder_bk``, ``token_generation_bk``).

A bucket is a sequence length a step function was captured for (one CUDA graph each).  Context encoding is routed by the
number of real (non-pad) tokens of the longest prompt in the batch, token generation by the largest cache position — the
smallest bucket that fits every sequence of the batch — and the padded ``[batch, seq]`` inputs are cut to the bucket from the
side opposite to the padding.  1-D inputs (sequence ids) pass through untouched.

The choice is a host decision (it selects which captured program to replay): with device tensors it costs one scalar readback per
call; pass CPU tensors / lengths to avoid it.
"""
from __future__ import annotations

import math
from typing import Callable, List, Sequence, Tuple

import torch


def generate_buckets(min_len: int, max_len: int) -> List[int]:
    """Powers of two from ``min_len`` up, then ``max_len`` itself.  A power of two closer than a factor √2 to ``max_len`` is
    dropped — two captured programs that close together waste memory for no latency gain (e.g. 128…512, 1100 — not 1024, 1100)."""
    if min_len >= max_len:
        return [max_len]
    out, b = [], 1 << max(0, int(math.log2(max(1, min_len))))
    while b * math.sqrt(2.0) <= max_len:
        out.append(b)
        b *= 2
    out.append(max_len)
    return out


def pick_bucket(buckets: Sequence[int], length: int) -> int:
    for b in buckets:
        if length <= b:
            return b
    raise ValueError(f"length {length} exceeds the largest bucket {buckets[-1]}")


def _bucket_index(buckets: Sequence[int], need: int) -> int:
    for i, b in enumerate(buckets):
        if need <= b:
            return i
    return len(buckets) - 1


def _cut(t: torch.Tensor, bucket: int, largest: int, padding_side: str) -> torch.Tensor:
    if t.dim() < 2 or t.shape[1] <= bucket:
        return t
    return t[:, :bucket] if padding_side == "right" else t[:, largest - bucket:largest] if t.shape[1] >= largest else t[:, -bucket:]


def context_encoding_router(tensors: List[torch.Tensor], buckets: Sequence[int], padding_side: str = "right",
                            pad_token: int = 0) -> Tuple[List[torch.Tensor], int]:
    """``tensors[0]`` = ``input_ids [B, S_max]``.  Returns the inputs cut to the chosen bucket and the bucket index."""
    assert padding_side in ("left", "right")
    ids = tensors[0]
    longest = int((ids != pad_token).sum(dim=1).max())
    idx = _bucket_index(buckets, longest)
    b = buckets[idx]
    return [_cut(t, b, buckets[-1], padding_side) for t in tensors], idx


def token_generation_router(tensors: List[torch.Tensor], buckets: Sequence[int], padding_side: str = "right"
                            ) -> Tuple[List[torch.Tensor], int]:
    """``tensors`` = ``[input_ids [B,1], (attention_mask [B,S_max],) position_ids [B,1], …]``.  The bucket must hold every
    sequence's NEXT position (``position + 1`` cache entries); only the attention mask is cut — ids / positions are ``[B,1]``."""
    assert padding_side in ("left", "right")
    has_mask = tensors[1].dim() == 2 and tensors[1].shape[1] != 1
    pos = tensors[2] if has_mask else tensors[1]
    idx = _bucket_index(buckets, int(pos.max()) + 1)
    out = list(tensors)
    if has_mask:
        out[1] = _cut(tensors[1], buckets[idx], buckets[-1], padding_side)
    return out, idx


def get_context_encoder_bk() -> Callable:
    return context_encoding_router


def get_token_generation_bk() -> Callable:
    return token_generation_router
