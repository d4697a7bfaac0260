"""Sequence-length buckets (role of reference ``examples/inference/modules/autobucketing.py:33-60``): powers of two
from 128 up to the maximum; the router picks the smallest bucket that fits."""
from __future__ import annotations

from typing import List


def generate_buckets(min_len: int, max_len: int) -> List[int]:
    if min_len >= max_len:
        return [max_len]
    out, b = [], max(1, min_len)
    while b < max_len:
        out.append(b)
        b *= 2
    out.append(max_len)
    return out


def pick_bucket(buckets: List[int], length: int) -> int:
    for b in buckets:
        if length <= b:
            return b
    raise ValueError(f"length {length} exceeds the largest bucket {buckets[-1]}")
