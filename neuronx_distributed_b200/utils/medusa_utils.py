"""Medusa speculative-decoding tree buffers (reference ``utils/medusa_utils.py``): from a list of candidate
paths (each a tuple of per-head top-k choices) build the tree attention mask, the index of every tree node in
the flattened candidate tensor, per-node position offsets and the retrieval indices of every root→leaf path."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

TOPK = 10


def pad_path(path: Sequence[int], length: int, pad_value: int = -2) -> List[int]:
    return list(path) + [pad_value] * (length - len(path))


def generate_medusa_buffers(medusa_choices: Sequence[Sequence[int]], device="cpu", topk: int = TOPK) -> Dict[str, torch.Tensor]:
    sorted_choices = sorted(medusa_choices, key=lambda x: (len(x), x))
    n = len(sorted_choices) + 1
    depth_counts: List[int] = []
    prev = 0
    for p in sorted_choices:
        if len(p) != prev:
            depth_counts.append(0)
            prev = len(p)
        depth_counts[-1] += 1
    attn = torch.eye(n, n)
    attn[:, 0] = 1
    start = 0
    for d, cnt in enumerate(depth_counts):
        for j in range(cnt):
            cur = sorted_choices[start + j]
            if len(cur) == 1:
                continue
            anc = [sorted_choices.index(cur[: c + 1]) + 1 for c in range(len(cur) - 1)]
            attn[start + j + 1, anc] = 1
        start += cnt
    tree_idx = torch.zeros(n, dtype=torch.long)
    start = 0
    for d, cnt in enumerate(depth_counts):
        for j in range(cnt):
            cur = sorted_choices[start + j]
            tree_idx[start + j + 1] = cur[-1] + topk * d + 1
        start += cnt
    pos = torch.zeros(n, dtype=torch.long)
    start = 0
    for d, cnt in enumerate(depth_counts):
        pos[start + 1 : start + cnt + 1] = d + 1
        start += cnt
    paths: List[List[int]] = []
    seen: List[Tuple[int, ...]] = []
    for cur in reversed(sorted_choices):
        cur = tuple(cur)
        if any(s[: len(cur)] == cur for s in seen):
            continue
        seen.append(cur)
        paths.append([0] + [sorted_choices.index(list(cur[: c + 1])) + 1 if list(cur[: c + 1]) in [list(x) for x in sorted_choices]
                            else -1 for c in range(len(cur))])
    maxlen = max(len(p) for p in paths)
    retrieve = torch.tensor([pad_path(p, maxlen, -1) for p in paths], dtype=torch.long)
    return {
        "medusa_attn_mask": attn.unsqueeze(0).unsqueeze(0).to(device),
        "tree_indices": tree_idx.to(device),
        "medusa_position_ids": pos.to(device),
        "retrieve_indices": retrieve.to(device),
    }
