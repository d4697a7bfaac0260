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
    sorted_choices = sorted((tuple(c) for c in medusa_choices), key=lambda x: (len(x), x))
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
        paths.append([0] + [sorted_choices.index(cur[: c + 1]) + 1 if cur[: c + 1] in sorted_choices else -1
                            for c in range(len(cur))])
    maxlen = max(len(p) for p in paths)
    retrieve = torch.tensor([pad_path(p, maxlen, -1) for p in paths], dtype=torch.long)
    return {
        "medusa_attn_mask": attn.unsqueeze(0).unsqueeze(0).to(device),
        "tree_indices": tree_idx.to(device),
        "medusa_position_ids": pos.to(device),
        "retrieve_indices": retrieve.to(device),
    }


# ---------------------------------------------------------------------------------------------------------------------
# Per-step helpers of the Medusa loop (reference utils/medusa_utils.py:120-222).  The model's outputs here are token ids
# (the sampling / top-k runs on device inside the captured program), not probabilities:
#   ``logits``         – ``[..., ≥1]`` ids predicted by the LM head at the last accepted position (entry 0 = greedy token);
#   ``medusa_logits``  – ``[num_heads, ..., topk]`` top-k ids of every Medusa head.
# ---------------------------------------------------------------------------------------------------------------------
def generate_candidates(medusa_logits: torch.Tensor, logits: torch.Tensor, tree_indices: torch.Tensor,
                        retrieve_indices: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Lay the greedy token and the heads' top-k ids out on the tree.

    Returns ``(cart_candidates [num_paths, depth+1], tree_candidates [1, num_nodes])``: every root→leaf path as a row of
    token ids (``retrieve_indices == -1`` pads with token 0) and the flat node tokens that are fed to the tree-attention
    verification pass."""
    root = logits.reshape(-1)[:1]
    heads = medusa_logits.reshape(medusa_logits.shape[0], -1)                    # [num_heads, topk]
    flat = torch.cat([root, heads.reshape(-1)])                                  # index 0 = root, 1 + h·topk + j = head h choice j
    tree = flat[tree_indices]
    padded = torch.cat([tree, tree.new_zeros(1)])                                # slot -1 → 0
    return padded[retrieve_indices], tree.unsqueeze(0)


def evaluate_posterior(logits: torch.Tensor, candidates: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Greedy acceptance: a path is accepted up to the first position where its token differs from what the model
    predicted after the previous token.  ``logits [num_paths, depth+1, ≥1]`` are the verified ids along each path.
    Returns ``(best_candidate, accept_length)`` (0-d tensors; path 0 when nothing beyond the root is accepted)."""
    predicted = logits[:, :-1, 0].long()
    match = (candidates[:, 1:] == predicted).int()
    accepted = torch.cumprod(match, dim=1).sum(dim=1)
    accept_length = accepted.max()
    best = torch.where(accept_length == 0, torch.zeros_like(accept_length), torch.argmax(accepted))
    return best.to(torch.long), accept_length


def update_inference_inputs(input_ids: torch.Tensor, candidates: torch.Tensor, best_candidate: torch.Tensor,
                            accept_length: torch.Tensor, retrieve_indices: torch.Tensor, outputs, logits: torch.Tensor,
                            medusa_logits: torch.Tensor, new_token: int):
    """Commit the accepted prefix of the best path: append its tokens to ``input_ids`` and select, from the verification
    outputs, the LM / Medusa predictions at the last accepted node (the next step's roots).  ``select_indices`` are the
    absolute KV-cache positions of the accepted tree nodes (to be compacted to the front of the window)."""
    n = int(accept_length) + 1
    best = int(best_candidate)
    select_indices = retrieve_indices[best, :n] + input_ids.shape[1]
    input_ids = torch.cat([input_ids, candidates[None, best, :n]], dim=-1)
    logits = logits[None, best, n - 1:n]
    medusa_logits = medusa_logits[:, None, best, n - 1:n]
    return input_ids, logits, medusa_logits, new_token + n, select_indices
