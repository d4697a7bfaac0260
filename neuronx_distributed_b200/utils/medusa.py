"""Medusa decoding (reference ``utils/medusa_utils.py`` buffers + the Medusa path of ``examples/inference``): ``K`` extra heads on
the last hidden state guess the tokens at offsets +2 … +K+1; their top-k choices form a candidate TREE that the base model
verifies in one forward with a tree attention mask; the longest root→leaf path whose tokens match the base model's own greedy
predictions is accepted, the KV entries of that path are compacted into place, and the base model's prediction after the
last accepted node becomes the next root.  Greedy acceptance ⇒ the output equals plain greedy decoding token for token."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
from torch import nn

from ..parallel_layers.layers import ColumnParallelLinear
from .medusa_utils import generate_medusa_buffers


class MedusaHeads(nn.Module):
    """``num_heads`` × (residual SiLU block → vocab projection); the projection is column-parallel with gathered output."""

    def __init__(self, hidden_size: int, vocab_size: int, num_heads: int, dtype=torch.float32, device=None):
        super().__init__()
        self.blocks = nn.ModuleList([nn.Linear(hidden_size, hidden_size, dtype=dtype, device=device) for _ in range(num_heads)])
        self.proj = nn.ModuleList([ColumnParallelLinear(hidden_size, vocab_size, bias=False, gather_output=True, dtype=dtype, device=device)
                                   for _ in range(num_heads)])
        for b in self.blocks:
            nn.init.zeros_(b.weight)          # identity at init, as in the Medusa recipe
            nn.init.zeros_(b.bias)

    def forward(self, h: torch.Tensor) -> List[torch.Tensor]:
        return [p(h + torch.nn.functional.silu(b(h))) for b, p in zip(self.blocks, self.proj)]


@torch.no_grad()
def medusa_generate(model, heads: MedusaHeads, prompt_ids: torch.Tensor, max_new_tokens: int,
                    medusa_choices: Sequence[Sequence[int]], topk: int = 10) -> Tuple[torch.Tensor, float]:
    """``model``: :class:`LlamaForInference` (batch 1, greedy on-device sampling).  Returns (ids ``[1, n]``, mean accepted
    tree tokens per verification forward)."""
    assert prompt_ids.shape[0] == 1
    dev = prompt_ids.device
    buf = generate_medusa_buffers(medusa_choices, device=dev, topk=topk)
    tree_mask = buf["medusa_attn_mask"][0, 0].bool()                 # [W, W]
    tree_idx, depth, retrieve = buf["tree_indices"], buf["medusa_position_ids"], buf["retrieve_indices"]
    W = tree_idx.numel()
    S = prompt_ids.shape[1]
    # prefill: next token + hidden state of the last prompt position (input of the Medusa heads)
    h_full = model._body(prompt_ids, None, True, None)               # [S, 1, H]
    h_last = h_full[-1]                                              # [1, H]
    root = model.sampler.sample(model.lm.lm_head(h_last.unsqueeze(0))[0].float())      # [1]
    out: List[int] = [int(root)]
    pos = S                                                          # slot of the root token (not yet in the cache)
    accepted_total, rounds = 0, 0
    while len(out) < max_new_tokens:
        # candidates: [root, head0 top-k, head1 top-k, …] → tree node tokens
        flat = [root.view(1)]
        for lg in heads(h_last):
            flat.append(torch.topk(lg[0].float(), topk).indices)
        cand = torch.cat(flat)[tree_idx].view(1, W)
        pred, hid = model.speculation_forward(cand, torch.tensor([pos], device=dev), tree_mask=tree_mask, tree_depth=depth,
                                              return_hidden=True)    # pred [1, W]: base model's next token after every node
        pred, cand1 = pred[0], cand[0]
        best_path, best_len = retrieve[0], 0
        for path in retrieve:
            nodes = path[path >= 0]
            n_ok = 0
            while n_ok + 1 < nodes.numel() and cand1[nodes[n_ok + 1]] == pred[nodes[n_ok]]:
                n_ok += 1
            if n_ok > best_len or (n_ok == best_len and best_len == 0 and path is retrieve[0]):
                best_path, best_len = nodes, n_ok
        nodes = best_path[best_path >= 0][: best_len + 1]            # accepted nodes incl. the root
        last = nodes[-1]
        new = [int(cand1[i]) for i in nodes[1:]] + [int(pred[last])]
        out.extend(new)
        accepted_total += best_len
        rounds += 1
        # make the accepted path contiguous in the cache (root stays in slot pos; node i of the path moves to pos+i)
        model.kv.compact_window(torch.tensor([pos], device=dev), nodes.view(1, -1))
        h_last = hid[:, last]                                        # hidden state that produced the bonus token
        root = pred[last].view(1)
        pos += nodes.numel()
    return torch.tensor([out[:max_new_tokens]], device=dev), accepted_total / max(rounds, 1)
