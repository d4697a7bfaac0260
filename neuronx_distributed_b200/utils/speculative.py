"""Greedy speculative decoding with a small draft model (reference ``examples/inference/run_llama_speculative.py`` and the
``speculation_length`` path of its model wrapper): the draft proposes ``k`` tokens one by one, the target verifies the
whole window in ONE forward (``speculation_forward``) and accepts the longest matching prefix plus its own next token —
the output is token-for-token identical to the target's greedy decoding."""
from __future__ import annotations

from typing import Tuple

import torch


@torch.no_grad()
def speculative_generate(target, draft, prompt_ids: torch.Tensor, max_new_tokens: int, speculation_length: int = 4
                         ) -> Tuple[torch.Tensor, float]:
    """``target`` / ``draft``: :class:`models.llama_inference.LlamaForInference` with batch size 1 (on-device greedy
    sampling).  Returns ``(generated ids [1, n], mean accepted draft tokens per target forward)``."""
    assert prompt_ids.shape[0] == 1, "speculative decoding example runs batch 1"
    dev = prompt_ids.device
    S = prompt_ids.shape[1]
    t_next = target.context_encoding(prompt_ids)              # token at position S
    draft.context_encoding(prompt_ids)
    out = [int(t_next)]
    pos = S                                                   # position of the last emitted token (t_next)
    accepted_total, rounds = 0, 0
    while len(out) < max_new_tokens:
        k = min(speculation_length, max_new_tokens - len(out))
        # draft proposes k tokens after t_next
        proposal, tok, p = [], torch.tensor([[out[-1]]], device=dev), pos
        for _ in range(k):
            tok = draft.token_generation(tok.view(1, 1), torch.tensor([p], device=dev)).view(1, 1)
            proposal.append(int(tok))
            p += 1
        # target verifies [t_next, proposal...] in one forward: prediction after each of the k+1 positions
        window = torch.tensor([[out[-1]] + proposal], device=dev)
        pred = target.speculation_forward(window, torch.tensor([pos], device=dev))[0].tolist()     # k+1 tokens
        n_ok = 0
        while n_ok < k and proposal[n_ok] == pred[n_ok]:
            n_ok += 1
        new = proposal[:n_ok] + [pred[n_ok]]
        out.extend(new)
        accepted_total += n_ok
        rounds += 1
        # draft cache: positions pos..pos+k-1 are written ([last token, proposals[:-1]]); rejected entries get overwritten by
        # the next round, but when everything was accepted the last proposal (position pos+k) was never fed to the draft
        if n_ok == k:
            draft.token_generation(torch.tensor([[proposal[-1]]], device=dev), torch.tensor([pos + k], device=dev))
        pos += len(new)
    return torch.tensor([out[:max_new_tokens]], device=dev), accepted_total / max(rounds, 1)
