"""Auxiliary load-balancing loss (reference ``modules/moe/loss_function.py:5``): Switch-Transformer form,
``E · Σ_e f_e · P_e`` with f_e the fraction of (token, slot) assignments routed to expert e and P_e the mean
router probability of expert e."""
from __future__ import annotations

from typing import Sequence, Union

import torch
import torch.nn.functional as F


def load_balancing_loss_func(router_logits: Union[torch.Tensor, Sequence[torch.Tensor]], num_experts: int,
                             top_k: int) -> torch.Tensor:
    concatenated_gate_logits = router_logits      # reference parameter names in the signature
    if not isinstance(concatenated_gate_logits, torch.Tensor):
        concatenated_gate_logits = torch.cat([g.reshape(-1, num_experts) for g in concatenated_gate_logits], dim=0)
    logits = concatenated_gate_logits.reshape(-1, num_experts).float()
    probs = torch.softmax(logits, dim=-1)
    _, selected = torch.topk(probs, top_k, dim=-1)
    mask = F.one_hot(selected, num_experts).float()          # [T, k, E]
    tokens_per_expert = mask.mean(dim=0)                      # [k, E]
    router_prob_per_expert = probs.mean(dim=0)                # [E]
    return (tokens_per_expert * router_prob_per_expert.unsqueeze(0)).sum() * num_experts
