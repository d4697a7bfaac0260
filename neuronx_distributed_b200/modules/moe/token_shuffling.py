"""BASE-layers style token shuffling (reference ``modules/moe/token_shuffling.py``): a random permutation followed
by an all-to-all over the token-shuffle group decorrelates the tokens each DP rank feeds its routers; ``unshuffle``
is the exact inverse."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ...parallel_layers import mappings
from ...parallel_layers import parallel_state as ps


def all_to_all_for_shuffle(hidden_states: torch.Tensor, input_is_sequence_parallel: bool = True, dim: int = 0) -> torch.Tensor:
    """The exchange step on its own (reference token_shuffling.py:90-99): all-to-all over the token-shuffle group along
    ``dim`` — a self-inverse permutation of equal chunks, differentiable.  With ``input_is_sequence_parallel=False`` the
    (replicated) input is first split over the TP group and the result gathered back."""
    if not input_is_sequence_parallel:
        hidden_states = mappings.scatter_to_sequence_parallel_region(hidden_states, dim)
    out = mappings.all_to_all_in_expert_parallel_region(hidden_states, dim, dim, ps.get_token_shuffle_group())
    if not input_is_sequence_parallel:
        out = mappings.gather_from_sequence_parallel_region(out, dim, to_model_parallel=False)
    return out


def token_shuffle(hidden_states: torch.Tensor, seed: Optional[int] = None, dim: int = 0
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    T = hidden_states.shape[dim]
    g = torch.Generator(device="cpu")
    if seed is not None:
        g.manual_seed(seed)
    perm = torch.randperm(T, generator=g).to(hidden_states.device)
    x = hidden_states.index_select(dim, perm)
    x = mappings.all_to_all_in_expert_parallel_region(x, dim, dim, ps.get_token_shuffle_group())
    return x, perm


def token_unshuffle(permuted_states: torch.Tensor, permutation: torch.Tensor, dim: int = 0) -> torch.Tensor:
    hidden_states = permuted_states      # reference parameter names in the signature
    x = mappings.all_to_all_in_expert_parallel_region(hidden_states, dim, dim, ps.get_token_shuffle_group())
    inv = torch.empty_like(permutation)
    inv[permutation] = torch.arange(permutation.numel(), device=permutation.device)
    return x.index_select(dim, inv)
