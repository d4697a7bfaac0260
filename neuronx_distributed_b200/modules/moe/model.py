"""The MoE layer (reference ``modules/moe/model.py:154-258``):
(rmsnorm) → (token shuffle) → SP all-gather → router → routed experts (+ shared experts) → ONE delayed
reduce-scatter / all-reduce over TP (over the world when EP is on) → (unshuffle)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from ...parallel_layers import mappings
from ...parallel_layers import parallel_state as ps
from . import token_shuffling


class MoE(nn.Module):
    def __init__(self, router: nn.Module, expert_mlps: nn.Module, shared_experts: Optional[nn.Module] = None,
                 rmsnorm: Optional[nn.Module] = None, sequence_parallel_enabled: bool = False,
                 sequence_dimension: Optional[int] = None, return_router_logits: bool = False,
                 return_expert_index: bool = False, token_shuffle_group_size: int = 1, token_shuffle_seed=None,
                 tensor_model_parallel_group=None, init_tkg_module: bool = False, tkg_config=None):
        super().__init__()
        self.router, self.expert_mlps, self.shared_experts, self.rmsnorm = router, expert_mlps, shared_experts, rmsnorm
        self.sequence_parallel_enabled = sequence_parallel_enabled
        self.sequence_dimension = 0 if sequence_dimension is None else sequence_dimension
        self.return_router_logits, self.return_expert_index = return_router_logits, return_expert_index
        self.token_shuffle_group_size, self.token_shuffle_seed = token_shuffle_group_size, token_shuffle_seed
        self.tensor_parallel_group = tensor_model_parallel_group
        if token_shuffle_group_size > 1:
            ps.initialize_token_shuffle_group(token_shuffle_group_size)
        self.moe_fused_tkg = None
        if init_tkg_module:
            from .moe_fused_tkg import MoEFusedTKG

            self.moe_fused_tkg = MoEFusedTKG(router, expert_mlps, shared_experts, rmsnorm, tkg_config,
                                             sequence_dimension=self.sequence_dimension,
                                             tensor_model_parallel_group=tensor_model_parallel_group,
                                             return_router_logits=return_router_logits, return_expert_index=return_expert_index)

    def _shared(self, full: torch.Tensor, seq_len: int) -> torch.Tensor:
        try:
            return self.shared_experts(full, seq_len)
        except TypeError:                                   # user-supplied module with the one-argument signature
            return self.shared_experts(full)

    def _reduce(self, y: torch.Tensor) -> torch.Tensor:
        ep = ps.get_expert_model_parallel_size()
        tp_group = self.tensor_parallel_group if self.tensor_parallel_group is not None else ps.get_tensor_model_parallel_group()
        if self.sequence_parallel_enabled:
            y = mappings.reduce_scatter_to_sequence_parallel_region(y, self.sequence_dimension, tp_group)
        else:
            y = mappings.reduce_from_tensor_model_parallel_region(y, tp_group)
        if ep > 1 and not self.training:
            y = mappings.reduce_from_tensor_model_parallel_region(y, ps.get_expert_model_parallel_group())
        return y

    def forward(self, hidden_states: torch.Tensor, padding_mask: Optional[torch.Tensor] = None,
                is_speculative_decoding: bool = False, residual: Optional[torch.Tensor] = None):
        """Returns ``(output, [router_logits], [expert_index], [residual])``.  ``residual`` (same layout as ``hidden_states``):
        the layer computes on ``hidden_states + residual`` and also returns that sum, the residual stream after the attention
        block (fused into the decode kernel on the token-generation path).  ``is_speculative_decoding``: a short verification
        window is served by the decode block as well (reference model.py:260-303)."""
        decode = hidden_states.shape[self.sequence_dimension] == 1 or is_speculative_decoding
        if self.moe_fused_tkg is not None and not self.training and decode:
            return self.moe_fused_tkg(hidden_states) if residual is None else self.moe_fused_tkg(hidden_states, residual=residual)
        if residual is not None:
            hidden_states = hidden_states + residual
            return self._forward_compute_bound(hidden_states, padding_mask) + (hidden_states,)
        return self._forward_compute_bound(hidden_states, padding_mask)

    def _forward_compute_bound(self, hidden_states: torch.Tensor, padding_mask: Optional[torch.Tensor] = None) -> Tuple:
        x = self.rmsnorm(hidden_states) if self.rmsnorm is not None else hidden_states
        perm = None
        if self.token_shuffle_group_size > 1:
            x, perm = token_shuffling.token_shuffle(x, self.token_shuffle_seed, self.sequence_dimension)
        router_logits, aff, idx = self.router(x)            # router gathers over SP itself
        full = x
        if self.sequence_parallel_enabled:
            full = mappings.gather_from_sequence_parallel_region(x, self.sequence_dimension, to_model_parallel=True,
                                                                 process_group=self.tensor_parallel_group)
        shape = full.shape
        seq_len = shape[self.sequence_dimension]
        y = self.expert_mlps(full.reshape(-1, shape[-1]), aff, idx, seq_len=seq_len, padding_mask=padding_mask).view(shape)
        shared_complete = None
        if self.shared_experts is not None:
            if (getattr(self.shared_experts, "sequence_parallel_enabled", False) and self.sequence_parallel_enabled
                    and seq_len > 1):
                # SP shared experts hold replicated weights: run on the LOCAL sequence shard, complete result, no collective;
                # added after the routed experts' reduce-scatter (reference shared_experts.py SP flow)
                shared_complete = self._shared(x, seq_len)
            else:
                y = y + self._shared(full, seq_len)
        y = self._reduce(y)
        if shared_complete is not None:
            y = y + shared_complete
        if perm is not None:
            y = token_shuffling.token_unshuffle(y, perm, self.sequence_dimension)
        out: Tuple = (y,)
        if self.return_router_logits:
            out += (router_logits,)
        if self.return_expert_index:
            out += (idx,)
        return out
