"""Decode-time fused MoE block (role of reference ``moe_fused_tkg.py:85`` / K8 ``moe_block_tkg``): for a handful of
tokens the step is weight-bandwidth bound, so RMSNorm → router → top-k → all-local-experts GLU → shared experts run
back-to-back on one stream with no collectives in between and a single reduction at the end; the sequence is
CUDA-graph capturable (static shapes, no host sync) and is what the inference runtime captures per decode bucket."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from ...parallel_layers import mappings
from .moe_configs import MoEFusedTKGConfig


class MoEFusedTKG(nn.Module):
    def __init__(self, router: nn.Module, expert_mlps: nn.Module, shared_experts: Optional[nn.Module] = None,
                 rmsnorm: Optional[nn.Module] = None, config: Optional[MoEFusedTKGConfig] = None):
        super().__init__()
        # references only (no new parameters): reuse the prefill modules' weights
        object.__setattr__(self, "_router", router)
        object.__setattr__(self, "_experts", expert_mlps)
        object.__setattr__(self, "_shared", shared_experts)
        object.__setattr__(self, "_norm", rmsnorm)
        self.config = config or MoEFusedTKGConfig()

    def preshard_hook(self, model_state_dict, prefix: str) -> None:
        """Nothing to re-arrange: the block holds references to the prefill modules' parameters (whose own hooks run) and
        owns none itself (reference ``moe_fused_tkg.py:449-450``)."""

    def forward(self, hidden_states: torch.Tensor, residual: Optional[torch.Tensor] = None):
        x = hidden_states if residual is None else hidden_states + residual
        h = self._norm(x) if self._norm is not None else x
        _, aff, idx = self._router(h)
        flat = h.reshape(-1, h.shape[-1])
        y = self._experts.forward_all_experts(flat, aff, idx).view(h.shape)
        if self._shared is not None:
            y = y + self._shared(h)
        y = mappings.reduce_from_tensor_model_parallel_region(y)
        return (y,) if residual is None else (y, x)
