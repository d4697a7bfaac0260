"""Routed expert MLPs — dispatch over execution modes (reference ``modules/moe/expert_mlps_v2.py:1407-1500``):

* **all-experts**: every local expert processes every token, results masked/weighted — best for tiny token counts
  (decode) where the weights, not the activations, dominate memory traffic;
* **capacity-factor**: each expert takes at most ``C = ceil(cf·T·k/E)`` tokens in routing order, the rest are
  dropped (position-in-expert via cumsum, reference :484-593); batched ``[E, C, H]`` GEMMs;
* **blockwise (dropless)**: :mod:`.blockwise`;
* **expert parallel** (training): capacity-factor layout + all-to-all dispatch/combine over the EP group
  (``enter/exit_expert_parallel_region``, reference experts.py:174-214); inference masks to local experts instead.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from ...parallel_layers import mappings
from ...parallel_layers import parallel_state as ps
from .blockwise import blockwise_expert_mlp
from .experts import Experts
from .moe_configs import BlockwiseMatmulConfig, RoutedExpertsMLPOpsConfig


class ExpertMLPsV2(nn.Module):
    def __init__(self, routed_experts_mlp_config: RoutedExpertsMLPOpsConfig,
                 blockwise_matmul_config: Optional[BlockwiseMatmulConfig] = None, sequence_parallel_enabled: bool = False,
                 dtype: torch.dtype = torch.float32, device=None, tensor_model_parallel_group=None,
                 expert_model_parallel_group=None, is_prefill: bool = True, return_bias: bool = False):
        super().__init__()
        c = routed_experts_mlp_config
        self.cfg = c
        self.bw = blockwise_matmul_config or BlockwiseMatmulConfig.default()
        self.num_experts, self.top_k = c.num_experts, c.top_k
        self.capacity_factor = c.capacity_factor
        self.sequence_parallel_enabled = sequence_parallel_enabled
        self.ep = ps.get_expert_model_parallel_size() if ps.model_parallel_is_initialized() else 1
        self.mlp_op = Experts(
            c.num_experts, c.hidden_size, c.intermediate_size, c.hidden_act, c.glu_mlp, c.glu_type, c.capacity_factor,
            reduce_output=False, dtype=dtype, device=device, input_layer_init_method=c.input_layer_init_method,
            output_layer_init_method=c.output_layer_init_method, tensor_model_parallel_group=tensor_model_parallel_group,
            hidden_act_scaling_factor=c.hidden_act_scaling_factor, hidden_act_bias=c.hidden_act_bias,
            gate_clamp_upper_limit=c.gate_clamp_upper_limit, gate_clamp_lower_limit=c.gate_clamp_lower_limit,
            up_clamp_upper_limit=c.up_clamp_upper_limit, up_clamp_lower_limit=c.up_clamp_lower_limit)
        self.local_expert_ids = self.mlp_op.down_proj.local_expert_ids

    # ------------------------------------------------------------------ helpers
    def _topk_affinities(self, aff: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """[T, E] affinities keeping only the chosen experts (optionally re-normalised over them)."""
        mask = torch.zeros_like(aff).scatter_(1, idx, 1.0)
        a = aff * mask
        if self.cfg.normalize_top_k_affinities:
            a = a / a.sum(-1, keepdim=True).clamp(min=1e-9)
        return a

    # ------------------------------------------------------------------ modes
    def forward_all_experts(self, x: torch.Tensor, aff: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        a = self._topk_affinities(aff, idx)                                   # [T, E]
        local = torch.as_tensor(self.local_expert_ids, device=x.device)
        xe = x.unsqueeze(0).expand(len(self.local_expert_ids), *x.shape)      # [E_l, T, H]
        y = self.mlp_op(xe)                                                   # [E_l, T, H]
        w = a[:, local].t().unsqueeze(-1).to(y.dtype)                         # [E_l, T, 1]
        return (y * w).sum(0)

    def forward_capacity_factor(self, x: torch.Tensor, aff: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        T, H = x.shape
        E, k = self.num_experts, self.top_k
        C = min(T, max(1, math.ceil(self.capacity_factor * T * k / E)))
        a = self._topk_affinities(aff, idx)
        # position of every (token, slot) inside its expert's queue, in slot-major routing order
        onehot = torch.zeros(k, T, E, dtype=torch.long, device=x.device)
        onehot.scatter_(2, idx.t().unsqueeze(-1), 1)
        flat = onehot.reshape(k * T, E)
        pos = (torch.cumsum(flat, dim=0) - 1) * flat                           # [k·T, E]
        keep = (flat == 1) & (pos < C)
        pos = pos.reshape(k, T, E)
        keep = keep.reshape(k, T, E)
        keep_te = keep.any(0)                                                  # [T, E]
        pos_te = (pos * keep.long()).sum(0)                                    # [T, E]
        # dispatch: [E, C, H]
        disp = torch.zeros(E, C, H, dtype=x.dtype, device=x.device)
        t_idx, e_idx = keep_te.nonzero(as_tuple=True)
        disp[e_idx, pos_te[t_idx, e_idx]] = x[t_idx]
        if self.ep > 1 and self.training:
            d = mappings.enter_expert_parallel_region(disp, scatter_gather=False)        # [E/ep, ep, C, H]
            e_l, ep, _, _ = d.shape
            y = self.mlp_op(d.reshape(e_l, ep * C, H)).reshape(e_l, ep, C, H)
            y = mappings.exit_expert_parallel_region(y, scatter_gather=False)            # [E, C, H]
        else:
            local = torch.as_tensor(self.local_expert_ids, device=x.device)
            y_l = self.mlp_op(disp[local])
            y = torch.zeros(E, C, H, dtype=y_l.dtype, device=x.device)
            y[local] = y_l
        out = torch.zeros(T, H, dtype=y.dtype, device=x.device)
        out.index_add_(0, t_idx, y[e_idx, pos_te[t_idx, e_idx]] * a[t_idx, e_idx].unsqueeze(-1).to(y.dtype))
        return out

    def forward_blockwise(self, x: torch.Tensor, aff: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        a = self._topk_affinities(aff, idx)
        if self.ep > 1:
            # inference-style EP: mask routing to the experts this rank owns, results are summed over EP by the caller
            local = torch.as_tensor(self.local_expert_ids, device=x.device)
            remap = torch.full((self.num_experts,), -1, dtype=torch.long, device=x.device)
            remap[local] = torch.arange(len(self.local_expert_ids), device=x.device)
            return self.forward_all_experts(x, aff, idx)
        return blockwise_expert_mlp(x, a, idx, self.mlp_op, min(self.bw.block_size, max(16, x.shape[0])))

    def forward(self, hidden_states: torch.Tensor, expert_affinities: torch.Tensor, expert_index: torch.Tensor,
                seq_len: Optional[int] = None, padding_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """hidden ``[T, H]`` (already gathered over SP), affinities ``[T, E]``, index ``[T, k]`` → ``[T, H]`` partial sums
        over TP (and EP) — the MoE layer performs the delayed reduction."""
        x = hidden_states.reshape(-1, hidden_states.shape[-1])
        T = x.shape[0]
        if self.capacity_factor is not None and self.capacity_factor > 0 and (self.training or self.ep == 1):
            out = self.forward_capacity_factor(x, expert_affinities, expert_index)
        elif T * self.top_k <= self.num_experts * 4 or (self.ep > 1 and not self.training):
            out = self.forward_all_experts(x, expert_affinities, expert_index)
        else:
            out = self.forward_blockwise(x, expert_affinities, expert_index)
        return out.view(hidden_states.shape)


ExpertMLPs = ExpertMLPsV2
