"""Expert MLP weights + activation (reference ``modules/moe/experts.py``): fused gate|up column projection,
GLU / SwiGLU-with-clamps activation, row down-projection."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .model_utils import ACT2FN  # noqa: F401  (single activation table for the MoE package)
from .moe_parallel_layers import ExpertFusedColumnParallelLinear, ExpertFusedRowParallelLinear


class Experts(nn.Module):
    def __init__(self, num_experts: int, hidden_size: int, intermediate_size: int, hidden_act: str = "silu",
                 glu_mlp: bool = True, glu_type: str = "glu", capacity_factor=None, reduce_output: bool = False,
                 dtype=torch.float32, device=None, input_layer_init_method=None, output_layer_init_method=None,
                 tensor_model_parallel_group=None, hidden_act_scaling_factor: float = 1.0, hidden_act_bias: float = 0.0,
                 gate_clamp_upper_limit=None, gate_clamp_lower_limit=None, up_clamp_upper_limit=None,
                 up_clamp_lower_limit=None, bias: bool = False, expert_model_parallel_group=None, is_prefill: bool = True,
                 glu: Optional[bool] = None, activation_fn=None, expert_distribution=None):
        """``glu`` / ``activation_fn`` (a callable) are the reference's names for ``glu_mlp`` / the activation (experts.py:24-50);
        ``hidden_act`` may itself be a callable.  ``expert_distribution`` is forwarded to the expert-fused linears' placement."""
        super().__init__()
        if glu is not None:
            glu_mlp = bool(glu)
        if activation_fn is None and callable(hidden_act):
            activation_fn, hidden_act = hidden_act, getattr(hidden_act, "__name__", "custom")
        glu_type = getattr(glu_type, "value", glu_type)
        glu_type = glu_type.lower() if isinstance(glu_type, str) else glu_type
        self.glu_mlp, self.glu_type = glu_mlp, glu_type
        self.act = activation_fn if activation_fn is not None else ACT2FN[hidden_act]
        self.expert_distribution = expert_distribution
        self.scale, self.act_bias = hidden_act_scaling_factor, hidden_act_bias
        self.gc_hi, self.gc_lo, self.uc_hi, self.uc_lo = (gate_clamp_upper_limit, gate_clamp_lower_limit,
                                                          up_clamp_upper_limit, up_clamp_lower_limit)
        if glu_mlp:
            self.gate_up_proj = ExpertFusedColumnParallelLinear(
                num_experts, hidden_size, 2 * intermediate_size, dtype=dtype, device=device, stride=2,
                init_method=input_layer_init_method, tensor_model_parallel_group=tensor_model_parallel_group, bias=bias,
                expert_model_parallel_group=expert_model_parallel_group, is_prefill=is_prefill, is_fused_gate_up=True)
        else:
            self.up_proj = ExpertFusedColumnParallelLinear(
                num_experts, hidden_size, intermediate_size, dtype=dtype, device=device,
                init_method=input_layer_init_method, tensor_model_parallel_group=tensor_model_parallel_group, bias=bias,
                expert_model_parallel_group=expert_model_parallel_group, is_prefill=is_prefill)
        self.down_proj = ExpertFusedRowParallelLinear(
            num_experts, intermediate_size, hidden_size, reduce_output=reduce_output, dtype=dtype, device=device,
            init_method=output_layer_init_method, tensor_model_parallel_group=tensor_model_parallel_group, bias=bias,
            expert_model_parallel_group=expert_model_parallel_group, is_prefill=is_prefill)

        self.local_expert_ids = self.down_proj.local_expert_ids

    def activation(self, h: torch.Tensor) -> torch.Tensor:
        if not self.glu_mlp:
            return self.act(h)
        gate, up = h.chunk(2, dim=-1)
        if self.gc_hi is not None or self.gc_lo is not None:
            gate = gate.clamp(min=self.gc_lo, max=self.gc_hi)
        if self.uc_hi is not None or self.uc_lo is not None:
            up = up.clamp(min=self.uc_lo, max=self.uc_hi)
        if self.glu_type == "swiglu":      # gpt-oss style: gate·σ(α·gate)·(up + β)
            return gate * torch.sigmoid(self.scale * gate) * (up + self.act_bias)
        return self.act(gate) * up

    def forward(self, hidden_states: torch.Tensor, expert_indices: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x ``[E_local, C, H]`` → ``[E_local, C, H]`` (partial over TP unless ``reduce_output``)."""
        x = hidden_states      # reference parameter names in the signature
        proj = self.gate_up_proj if self.glu_mlp else self.up_proj
        return self.down_proj(self.activation(proj(x, expert_indices)), expert_indices)
