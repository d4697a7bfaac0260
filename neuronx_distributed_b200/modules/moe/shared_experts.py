"""Always-on shared experts (reference ``modules/moe/shared_experts.py:73``): a dense GLU MLP on TP layers whose
output is added to the routed experts' output *before* the single delayed reduction."""
from __future__ import annotations

import torch
from torch import nn

from ...parallel_layers.layers import ColumnParallelLinear, RowParallelLinear
from .experts import ACT2FN


class SharedExperts(nn.Module):
    def __init__(self, hidden_size: int, intermediate_size: int, num_shared_experts: int = 1, hidden_act: str = "silu",
                 dtype=torch.float32, device=None, reduce_dtype=torch.float32, fused_gate_up_projection: bool = True,
                 sequence_parallel_enabled: bool = False, transpose_weights: bool = False, tensor_model_parallel_group=None):
        super().__init__()
        inter = intermediate_size * num_shared_experts
        self.act = ACT2FN[hidden_act]
        self.gate_up_proj = ColumnParallelLinear(hidden_size, 2 * inter, bias=False, gather_output=False, stride=2,
                                                 dtype=dtype, device=device,
                                                 tensor_model_parallel_group=tensor_model_parallel_group)
        # reduce_output=False: the MoE layer reduces routed + shared together
        self.down_proj = RowParallelLinear(inter, hidden_size, bias=False, input_is_parallel=True, reduce_output=False,
                                           dtype=dtype, device=device, reduce_dtype=reduce_dtype,
                                           tensor_model_parallel_group=tensor_model_parallel_group)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        g, u = self.gate_up_proj(x).chunk(2, dim=-1)
        return self.down_proj(self.act(g) * u)
