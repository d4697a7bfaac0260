"""Expert-fused TP linears (reference ``modules/moe/moe_parallel_layers.py:18-431``): one 3-D weight holds all
local experts — ``[E_local, H, 2I/tp]`` for the fused gate|up column projection (stride 2) and
``[E_local, I/tp, H]`` for the row projection; forward is a batched ``e…h,ehi→e…i`` contraction; the row layer's
partial sums are reduced by the caller (MoE delays the reduction until after the shared experts)."""
from __future__ import annotations

import math
from typing import Callable, Optional

import torch
from torch import nn

from ...parallel_layers import mappings
from ...parallel_layers import parallel_state as ps
from ...parallel_layers.layers import BaseParallelLayer, _group_info
from ...parallel_layers.utils import create_local_weight, divide, set_tensor_model_parallel_attributes


def _local_experts(num_experts: int):
    ep, r = ps.get_expert_model_parallel_size(), ps.get_expert_model_parallel_rank()
    return ps.get_experts_for_expert_parallel_rank(r, num_experts, ep), ep


class _ExpertFusedBase(BaseParallelLayer):
    def _make(self, num_experts, full_shape, partition_dim, stride, dtype, device, init_method, group):
        self.tensor_parallel_group, self.tp, self.tp_rank = _group_info(group)
        self.num_experts = num_experts
        self.local_expert_ids, self.ep = _local_experts(num_experts)
        self.num_local_experts = len(self.local_expert_ids)
        local = list(full_shape)
        local[partition_dim] = divide(full_shape[partition_dim], self.tp)
        w = nn.Parameter(torch.empty(self.num_local_experts, *local[1:], dtype=dtype, device=device or torch.device("cpu")))
        set_tensor_model_parallel_attributes(w, True, partition_dim, stride, num_partitions=self.tp)
        if self.ep > 1:
            w.expert_model_parallel = True
        if w.device.type != "meta":
            with torch.no_grad():
                for li, e in enumerate(self.local_expert_ids):
                    g = torch.Generator().manual_seed(1000003 * (e + 1) + full_shape[1] * 31 + full_shape[2])
                    full = torch.empty(full_shape[1:], dtype=torch.float32)
                    if init_method is None:
                        bound = 1.0 / math.sqrt(full_shape[1])
                        full.uniform_(-bound, bound, generator=g)
                    else:
                        init_method(full)
                    shard = create_local_weight(full.to(dtype), partition_dim - 1, local[partition_dim], stride,
                                                rank=self.tp_rank, world_size=self.tp)
                    w[li].copy_(shard)
        return w


class ExpertFusedColumnParallelLinear(_ExpertFusedBase):
    def __init__(self, num_experts: int, input_size: int, output_size: int, dtype=torch.float32, device=None,
                 stride: int = 1, init_method: Optional[Callable] = None, tensor_model_parallel_group=None, bias: bool = False):
        super().__init__()
        self.input_size, self.output_size = input_size, output_size
        self.weight = self._make(num_experts, (num_experts, input_size, output_size), 2, stride, dtype, device, init_method,
                                 tensor_model_parallel_group)
        self.register_parameter("bias", None)

    def forward(self, x: torch.Tensor, expert_indices: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x ``[E_local, C, H]`` → ``[E_local, C, out/tp]`` (dgrad all-reduce over TP in backward)."""
        x = mappings.copy_to_tensor_model_parallel_region(x, self.tensor_parallel_group)
        w = self.weight if expert_indices is None else self.weight[expert_indices]
        return torch.einsum("e...h,ehi->e...i", x, w)


class ExpertFusedRowParallelLinear(_ExpertFusedBase):
    def __init__(self, num_experts: int, input_size: int, output_size: int, reduce_output: bool = True,
                 dtype=torch.float32, device=None, stride: int = 1, init_method: Optional[Callable] = None,
                 tensor_model_parallel_group=None, bias: bool = False):
        super().__init__()
        self.input_size, self.output_size, self.reduce_output = input_size, output_size, reduce_output
        self.weight = self._make(num_experts, (num_experts, input_size, output_size), 1, stride, dtype, device, init_method,
                                 tensor_model_parallel_group)
        self.register_parameter("bias", None)

    def forward(self, x: torch.Tensor, expert_indices: Optional[torch.Tensor] = None) -> torch.Tensor:
        w = self.weight if expert_indices is None else self.weight[expert_indices]
        out = torch.einsum("e...i,eih->e...h", x, w)
        if self.reduce_output:
            out = mappings.reduce_from_tensor_model_parallel_region(out, self.tensor_parallel_group)
        return out
