"""Always-on shared experts (reference ``modules/moe/shared_experts.py:14-330``): a dense GLU MLP on TP layers whose
output is added to the routed experts' output *before* the single delayed reduction (``reduce_output=False`` on the
down projection).

Options, as in the reference:

* ``fused_gate_up_projection`` — one ``gate_up_proj`` (stride-2 column layer) instead of ``gate_proj`` + ``up_proj``
  (checkpoint key names follow);
* ``transpose_weights`` — weights stored ``[in, out]`` (``*Transposed`` layers): the decode GEMV streams the weight in
  its K-major order without a transposed copy;
* ``sequence_parallel_enabled`` — weights are *replicated* and every rank runs the MLP on its own sequence shard during
  prefill (no collective at all); at decode (``seq_len == 1``) each rank slices its TP share out of the replicated
  weight (``SPMDRank`` + ``indices_split_along_dim``) and the caller's reduction sums the partial outputs.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
import torch.distributed as dist
from torch import nn

from ...parallel_layers import parallel_state as ps
from ...parallel_layers.layers import (ColumnParallelLinear, ProcessGroupSafeDeepcopy, RowParallelLinear, SPMDRank,
                                        tp_linear)
from ...parallel_layers.utils import indices_split_along_dim
from .model_utils import ACT2FN, create_spmd_ranks

weight_cache: Dict[str, Any] = {}


class ColumnParallelLinearTransposed(ColumnParallelLinear):
    """Column-parallel linear whose parameter is stored ``[in, out/tp]`` (reference :14-42).  Built by allocating the
    regular ``[out/tp, in]`` shard (same initialisation stream) and re-registering its transpose, partition dim 1."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        _store_transposed(self, new_partition_dim=1)

    def forward(self, input: torch.Tensor, slice_indices: Optional[torch.Tensor] = None, *_: Any):  # noqa: A002
        self._check_pad_false_for_training()
        w = self.weight if slice_indices is None else self.weight.index_select(1, slice_indices)
        tp = self.tensor_model_parallel_size
        in_mode = "gather" if self.sequence_parallel_enabled else ("copy" if tp > 1 else "none")
        out = tp_linear(input, w.t(), None, in_mode, "none", self.sequence_dimension, self.tensor_parallel_group, self.reduce_dtype)
        if self.gather_output:
            from ...parallel_layers import mappings

            out = mappings.gather_from_tensor_model_parallel_region(out, self.tensor_parallel_group)
        if self.skip_bias_add:
            return out, self.bias
        return out if self.bias is None else out + self.bias


class RowParallelLinearTransposed(RowParallelLinear):
    """Row-parallel linear whose parameter is stored ``[in/tp, out]`` (reference :45-72), partition dim 0."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        _store_transposed(self, new_partition_dim=0)

    def forward(self, input_: torch.Tensor, slice_indices: Optional[torch.Tensor] = None, *_: Any):
        self._check_pad_false_for_training()
        from ...parallel_layers import mappings

        x = input_ if self.input_is_parallel else mappings.scatter_to_tensor_model_parallel_region(input_, self.tensor_parallel_group)
        w = self.weight if slice_indices is None else self.weight.index_select(0, slice_indices)
        if not self.reduce_output or self.tensor_model_parallel_size == 1:
            out_mode = "none"
        else:
            out_mode = "scatter" if self.sequence_parallel_enabled else "reduce"
        out = tp_linear(x, w.t(), None, "none", out_mode, self.sequence_dimension, self.tensor_parallel_group, self.reduce_dtype)
        if self.skip_bias_add:
            return out, self.bias
        return out if self.bias is None else out + self.bias


def _store_transposed(layer: nn.Module, new_partition_dim: int) -> None:
    old = layer.weight
    new = nn.Parameter(old.data.t().contiguous(), requires_grad=old.requires_grad)
    for k, v in old.__dict__.items():
        setattr(new, k, v)
    new.partition_dim = new_partition_dim
    layer.weight = new
    layer.weight_partition_dim = new_partition_dim
    if getattr(layer, "master_weight", None) is not None:
        layer.master_weight = layer.master_weight.t().contiguous()


def _self_group():
    """A process group containing only this rank (replicated-weight layers).  Collective: every rank creates all of
    them, cached per world."""
    key = f"self_groups_{dist.get_world_size()}"
    if key not in weight_cache:
        mine = None
        for r in range(dist.get_world_size()):
            g = dist.new_group([r])
            if r == dist.get_rank():
                mine = g
        weight_cache[key] = mine
    return weight_cache[key]


class SharedExperts(ProcessGroupSafeDeepcopy, nn.Module):
    def __init__(self, hidden_size: int, intermediate_size: int, num_shared_experts: int = 1, hidden_act: str = "silu",
                 dtype: torch.dtype = torch.float32, tensor_model_parallel_group=None, reduce_dtype: torch.dtype = torch.float32,
                 fused_gate_up_projection: bool = False, sequence_parallel_enabled: bool = False,
                 transpose_weights: bool = False, device=None):
        super().__init__()
        self.hidden_size, self.intermediate_size, self.num_shared_experts = hidden_size, intermediate_size, num_shared_experts
        self.act_fn = self.act = ACT2FN[hidden_act]
        self.dtype, self.reduce_dtype, self.device = dtype, reduce_dtype, device
        self.fused_gate_up_projection, self.transpose_weights = fused_gate_up_projection, transpose_weights
        self.sequence_parallel_enabled = sequence_parallel_enabled
        if transpose_weights and fused_gate_up_projection:
            raise ValueError("transpose_weights requires separate gate / up projections")
        self.world_size = dist.get_world_size(ps.get_world_group()) if ps.model_parallel_is_initialized() else 1
        if sequence_parallel_enabled:
            self.slice_size = dist.get_world_size(tensor_model_parallel_group or ps.get_tensor_model_parallel_group())
            self.spmd_rank = SPMDRank(world_size=self.world_size, tensor_model_parallel_size=self.slice_size)
            self.tensor_parallel_group = _self_group() if self.world_size > 1 else ps.get_tensor_model_parallel_group()
        else:
            self.tensor_parallel_group = tensor_model_parallel_group or ps.get_tensor_model_parallel_group()
        self._initialize_parallel_layers()

    # ---------------------------------------------------------------------------------------------- construction
    def _initialize_parallel_layers(self) -> None:
        inter = self.intermediate_size * self.num_shared_experts
        if self.fused_gate_up_projection:
            self.gate_up_proj = self._create_column_parallel_linear(inter * 2, stride=2)
            self.down_proj = self._create_row_parallel_linear(inter)
        elif self.transpose_weights:
            # stored transposed: gate/up hold [H, inter/tp], down holds [inter/tp, H] — the "row" / "column" roles of the
            # *stored* matrices swap, which is how the reference builds them
            self.gate_proj = self._create_column_parallel_linear(inter)
            self.up_proj = self._create_column_parallel_linear(inter)
            self.down_proj = self._create_row_parallel_linear(inter)
        else:
            self.gate_proj = self._create_column_parallel_linear(inter)
            self.up_proj = self._create_column_parallel_linear(inter)
            self.down_proj = self._create_row_parallel_linear(inter)

    def _create_column_parallel_linear(self, output_size: int, stride: int = 1):
        cls = ColumnParallelLinearTransposed if self.transpose_weights else ColumnParallelLinear
        return cls(self.hidden_size, output_size, stride=stride, bias=False, gather_output=False, dtype=self.dtype,
                   device=self.device, reduce_dtype=self.reduce_dtype, tensor_model_parallel_group=self.tensor_parallel_group)

    def _create_row_parallel_linear(self, input_size: int):
        cls = RowParallelLinearTransposed if self.transpose_weights else RowParallelLinear
        # reduce_output=False: the MoE layer reduces routed + shared together
        return cls(input_size, self.hidden_size, bias=False, input_is_parallel=True, reduce_output=False, dtype=self.dtype,
                   device=self.device, reduce_dtype=self.reduce_dtype, tensor_model_parallel_group=self.tensor_parallel_group)

    def preshard_hook(self, model_state_dict: Dict[str, Any], prefix: str) -> None:
        """Checkpoints store ``[out, in]``; transpose for ``transpose_weights`` and add the SPMD-rank entry for SP."""
        base = prefix[: prefix.rfind(".") + 1] if "." in prefix else ""
        base = base[: -len("gate_proj.")] if base.endswith("gate_proj.") else base
        base = base[: -len("gate_up_proj.")] if base.endswith("gate_up_proj.") else base
        base = base[: -len("spmd_rank.")] if base.endswith("spmd_rank.") else base
        if self.transpose_weights:
            for name, layer in (("gate_proj", self.gate_proj), ("up_proj", self.up_proj), ("down_proj", self.down_proj)):
                key = f"{base}{name}.weight"
                full_t = (layer.weight.shape[0] * (layer.tensor_model_parallel_size if layer.weight.partition_dim == 0 else 1),
                          layer.weight.shape[1] * (layer.tensor_model_parallel_size if layer.weight.partition_dim == 1 else 1))
                if key in model_state_dict and tuple(model_state_dict[key].shape) != full_t:
                    model_state_dict[key] = model_state_dict[key].t().contiguous()
        if self.sequence_parallel_enabled:
            create_spmd_ranks(model_state_dict, base, self.world_size)

    # ---------------------------------------------------------------------------------------------- forward
    def forward(self, x: torch.Tensor, seq_len: Optional[int] = None) -> torch.Tensor:
        if seq_len == 1 and self.sequence_parallel_enabled:
            return self._forward_token_gen_replicated_weights(x)
        return self._forward(x)

    def _fused_activation(self, x: torch.Tensor) -> torch.Tensor:
        gate, up = torch.chunk(x, 2, dim=-1)
        return self.act_fn(gate) * up

    def _forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.fused_gate_up_projection:
            h = self._fused_activation(self.gate_up_proj(x))
        else:
            h = self.act_fn(self.gate_proj(x)) * self.up_proj(x)
        return self.down_proj(h)

    def get_split_indices(self, weight: torch.Tensor, dim: int) -> torch.Tensor:
        return indices_split_along_dim(weight, dim, rank=self.spmd_rank.rank.data.to(torch.long), num_partitions=self.slice_size)

    def _fused_split_indices(self, weight: torch.Tensor, dim: int) -> torch.Tensor:
        """gate|up halves are concatenated along ``dim``: take this rank's slice of each half."""
        half = weight.size(dim) // 2
        per = half // self.slice_size
        r = self.spmd_rank.rank.data.to(torch.long)
        base = torch.arange(per, device=weight.device) + r * per
        return torch.cat([base, base + half])

    def _forward_token_gen_replicated_weights(self, x: torch.Tensor) -> torch.Tensor:
        out_dim = 1 if self.transpose_weights else 0
        if self.fused_gate_up_projection:
            h = self._fused_activation(self.gate_up_proj(x, self._fused_split_indices(self.gate_up_proj.weight, out_dim)))
        else:
            gate = self.gate_proj(x, self.get_split_indices(self.gate_proj.weight, out_dim))
            up = self.up_proj(x, self.get_split_indices(self.up_proj.weight, out_dim))
            h = self.act_fn(gate) * up
        return self.down_proj(h, self.get_split_indices(self.down_proj.weight, 1 - out_dim))
