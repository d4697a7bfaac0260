"""Expert-fused TP linears (reference ``modules/moe/moe_parallel_layers.py:18-431``): one 3-D weight holds all
local experts — ``[E_local, H, 2I/tp]`` for the fused gate|up column projection (stride 2) and
``[E_local, I/tp, H]`` for the row projection; forward is a batched ``e…h,ehi→e…i`` contraction; the row layer's
partial sums are reduced by the caller unless ``reduce_output`` (MoE delays the reduction until after the shared
experts are added).  Optional per-expert bias ``[E_local, out]`` is broadcast over the token dims."""
from __future__ import annotations

import math
from typing import Any, Callable, List, Optional

import torch
import torch.distributed as dist
from torch import nn

from ...parallel_layers import mappings
from ...parallel_layers import parallel_state as ps
from ...parallel_layers.layers import BaseParallelLayer, _group_info
from ...parallel_layers.utils import create_local_weight, divide, set_tensor_model_parallel_attributes


class ExpertFusedLinearWithAsyncCommunication(torch.autograd.Function):
    """``out[e] = x[e] @ W[e]`` for all experts in one contraction (reference :18-139).

    input ``(E, …, H)``, weight ``(E, H, I)`` → ``(E, …, I)``; ``E == 1`` on the input broadcasts one token block over all
    experts.  Backward: dgrad contraction, optional all-reduce of the input gradient over the TP group (column layers —
    the forward "copy to TP region"), then the wgrad contraction, so the collective overlaps the wgrad GEMM on the NCCL
    stream."""

    @staticmethod
    def forward(ctx, input, weight, bias, async_grad_allreduce, sequence_parallel_enabled, sequence_dimension=0,  # noqa: A002
                save_for_backward=True, process_group=None, reduce_dtype=torch.float32):
        if bias is not None:
            raise NotImplementedError("bias is added by the layer, not by the fused contraction")
        if sequence_parallel_enabled:
            raise NotImplementedError("sequence parallelism is exited before the expert MLPs; not supported here")
        if input.shape[0] != weight.shape[0] and input.shape[0] > 1:
            raise RuntimeError(f"input and weight disagree on the number of experts: input_shape={tuple(input.shape)}, "
                               f"weight_shape={tuple(weight.shape)}")
        ctx.async_grad_allreduce = async_grad_allreduce
        ctx.compute_weight_gradient = weight.requires_grad
        ctx.process_group = process_group if process_group is not None else ps.get_tensor_model_parallel_group()
        if save_for_backward:
            ctx.save_for_backward(*((input, weight) if ctx.compute_weight_gradient else (weight,)))
        return torch.einsum("e...h,ehi->e...i", input, weight)

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.compute_weight_gradient:
            x, weight = ctx.saved_tensors
        else:
            (weight,), x = ctx.saved_tensors, None
        grad_input = torch.einsum("e...i,ehi->e...h", grad_output, weight)
        if x is not None and x.shape[0] == 1 and weight.shape[0] > 1:
            grad_input = grad_input.sum(0, keepdim=True)
        work = None
        if ctx.async_grad_allreduce and dist.get_world_size(ctx.process_group) > 1:
            grad_input = grad_input.contiguous()
            work = dist.all_reduce(grad_input, group=ctx.process_group, async_op=True)
        grad_weight = None
        if ctx.compute_weight_gradient:
            xe = x.expand(weight.shape[0], *x.shape[1:]) if x.shape[0] == 1 else x
            grad_weight = torch.einsum("e...h,e...i->ehi", xe, grad_output)
        if work is not None:
            work.wait()
        return grad_input, grad_weight, None, None, None, None, None, None, None


class ExpertFusedLinear(nn.Module):
    """Mixin: tags parameters as expert-parallel (``param.expert_model_parallel``) so that gradient reduction, ZeRO-1
    sharding and checkpointing treat them over the expert-data-parallel group (reference :141-175)."""

    def _mark_expert_parallel_weights(self, iterable=None, expert_parallel_group_size: Optional[int] = None,
                                      is_prefill: bool = True, expert_distribution=None) -> None:
        if expert_parallel_group_size is None:
            expert_parallel_group_size = ps.get_expert_model_parallel_size()
        if expert_parallel_group_size <= 1:
            return
        for p in (self.parameters() if iterable is None else iterable):
            p.expert_model_parallel = True
            if is_prefill:
                p.is_prefill = True
            p.expert_distribution = expert_distribution

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)           # .to()/.cuda() may re-create parameters: re-tag them
        self._mark_expert_parallel_weights(expert_parallel_group_size=getattr(self, "ep", None),
                                           is_prefill=getattr(self, "is_prefill", True),
                                           expert_distribution=getattr(self, "expert_distribution", None))
        return out


def _local_experts(num_experts: int, expert_model_parallel_group=None, expert_distribution=None):
    if expert_model_parallel_group is not None:
        ep, r = dist.get_world_size(expert_model_parallel_group), dist.get_rank(expert_model_parallel_group)
    else:
        ep, r = ps.get_expert_model_parallel_size(), ps.get_expert_model_parallel_rank()
    if expert_distribution is not None:
        return list(expert_distribution[r]), ep
    return ps.get_experts_for_expert_parallel_rank(r, num_experts, ep), ep


class _ExpertFusedBase(BaseParallelLayer, ExpertFusedLinear):
    autograd_func_class = ExpertFusedLinearWithAsyncCommunication

    def _configure(self, num_experts, input_size, output_size, dtype, device, stride, init_method, keep_master_weight,
                   tensor_model_parallel_group, expert_model_parallel_group, is_prefill, is_fused_gate_up,
                   expert_distribution) -> None:
        self.num_experts, self.input_size, self.output_size = num_experts, input_size, output_size
        self.dtype, self.stride, self.arg_init_method = dtype, stride, init_method
        self.device = device if device is not None else torch.device("cpu")
        self.keep_master_weight = keep_master_weight
        self.expert_model_parallel_group = expert_model_parallel_group
        self.is_prefill, self.is_fused_gate_up, self.expert_distribution = is_prefill, is_fused_gate_up, expert_distribution
        self.tensor_parallel_group, self.tp, self.tp_rank = _group_info(tensor_model_parallel_group)
        self.tensor_model_parallel_size = self.tp
        self.local_expert_ids, self.ep = _local_experts(num_experts, expert_model_parallel_group, expert_distribution)
        self.num_local_experts = self._n_local_experts = len(self.local_expert_ids)
        self.sequence_parallel_enabled, self.sequence_dimension = False, 0
        self.master_weight: Optional[torch.Tensor] = None

    def _make(self, full_shape, partition_dim):
        """Allocate the local ``[E_local, …]`` shard and fill it from a per-expert seeded full weight (every TP rank
        draws the same full matrix for an expert and keeps its slice — TP-degree-independent initialisation)."""
        local = list(full_shape)
        local[partition_dim] = divide(full_shape[partition_dim], self.tp)
        w = nn.Parameter(torch.empty(self.num_local_experts, *local[1:], dtype=self.dtype, device=self.device))
        set_tensor_model_parallel_attributes(w, True, partition_dim, self.stride, num_partitions=self.tp)
        masters: List[torch.Tensor] = []
        if w.device.type != "meta":
            with torch.no_grad():
                for li, e in enumerate(self.local_expert_ids):
                    g = torch.Generator().manual_seed(1000003 * (e + 1) + full_shape[1] * 31 + full_shape[2])
                    full = torch.empty(full_shape[1:], dtype=torch.float32)
                    if self.arg_init_method is None:
                        bound = 1.0 / math.sqrt(full_shape[1])
                        full.uniform_(-bound, bound, generator=g)
                    else:
                        self.arg_init_method(full)
                    if self.keep_master_weight:
                        masters.append(full.to(self.dtype))
                    shard = create_local_weight(full.to(self.dtype), partition_dim - 1, local[partition_dim], self.stride,
                                                rank=self.tp_rank, world_size=self.tp)
                    w[li].copy_(shard)
        if masters:
            self.master_weight = torch.stack(masters)
        return w

    def set_weight_and_bias_config(self) -> None:
        """Override point: full logical weight shape ``(E, in, out)`` and the TP partition dim (reference :234-246)."""
        self.weight_shape = (self.num_experts, self.input_size, self.output_size)

    def init_weight_cpu(self) -> None:
        """Re-draw this rank's expert shards from the per-expert seeded full weights."""
        with torch.no_grad():
            self.weight.data.copy_(self._make(self.weight_shape, self.weight_partition_dim).data)

    def initialize_weight_and_bias(self) -> None:
        self.set_weight_and_bias_config()
        self.init_weight_cpu()
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()

    def preshard_hook(self, model_state_dict, prefix: str) -> None:
        """Full checkpoints may store per-expert tensors (``….experts.<e>.<proj>.weight``, HF Mixtral style ``[out, in]``):
        stack them into this layer's single ``[E, in, out]`` entry when that entry is missing."""
        if prefix in model_state_dict:
            return
        base = prefix[: prefix.rfind(".") + 1]
        pat = base + "{e}.weight"
        if all(pat.format(e=e) in model_state_dict for e in range(self.num_experts)):
            model_state_dict[prefix] = torch.stack([model_state_dict.pop(pat.format(e=e)).t() for e in range(self.num_experts)])

    def _make_bias(self, size: int, partitioned: bool):
        b = nn.Parameter(torch.zeros(self.num_local_experts, size, dtype=self.dtype, device=self.device))
        if partitioned:
            set_tensor_model_parallel_attributes(b, True, 1, self.stride, num_partitions=self.tp)
        return b

    def _bias_for(self, expert_indices: Optional[torch.Tensor], ndim: int) -> Optional[torch.Tensor]:
        if self.bias is None:
            return None
        b = self.bias if expert_indices is None else self.bias[expert_indices]
        return b.reshape(b.shape[0], *([1] * (ndim - 2)), b.shape[1])       # (e, 1…, out) against (e, …, out)


class ExpertFusedColumnParallelLinear(_ExpertFusedBase):
    def __init__(self, num_experts: int, input_size: int, output_size: int, bias: bool = False, dtype=torch.float32,
                 device=None, stride: int = 1, init_method: Optional[Callable] = None, keep_master_weight: bool = False,
                 tensor_model_parallel_group=None, expert_model_parallel_group=None, is_prefill: bool = True,
                 is_fused_gate_up: bool = False, expert_distribution: Optional[List[List[int]]] = None):
        super().__init__()
        self._configure(num_experts, input_size, output_size, dtype, device, stride, init_method, keep_master_weight,
                        tensor_model_parallel_group, expert_model_parallel_group, is_prefill, is_fused_gate_up or stride == 2,
                        expert_distribution)
        self.gather_output = False
        self.output_size_per_partition = divide(output_size, self.tp)
        self.weight_partition_dim = 2
        self.async_tensor_model_parallel_allreduce = self.tp > 1
        self.weight = self._make((num_experts, input_size, output_size), 2)
        if bias:
            self.bias = self._make_bias(self.output_size_per_partition, True)
        else:
            self.register_parameter("bias", None)
        self._mark_expert_parallel_weights(expert_parallel_group_size=self.ep, is_prefill=is_prefill,
                                           expert_distribution=expert_distribution)

    def forward(self, input_: torch.Tensor, expert_indices: Optional[torch.Tensor] = None, *_: Any) -> torch.Tensor:
        """x ``[E_local, …, H]`` → ``[E_local, …, out/tp]`` (dgrad all-reduce over TP in backward)."""
        x = input_      # reference parameter names in the signature
        w = self.weight if expert_indices is None else self.weight[expert_indices]
        out = self.autograd_func_class.apply(x, w, None, self.async_tensor_model_parallel_allreduce, False, 0, True,
                                             self.tensor_parallel_group)
        b = self._bias_for(expert_indices, out.dim())
        return out if b is None else out + b


class ExpertFusedRowParallelLinear(_ExpertFusedBase):
    def __init__(self, num_experts: int, input_size: int, output_size: int, bias: bool = False, reduce_output: bool = True,
                 dtype=torch.float32, device=None, stride: int = 1, init_method: Optional[Callable] = None,
                 keep_master_weight: bool = False, tensor_model_parallel_group=None, expert_model_parallel_group=None,
                 is_prefill: bool = True, is_fused_gate_up: bool = False,
                 expert_distribution: Optional[List[List[int]]] = None):
        super().__init__()
        self._configure(num_experts, input_size, output_size, dtype, device, stride, init_method, keep_master_weight,
                        tensor_model_parallel_group, expert_model_parallel_group, is_prefill, is_fused_gate_up,
                        expert_distribution)
        self.reduce_output, self.input_is_parallel = reduce_output, True
        self.input_size_per_partition = divide(input_size, self.tp)
        self.weight_partition_dim = 1
        self.weight = self._make((num_experts, input_size, output_size), 1)
        if bias:
            # with ``reduce_output=False`` every TP rank adds the bias to its partial sum: the caller's reduction would
            # count it tp times, so the stored bias is the full one and it is scaled by 1/tp when applied
            self.bias = self._make_bias(output_size, False)
        else:
            self.register_parameter("bias", None)
        self._mark_expert_parallel_weights(expert_parallel_group_size=self.ep, is_prefill=is_prefill,
                                           expert_distribution=expert_distribution)

    def forward(self, input_: torch.Tensor, expert_indices: Optional[torch.Tensor] = None, *_: Any) -> torch.Tensor:
        x = input_      # reference parameter names in the signature
        w = self.weight if expert_indices is None else self.weight[expert_indices]
        out = self.autograd_func_class.apply(x, w, None, False, False, 0, True, self.tensor_parallel_group)
        if self.reduce_output:
            out = mappings.reduce_from_tensor_model_parallel_region(out, self.tensor_parallel_group)
        b = self._bias_for(expert_indices, out.dim())
        if b is None:
            return out
        return out + (b if self.reduce_output else b / self.tp)
