"""Quantised tensor-parallel layers (inference only; reference ``quantization/quantization_layers.py:626-966``).

Weights are stored quantised (int8 / fp8) with fp32 scales; forward either de-quantises the weight tile-wise into
the GEMM (weight-only) or, with dynamic activation quantisation, runs an fp8×fp8 GEMM and applies the outer
product of per-row activation scales and per-channel weight scales in the epilogue (``ops.gemm_fp8``)."""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
from torch import nn

from ..parallel_layers import mappings
from ..parallel_layers.layers import ColumnParallelLinear, RowParallelLinear
from ..parallel_layers.utils import set_tensor_model_parallel_attributes
from .quantization_config import ActivationQuantizationType, QuantizationType, QuantizedDtype
from .quantization_utils import (dequantize_blockwise, quantize_activation_dynamic, quantize_blockwise,
                                 quantize_per_channel, quantize_per_tensor)


class _QuantizedParallelBase(nn.Module):
    def _setup(self, base: nn.Module, q_config: Dict[str, Any]) -> None:
        self.q_config = q_config
        self.qtype: QuantizationType = q_config["quantization_type"]
        self.qdtype: torch.dtype = QuantizedDtype.get_dtype(q_config.get("quantized_dtype", QuantizedDtype.INT8)).value
        self.act_q: ActivationQuantizationType = q_config.get("activation_quantization_type", ActivationQuantizationType.NONE)
        self.clamp_bound = q_config.get("clamp_bound")
        self.dequantized_dtype = base.weight.dtype
        w = base.weight.data
        if self.qtype == QuantizationType.PER_TENSOR_SYMMETRIC:
            q, s = quantize_per_tensor(w, self.qdtype)
        elif self.qtype == QuantizationType.BLOCKWISE_SYMMETRIC:
            q, s = quantize_blockwise(w, self.qdtype, q_config["block_axis"], q_config["block_size"])
        else:
            q, s = quantize_per_channel(w, self.qdtype, q_config.get("quantization_per_channel_axis", 0))
        self.weight = nn.Parameter(q, requires_grad=False)
        self.scale = nn.Parameter(s.float(), requires_grad=False)
        set_tensor_model_parallel_attributes(self.weight, True, base.weight.partition_dim, base.weight.partition_stride,
                                             num_partitions=base.weight.num_partitions)
        per_channel_sharded = self.qtype != QuantizationType.PER_TENSOR_SYMMETRIC and s.dim() == w.dim() and \
            s.shape[base.weight.partition_dim] > 1
        set_tensor_model_parallel_attributes(self.scale, per_channel_sharded, base.weight.partition_dim if per_channel_sharded else 0,
                                             1, num_partitions=base.weight.num_partitions if per_channel_sharded else 1)
        self.bias = base.bias

    def _dequant_weight(self) -> torch.Tensor:
        if self.qtype == QuantizationType.BLOCKWISE_SYMMETRIC:
            return dequantize_blockwise(self.weight, self.scale, self.q_config["block_axis"], self.q_config["block_size"],
                                        self.dequantized_dtype)
        return (self.weight.float() * self.scale).to(self.dequantized_dtype)

    def _matmul(self, x: torch.Tensor) -> torch.Tensor:
        if self.act_q == ActivationQuantizationType.DYNAMIC and self.qdtype in (torch.float8_e4m3fn, torch.float8_e5m2) \
                and self.qtype in (QuantizationType.PER_CHANNEL_SYMMETRIC, QuantizationType.PER_TENSOR_SYMMETRIC):
            from ..ops import gemm_fp8

            xq, xs = quantize_activation_dynamic(x, self.qdtype, self.clamp_bound)
            w_scale = self.scale.reshape(-1) if self.scale.numel() > 1 else self.scale.reshape(1).expand(self.weight.shape[0])
            return gemm_fp8.scaled_linear(xq, xs, self.weight, w_scale, out_dtype=self.dequantized_dtype)
        return torch.matmul(x.to(self.dequantized_dtype), self._dequant_weight().t())


class QuantizedColumnParallel(_QuantizedParallelBase):
    def __init__(self, base: ColumnParallelLinear, q_config: Dict[str, Any]):
        super().__init__()
        self._setup(base, q_config)
        self.gather_output, self.group = base.gather_output, base.tensor_parallel_group
        self.sequence_parallel_enabled, self.sequence_dimension = base.sequence_parallel_enabled, base.sequence_dimension

    @classmethod
    def from_float(cls, mod: ColumnParallelLinear, q_config: Dict[str, Any]) -> "QuantizedColumnParallel":
        return cls(mod, q_config)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.sequence_parallel_enabled:
            x = mappings.gather_from_sequence_parallel_region(x, self.sequence_dimension, True, self.group)
        y = self._matmul(x)
        if self.gather_output:
            y = mappings.gather_from_tensor_model_parallel_region(y, self.group)
        return y if self.bias is None else y + self.bias


class QuantizedRowParallel(_QuantizedParallelBase):
    def __init__(self, base: RowParallelLinear, q_config: Dict[str, Any]):
        super().__init__()
        self._setup(base, q_config)
        self.input_is_parallel, self.group = base.input_is_parallel, base.tensor_parallel_group
        self.sequence_parallel_enabled, self.sequence_dimension = base.sequence_parallel_enabled, base.sequence_dimension
        self.reduce_output = base.reduce_output

    @classmethod
    def from_float(cls, mod: RowParallelLinear, q_config: Dict[str, Any]) -> "QuantizedRowParallel":
        return cls(mod, q_config)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.input_is_parallel:
            x = mappings.scatter_to_tensor_model_parallel_region(x, self.group)
        y = self._matmul(x)
        if self.reduce_output:
            if self.sequence_parallel_enabled:
                y = mappings.reduce_scatter_to_sequence_parallel_region(y, self.sequence_dimension, self.group)
            else:
                y = mappings.reduce_from_tensor_model_parallel_region(y, self.group)
        return y if self.bias is None else y + self.bias


class QuantizedExpertFusedColumnParallel(nn.Module):
    """Expert-wise per-channel quantised ``[E, H, I]`` weights (reference quantization_layers.py expert classes)."""

    def __init__(self, base: nn.Module, q_config: Dict[str, Any]):
        super().__init__()
        dt = QuantizedDtype.get_dtype(q_config.get("quantized_dtype", QuantizedDtype.INT8)).value
        w = base.weight.data
        scale = (w.abs().amax(dim=1, keepdim=True).float() / {torch.int8: 127.0}.get(dt, 448.0)).clamp(min=1e-12)
        from .quantization_utils import _cast

        self.weight = nn.Parameter(_cast(w.float() / scale, dt), requires_grad=False)
        self.scale = nn.Parameter(scale, requires_grad=False)
        self.dequantized_dtype = w.dtype
        self.base = [base]

    @classmethod
    def from_float(cls, mod, q_config):
        return cls(mod, q_config)

    def forward(self, x: torch.Tensor, expert_indices: Optional[torch.Tensor] = None) -> torch.Tensor:
        w = (self.weight.float() * self.scale).to(self.dequantized_dtype)
        if expert_indices is not None:
            w = w[expert_indices]
        pat = "e...h,ehi->e...i"
        return torch.einsum(pat, x, w)


QuantizedExpertFusedRowParallel = QuantizedExpertFusedColumnParallel
