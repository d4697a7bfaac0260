"""Tensor-parallel LoRA (reference ``modules/lora/tp_layer.py:15-173``).

* Column-parallel base ``W [out/tp, in]``:  A is dense ``[r, in]`` (replicated), B is ColumnParallel ``[out/tp, r]``
  → the adapter output is sharded exactly like the base output, no extra collective.
* Row-parallel base ``W [out, in/tp]``:  A is RowParallel ``[r, in/tp]`` (its tiny ``[.., r]`` partial sums are
  all-reduced), B is dense ``[out, r]``.
* GQA-QKV: A dense, B one ColumnParallel per q/k/v.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from ...parallel_layers.layers import ColumnParallelLinear, RowParallelLinear
from ..qkv_linear import GQAQKVColumnParallelLinear
from .config import LoraConfig
from .layer import LoraLayer


class LoraParallelLinear(LoraLayer):
    def __init__(self, base_layer: nn.Module, config: Optional[LoraConfig] = None, lora_config: Optional[LoraConfig] = None, **_compat):
        config = config if config is not None else lora_config          # ``lora_config=`` is the reference's keyword
        assert config is not None, "a LoraConfig is required"
        super().__init__(base_layer, config)
        dt, dev = base_layer.weight.dtype, base_layer.weight.device
        self.is_row = isinstance(base_layer, RowParallelLinear)
        if self.is_row:
            self.lora_A = RowParallelLinear(base_layer.input_size, self.r, bias=False, input_is_parallel=base_layer.input_is_parallel,
                                            dtype=dt, device=dev, sequence_parallel_enabled=False,
                                            tensor_model_parallel_group=base_layer.tensor_parallel_group)
            self.lora_B = nn.Linear(self.r, base_layer.output_size, bias=False, dtype=dt, device=dev)
            self.init_lora_parameters(self.lora_A.weight, self.lora_B.weight)
            setattr(self.lora_B.weight, "sequence_parallel_enabled", base_layer.sequence_parallel_enabled)
        else:
            self.lora_A = nn.Linear(base_layer.input_size, self.r, bias=False, dtype=dt, device=dev)
            self.lora_B = ColumnParallelLinear(self.r, base_layer.output_size, bias=False, gather_output=base_layer.gather_output,
                                               dtype=dt, device=dev, stride=base_layer.stride,
                                               tensor_model_parallel_group=base_layer.tensor_parallel_group)
            self.init_lora_parameters(self.lora_A.weight, self.lora_B.weight)
            setattr(self.lora_A.weight, "sequence_parallel_enabled", base_layer.sequence_parallel_enabled)

    def delta_weight(self) -> torch.Tensor:
        return (self.lora_B.weight @ self.lora_A.weight) * self.scaling   # local shard of ΔW in both cases

    def forward(self, x: torch.Tensor, *a, **k):
        y = self.base_layer(x, *a, **k)
        bias = None
        if isinstance(y, tuple):
            y, bias = y
        if not self.merged:
            from ...parallel_layers import mappings

            base = self.base_layer
            xin = self.dropout(x)
            if self.is_row:
                lo = self.lora_B(self.lora_A(xin))                 # all-reduced [.., r] → dense B
                if base.sequence_parallel_enabled:
                    lo = mappings.scatter_to_sequence_parallel_region(lo, base.sequence_dimension, base.tensor_parallel_group)
            else:
                if base.sequence_parallel_enabled:
                    xin = mappings.gather_from_sequence_parallel_region(xin, base.sequence_dimension, True, base.tensor_parallel_group)
                lo = self.lora_B(self.lora_A(xin))
            y = y + lo * self.scaling
        return (y, bias) if bias is not None else y


class LoraGQAQKVParallelLinear(LoraLayer):
    def __init__(self, base_layer: GQAQKVColumnParallelLinear, config: Optional[LoraConfig] = None, lora_config: Optional[LoraConfig] = None, **_compat):
        config = config if config is not None else lora_config          # ``lora_config=`` is the reference's keyword
        assert config is not None, "a LoraConfig is required"
        nn.Module.__init__(self)
        self.base_layer, self.lora_config = base_layer, config
        self.r, self.scaling, self.merged = config.lora_rank, config.scaling, False
        self.dropout = nn.Dropout(config.lora_dropout) if config.lora_dropout > 0 else nn.Identity()
        for p in base_layer.parameters():
            p.requires_grad_(False)
        w = base_layer.weight_qkv if base_layer.fuse_qkv else base_layer.weight_q
        dt, dev = w.dtype, w.device
        self.lora_A = nn.Linear(base_layer.input_size, self.r, bias=False, dtype=dt, device=dev)
        q, kv = base_layer.output_sizes[0], base_layer.output_sizes[1] * base_layer.kv_size_multiplier
        mk = lambda out: ColumnParallelLinear(self.r, out, bias=False, gather_output=base_layer.gather_output, dtype=dt,
                                              device=dev, tensor_model_parallel_group=base_layer.tensor_parallel_group)
        self.lora_B_q, self.lora_B_k, self.lora_B_v = mk(q), mk(kv), mk(kv)
        for b in (self.lora_B_q, self.lora_B_k, self.lora_B_v):
            self.init_lora_parameters(self.lora_A.weight, b.weight)

    def forward(self, x: torch.Tensor):
        q, k, v = self.base_layer(x)
        if self.merged:
            return q, k, v
        from ...parallel_layers import mappings

        xin = self.dropout(x)
        if self.base_layer.sequence_parallel_enabled:
            xin = mappings.gather_from_sequence_parallel_region(xin, self.base_layer.sequence_dimension, True,
                                                                self.base_layer.tensor_parallel_group)
        a = self.lora_A(xin)
        s = self.scaling
        return q + self.lora_B_q(a) * s, k + self.lora_B_k(a) * s, v + self.lora_B_v(a) * s

    def get_qkv(self, layer):
        """The (local) Q / K / V weights of a GQA-QKV layer as three views (reference ``lora/tp_layer.py:137-146``); for the
        fused layout these are slices of ``weight_qkv`` so in-place updates land in the fused parameter."""
        if getattr(layer, "fuse_qkv", False):
            sizes = [layer.q_output_size_per_partition, layer.kv_output_size_per_partition, layer.kv_output_size_per_partition]
            return torch.split(layer.weight_qkv, sizes, dim=0)
        return layer.weight_q, layer.weight_k, layer.weight_v

    def get_delta_weight(self):
        """``(ΔW_q, ΔW_k, ΔW_v)`` of this rank: ``B_x · A · scaling`` (``B_x`` is the local column shard)."""
        a = self.lora_A.weight
        return tuple((b.weight @ a) * self.scaling for b in (self.lora_B_q, self.lora_B_k, self.lora_B_v))

    def merge(self, safe_merge: bool = False) -> None:
        """Fold the adapter into the base Q / K / V shards (reference :100-120).  ``safe_merge`` checks for non-finite values
        before touching the weights."""
        if self.merged:
            return
        deltas = self.get_delta_weight()
        targets = self.get_qkv(self.base_layer)
        if safe_merge:
            for t, d in zip(targets, deltas):
                if not torch.isfinite(t.data + d.to(t.dtype)).all():
                    raise ValueError("NaNs detected in the merged weights. The adapter seems to be broken")
        with torch.no_grad():
            for t, d in zip(targets, deltas):
                t.data.add_(d.to(t.dtype))
        self.merged = True

    def unmerge(self) -> None:
        if not self.merged:
            return
        with torch.no_grad():
            for t, d in zip(self.get_qkv(self.base_layer), self.get_delta_weight()):
                t.data.sub_(d.to(t.dtype))
        self.merged = False
