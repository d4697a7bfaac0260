"""``convert`` — swap TP layers for their quantised counterparts (reference ``quantization/quantize.py:18-146``,
``quantization_mappings.py:11-16``)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

from torch import nn

from ..modules.moe.moe_parallel_layers import ExpertFusedColumnParallelLinear, ExpertFusedRowParallelLinear
from ..parallel_layers.layers import ColumnParallelLinear, RowParallelLinear
from .quantization_config import get_default_per_tensor_custom_qconfig_dict
from .quantization_layers import (QuantizedColumnParallel, QuantizedExpertFusedColumnParallel,
                                  QuantizedExpertFusedRowParallel, QuantizedRowParallel)


def get_default_quant_module_mappings() -> Dict[type, type]:
    return {
        ColumnParallelLinear: QuantizedColumnParallel,
        RowParallelLinear: QuantizedRowParallel,
        ExpertFusedColumnParallelLinear: QuantizedExpertFusedColumnParallel,
        ExpertFusedRowParallelLinear: QuantizedExpertFusedRowParallel,
    }


def convert(module: nn.Module, q_config: Optional[Dict[str, Any]] = None, inplace: bool = False,
            mapping: Optional[Dict[type, type]] = None, include: Optional[List[str]] = None,
            modules_to_not_convert: Optional[List[str]] = None) -> nn.Module:
    q_config = q_config or get_default_per_tensor_custom_qconfig_dict()
    mapping = mapping or get_default_quant_module_mappings()
    if not inplace:
        import copy

        module = copy.deepcopy(module)
    skip = set(modules_to_not_convert or [])

    def _walk(parent: nn.Module, prefix: str) -> None:
        for name, child in list(parent.named_children()):
            full = f"{prefix}.{name}" if prefix else name
            if full in skip or name in skip:
                continue
            if include is not None and not any(full == i or full.endswith("." + i) or name == i for i in include):
                _walk(child, full)
                continue
            target = mapping.get(type(child))
            if target is not None:
                setattr(parent, name, target.from_float(child, q_config))
            else:
                _walk(child, full)

    _walk(module, "")
    return module
