"""Float layer → quantised layer table used by ``quantize.convert`` (reference ``quantization_mappings.py:11-21``)."""
from __future__ import annotations

from typing import Dict


def get_default_quant_module_mappings() -> Dict[type, type]:
    from ..modules.moe.moe_parallel_layers import ExpertFusedColumnParallelLinear, ExpertFusedRowParallelLinear
    from ..parallel_layers.layers import ColumnParallelLinear, RowParallelLinear
    from .quantization_layers import (QuantizedColumnParallel, QuantizedExpertFusedColumnParallel,
                                      QuantizedExpertFusedRowParallel, QuantizedRowParallel)

    return {
        ColumnParallelLinear: QuantizedColumnParallel,
        RowParallelLinear: QuantizedRowParallel,
        ExpertFusedColumnParallelLinear: QuantizedExpertFusedColumnParallel,
        ExpertFusedRowParallelLinear: QuantizedExpertFusedRowParallel,
    }
