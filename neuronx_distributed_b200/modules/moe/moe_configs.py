"""MoE configuration objects (reference ``modules/moe/moe_configs.py:22-273``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch


def to_torch_dtype(dtype_str) -> torch.dtype:
    """``"bfloat16"`` → ``torch.bfloat16`` (reference moe_configs.py:13-20); torch dtypes pass through."""
    if isinstance(dtype_str, torch.dtype):
        return dtype_str
    table = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}
    assert dtype_str in table, f"Unsupported dtype: {dtype_str}"
    return table[dtype_str]


@dataclass
class RouterConfig:
    act_fn: str = "softmax"                 # "softmax" | "sigmoid"
    dtype: torch.dtype = torch.float32
    top_k: int = 2
    normalize_top_k_affinities: bool = True
    jitter_eps: float = 0.0
    bias: bool = False
    apply_act_fn_over_topk: bool = False


@dataclass
class BlockwiseMatmulConfig:
    block_size: int = 512
    use_block_parallel: bool = False
    use_torch_block_wise: bool = False      # force the PyTorch reference instead of the CUDA grouped GEMM
    skip_dma_token: bool = False
    skip_dma_weight: bool = False
    logical_nc_config: int = 1
    parallelize_token_to_block_mapping: bool = True
    optimized_block_to_token_mapping: bool = True
    always_augment_inputs_for_blockwise_matmul: bool = False
    block_sharding_strategy: str = "HI_LO"

    @classmethod
    def default(cls) -> "BlockwiseMatmulConfig":
        return cls()


@dataclass
class RoutedExpertsMLPOpsConfig:
    num_experts: int = 8
    top_k: int = 2
    hidden_size: int = 1024
    intermediate_size: int = 4096
    hidden_act: str = "silu"
    glu_mlp: bool = True
    glu_type: str = "glu"                    # "glu" | "swiglu"
    bias: bool = False
    capacity_factor: Optional[float] = None  # None → dropless
    normalize_top_k_affinities: bool = True
    early_expert_affinity_modulation: bool = False
    hidden_act_scaling_factor: float = 1.0
    hidden_act_bias: float = 0.0
    gate_clamp_upper_limit: Optional[float] = None
    gate_clamp_lower_limit: Optional[float] = None
    up_clamp_upper_limit: Optional[float] = None
    up_clamp_lower_limit: Optional[float] = None
    enable_spmd_rank: bool = False
    input_layer_init_method: Optional[object] = None
    output_layer_init_method: Optional[object] = None


@dataclass
class MoEFusedTKGConfig:
    """Decode-time fused path (RMSNorm → router → experts → shared experts, reference K8)."""
    quantized: bool = False
    moe_fused_kernel_enabled: bool = True
    router_topk_kernel_enabled: bool = True
    expert_mlp_kernel_enabled: bool = True
    shared_mlp_kernel_enabled: bool = True
    norm_topk_prob: bool = True
    is_mxfp4_compute: bool = False
