"""MoE configuration objects (reference ``modules/moe/moe_configs.py:22-273``)."""
from __future__ import annotations

import enum
from dataclasses import dataclass
from typing import Any, NamedTuple, Optional

import torch


def to_torch_dtype(dtype_str) -> torch.dtype:
    """``"bfloat16"`` → ``torch.bfloat16`` (reference moe_configs.py:13-20); torch dtypes pass through."""
    if isinstance(dtype_str, torch.dtype):
        return dtype_str
    table = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}
    assert dtype_str in table, f"Unsupported dtype: {dtype_str}"
    return table[dtype_str]


def _from_kwargs(cls, kwargs: dict):
    """Build a config dataclass from a loose kwargs dict: known fields are consumed (popped), the rest is left for the
    caller — the reference's ``from_kwargs`` protocol (moe_configs.py), dtype strings accepted."""
    import dataclasses

    picked = {}
    for f in dataclasses.fields(cls):
        if f.name in kwargs:
            v = kwargs.pop(f.name)
            picked[f.name] = to_torch_dtype(v) if f.name == "dtype" and isinstance(v, str) else v
    return cls(**picked)


@dataclass
class RouterConfig:
    act_fn: str = "softmax"                 # "softmax" | "sigmoid"
    dtype: torch.dtype = torch.float32
    top_k: int = 2
    normalize_top_k_affinities: bool = True
    jitter_eps: float = 0.0
    bias: bool = False
    apply_act_fn_over_topk: bool = False

    @classmethod
    def from_kwargs(cls, **kwargs) -> "RouterConfig":
        kwargs = dict(kwargs)
        if "router_act_fn" in kwargs:
            kwargs.setdefault("act_fn", kwargs.pop("router_act_fn"))
        if "router_dtype" in kwargs:
            kwargs.setdefault("dtype", kwargs.pop("router_dtype"))
        return _from_kwargs(cls, kwargs)


# ---- kernel-side enums (reference: imported from the closed kernel library at blockwise.py:57-87 and re-exported there) ----
class ExpertAffinityScaleMode(enum.IntEnum):
    """Where the routing weight multiplies an expert's contribution (ints as documented at reference blockwise.py:906-909)."""
    NO_SCALE = 0
    POST_SCALE = 1          # after the down projection (default)
    PRE_SCALE = 2           # on the expert's input ("early affinity modulation")

    @classmethod
    def coerce(cls, v) -> "ExpertAffinityScaleMode":
        if isinstance(v, cls):
            return v
        if isinstance(v, str):
            return cls[v.upper()]
        return cls(int(v))


class BlockShardStrategy(enum.Enum):
    """How a 2-core kernel split the block list on the reference hardware.  The grouped GEMM here is one persistent launch
    over all blocks (tile scheduler), so both values run the same code; kept so configs round-trip."""
    HI_LO = "HI_LO"
    PING_PONG = "PING_PONG"


class SkipMode(NamedTuple):
    """Skip the loads of padded token rows / of weights for empty blocks.  The grouped GEMM never loads either (padding slots
    are masked in the gather, empty blocks are not in the block list), so the flags are accepted and ignored."""
    skip_token: bool = False
    skip_weight: bool = False


class ActFnType(enum.Enum):
    SiLU = "silu"
    GELU = "gelu"
    GELU_Tanh_Approx = "gelu_new"
    Swish = "sigmoid"        # x·σ(αx) with the configured scaling factor (the "swiglu" GLU type)

    def fn(self):
        import torch.nn.functional as F

        return {"silu": F.silu, "gelu": F.gelu, "gelu_new": lambda x: F.gelu(x, approximate="tanh"),
                "sigmoid": lambda x: x * torch.sigmoid(x)}[self.value]


ActivationFunction = ActFnType


class RouterActFnType(enum.IntEnum):
    """Router activation as the decode kernel numbers it (``csrc/moe_tkg.cu`` ``router_act``)."""
    SOFTMAX = 0
    SIGMOID = 1


ROUTER_ACT_FN_MAPPING = {"softmax": RouterActFnType.SOFTMAX, "sigmoid": RouterActFnType.SIGMOID}


@dataclass
class BlockwiseMatmulConfig:
    """Field order of reference moe_configs.py:22-99 (all required there; defaults = its ``default()`` values).  The flags that
    picked between the reference's kernel variants (2-core sharding, dynamic-while block loops, static block counts, padding
    the block count to even, the autograd class) have no effect: every variant is the one persistent grouped GEMM whose block
    list is data, not a compile-time count."""
    block_size: int = 512
    use_block_parallel: bool = False
    block_sharding_strategy: Any = "HI_LO"               # BlockShardStrategy or its name
    skip_dma_token: bool = False
    skip_dma_weight: bool = False
    logical_nc_config: int = 1
    blockwise_nki_autograd_cls: Any = None
    use_torch_block_wise: bool = False      # force the PyTorch reference instead of the CUDA grouped GEMM
    parallelize_token_to_block_mapping: bool = True
    optimized_block_to_token_mapping: bool = True
    always_augment_inputs_for_blockwise_matmul: bool = False
    use_shard_on_intermediate_dynamic_while: bool = False
    use_shard_on_block_dynamic_while: bool = False
    num_static_blocks: Optional[int] = None
    pad_num_blocks_to_even: bool = False

    @classmethod
    def default(cls) -> "BlockwiseMatmulConfig":
        return cls()

    @classmethod
    def from_kwargs(cls, **kwargs) -> "BlockwiseMatmulConfig":
        return _from_kwargs(cls, dict(kwargs))


@dataclass
class RoutedExpertsMLPOpsConfig:
    """Field order and defaults of reference moe_configs.py:102-160 (the first six are required there; the defaults here only
    add convenience).  ``*_actual`` / ``is_*_dim_shuffled`` describe a checkpoint whose hidden / intermediate dims were padded
    or permuted offline for the reference's kernels — recorded, not needed by the grouped GEMM."""
    num_experts: int = 8
    hidden_size: int = 1024
    intermediate_size: int = 4096
    top_k: int = 2
    hidden_act: str = "silu"
    glu_mlp: bool = True
    bias: bool = False
    glu_type: str = "glu"                    # "glu" | "swiglu" (or the GLUType enum)
    hidden_act_scaling_factor: float = 1.0
    hidden_act_bias: float = 0.0
    hidden_size_actual: Optional[int] = None
    intermediate_size_actual: Optional[int] = None
    is_hidden_dim_shuffled: Optional[bool] = None
    is_intermediate_dim_shuffled: Optional[bool] = None
    use_index_calc_kernel: bool = True
    gate_clamp_upper_limit: Optional[float] = None
    gate_clamp_lower_limit: Optional[float] = None
    up_clamp_upper_limit: Optional[float] = None
    up_clamp_lower_limit: Optional[float] = None
    normalize_top_k_affinities: bool = False
    early_expert_affinity_modulation: bool = False
    input_layer_init_method: Optional[object] = None
    output_layer_init_method: Optional[object] = None
    capacity_factor: Optional[float] = None  # None → dropless
    enable_spmd_rank: bool = False
    is_prefill: Optional[bool] = None
    expert_distribution: Optional[list] = None       # [ep_rank][slot] → logical expert id (redundant experts allowed)

    def __post_init__(self):
        self.local_redudancy_degree = self.bincount_2d(self.expert_distribution, self.num_experts) \
            if self.expert_distribution else None

    @staticmethod
    def bincount_2d(expert_distribution, num_experts: int):
        """``[ep_rank][expert]`` → how many replicas of each logical expert an EP rank hosts."""
        out = []
        for row in expert_distribution:
            counts = [0] * num_experts
            for e in row:
                counts[int(e)] += 1
            out.append(counts)
        return out


@dataclass
class MoEFusedTKGConfig:
    """Decode-time fused path (RMSNorm → router → experts → shared experts, reference K8; fields / defaults of reference
    moe_configs.py:236-273).  The ``*_kernel_enabled`` switches are tri-state: ``None`` = use the kernel whenever it applies."""
    quantized: bool = False
    moe_fused_kernel_enabled: Optional[bool] = None
    router_topk_kernel_enabled: Optional[bool] = None
    expert_mlp_kernel_enabled: Optional[bool] = None
    shared_mlp_kernel_enabled: Optional[bool] = None
    norm_topk_prob: bool = False
    is_mxfp4_compute: Optional[bool] = None
    router_mm_dtype: torch.dtype = torch.float32      # accumulation / output dtype of the router GEMV (the kernel keeps fp32)
