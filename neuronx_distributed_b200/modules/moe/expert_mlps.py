"""First-generation constructor of the routed experts (reference ``modules/moe/expert_mlps.py:15-160``): every option is a flat
keyword of ``ExpertMLPs(...)``.  The options are sorted into the two config objects :class:`ExpertMLPsV2` takes (routed-expert
MLP ops / blockwise matmul); everything else configures the module itself (parallel groups, dtype / device, hybrid sharding)."""
from __future__ import annotations

import dataclasses
from typing import Optional

from .expert_mlps_v2 import ExpertMLPsV2
from .moe_configs import BlockwiseMatmulConfig, RoutedExpertsMLPOpsConfig, _from_kwargs

_MODULE_KEYWORDS = ("sequence_parallel_enabled", "dtype", "device", "return_bias", "tensor_model_parallel_group",
                    "expert_model_parallel_group", "tkg_tensor_model_parallel_group", "tkg_expert_model_parallel_group",
                    "cte_tensor_model_parallel_group", "cte_expert_model_parallel_group", "is_prefill", "enabled_hybrid_sharding")
# Trainium scheduling knobs of the blockwise NKI kernels: accepted so that ported configurations construct, nothing to select here
_IGNORED_KEYWORDS = ("blockwise_nki_autograd_cls", "use_shard_on_intermediate_dynamic_while", "use_shard_on_block_dynamic_while")


class ExpertMLPs(ExpertMLPsV2):
    def __init__(self, num_experts: int, top_k: int, hidden_size: int, intermediate_size: int, hidden_act: str, glu_mlp: bool,
                 capacity_factor: Optional[float], **options):
        options.update(num_experts=num_experts, top_k=top_k, hidden_size=hidden_size, intermediate_size=intermediate_size,
                       hidden_act=hidden_act, glu_mlp=glu_mlp, capacity_factor=capacity_factor)
        for k in _IGNORED_KEYWORDS:
            options.pop(k, None)
        if options.get("block_size") is None:                      # None = the default block size
            options.pop("block_size", None)
        glu_type = options.get("glu_type")
        if glu_type is not None and not isinstance(glu_type, str):  # GLUType enum → its value
            options["glu_type"] = glu_type.value
        strategy = options.get("block_sharding_strategy")
        if strategy is not None and not isinstance(strategy, str):
            options["block_sharding_strategy"] = getattr(strategy, "name", str(strategy))
        for old, new in (("init_method", "input_layer_init_method"),):
            if old in options:
                options.setdefault(new, options.pop(old))
        module_kw = {k: options.pop(k) for k in _MODULE_KEYWORDS if k in options}
        # `early_expert_affinity_modulation` belongs to the MLP ops; the reference also mirrors it into the blockwise config
        shared = {k: options[k] for k in ("early_expert_affinity_modulation",) if k in options}
        routed = _from_kwargs(RoutedExpertsMLPOpsConfig, options)
        options.update({k: v for k, v in shared.items() if k in {f.name for f in dataclasses.fields(BlockwiseMatmulConfig)}})
        blockwise = _from_kwargs(BlockwiseMatmulConfig, options)
        if options:
            raise TypeError(f"ExpertMLPs got unexpected keyword arguments: {sorted(options)}")
        super().__init__(routed_experts_mlp_config=routed, blockwise_matmul_config=blockwise, **module_kw)
