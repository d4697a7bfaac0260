from .expert_mlps import ExpertMLPs  # noqa: F401
from .expert_mlps_v2 import ExpertMLPsV2  # noqa: F401
from .experts import ACT2FN, Experts  # noqa: F401
from .loss_function import load_balancing_loss_func  # noqa: F401
from .model import MoE  # noqa: F401
from .moe_configs import BlockwiseMatmulConfig, MoEFusedTKGConfig, RoutedExpertsMLPOpsConfig, RouterConfig  # noqa: F401
from .routing import GroupLimitedRouter, RouterSinkhorn, RouterTopK  # noqa: F401
from .shared_experts import SharedExperts  # noqa: F401
