from . import post_partition_hooks  # noqa: F401
from . import post_partition_hooks as hooks  # noqa: F401  (round-1 module name)
from .model import NxDModel  # noqa: F401
from .optimizer import NxDOptimizer  # noqa: F401
from .trainer import (  # noqa: F401
    initialize_parallel_model,
    initialize_parallel_optimizer,
    neuronx_distributed_config,
)
