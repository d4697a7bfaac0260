from . import hooks  # noqa: F401
from .model import NxDModel  # noqa: F401
from .optimizer import NxDOptimizer  # noqa: F401
from .trainer import (  # noqa: F401
    initialize_parallel_model,
    initialize_parallel_optimizer,
    neuronx_distributed_config,
)
