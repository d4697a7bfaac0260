"""Model-parallel layer interface (re-exports mirror reference ``parallel_layers/__init__.py:7-37``)."""
from typing import Callable, List, Type

import torch

from . import parallel_state  # noqa: F401
from . import comm, grads, layer_norm, layers, loss_functions, mappings, random, utils  # noqa: F401
from .grads import clip_grad_norm  # noqa: F401
from .layers import ColumnParallelLinear, ParallelEmbedding, RowParallelLinear  # noqa: F401
from .loss_functions import fused_linear_cross_entropy, parallel_cross_entropy  # noqa: F401
from .mappings import (  # noqa: F401
    copy_to_tensor_model_parallel_region,
    gather_from_tensor_model_parallel_region,
    reduce_from_tensor_model_parallel_region,
    scatter_to_tensor_model_parallel_region,
)
from .parallel_state import initialize_model_parallel  # noqa: F401
from .random import get_rng_tracker, get_xla_rng_tracker, model_parallel_manual_seed, model_parallel_xla_manual_seed  # noqa: F401
from .utils import (  # noqa: F401
    copy_tensor_model_parallel_attributes,
    move_model_to_device,
    set_defaults_if_not_set_tensor_model_parallel_attributes,
    set_tensor_model_parallel_attributes,
    split_tensor_along_last_dim,
)


def __getattr__(name):
    if name in ("load", "save", "checkpointing"):
        import importlib

        mod = importlib.import_module(".checkpointing", __name__)
        return mod if name == "checkpointing" else getattr(mod, name)
    if name == "pad":
        import importlib

        return importlib.import_module(".pad", __name__)
    raise AttributeError(name)


PARALLEL_MODULES: List[Type[torch.nn.Module]] = [ColumnParallelLinear, RowParallelLinear, ParallelEmbedding]
PARALLEL_FUNCTIONS: List[Callable] = [
    parallel_cross_entropy,
    fused_linear_cross_entropy,
    copy_to_tensor_model_parallel_region,
    gather_from_tensor_model_parallel_region,
    reduce_from_tensor_model_parallel_region,
    scatter_to_tensor_model_parallel_region,
]
