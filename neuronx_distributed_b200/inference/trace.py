"""v0 inference API names (reference ``trace/trace.py:242-371``): ``parallel_model_trace / save / load``.  On B200 a
"traced parallel model" is the per-rank module plus its captured bucket programs; every rank is its own process
(torchrun) instead of being spawned by the library."""
from __future__ import annotations

import os
from typing import Any, Callable, Tuple

import torch

from .model_builder import ModelBuilder
from .nxd_model import NxDModel


def parallel_model_trace(func: Callable[[], Tuple[torch.nn.Module, Any]], example_inputs: Any, tp_degree: int = 1,
                         **kwargs) -> NxDModel:
    module, _aliases = func()
    ex = example_inputs if isinstance(example_inputs, (tuple, list)) else (example_inputs,)
    return ModelBuilder(tp_degree=tp_degree).add("main", module, [tuple(ex)]).trace()


def parallel_model_save(model: NxDModel, save_dir: str) -> None:
    model.save(save_dir, save_weights=True)


def parallel_model_load(load_dir: str) -> Any:
    return torch.load(os.path.join(load_dir, "nxd_model_meta.pt"), weights_only=False)
