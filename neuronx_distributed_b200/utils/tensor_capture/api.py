"""Public tensor-capture API (reference ``utils/tensor_capture/api.py:16-95``)."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import nn

from ..logger import get_logger
from .model_modification import find_available_modules, modify_model_for_tensor_capture, restore_model
from .registry import TensorRegistry

logger = get_logger()


def enable_tensor_capture(model: nn.Module, modules_to_capture: Optional[List[str]] = None, max_tensors: Optional[int] = None,
                          capture_inputs: bool = False) -> nn.Module:
    """Hook ``modules_to_capture`` (dotted names; ``ValueError`` if one does not exist).  ``max_tensors`` is the budget of
    manually registered tensors (``None``: manual registration is ignored)."""
    return modify_model_for_tensor_capture(model, modules_to_capture, max_tensors, capture_inputs)


def disable_tensor_capture(model: Optional[nn.Module] = None) -> Optional[nn.Module]:
    reg = TensorRegistry.get_instance()
    reg.clear()
    reg.enabled = False
    return restore_model(model) if model is not None else None


def get_available_modules(model: nn.Module) -> List[str]:
    return find_available_modules(model)


def register_tensor(name: str, tensor: torch.Tensor) -> None:
    """Call from inside model code to expose an intermediate value (no-op unless capture is enabled)."""
    TensorRegistry.get_instance().register_tensor(name, tensor)


def get_captured_tensors_dict() -> Dict[str, torch.Tensor]:
    return TensorRegistry.get_instance().get_captured_tensors_dict()


def get_captured_tensors(clear: bool = True) -> Dict[str, torch.Tensor]:
    """Captured tensors as a plain dict; ``clear`` drops the values (hooks stay) so the next step starts empty."""
    reg = TensorRegistry.get_instance()
    out = dict(reg.get_captured_tensors_dict())
    if clear:
        reg.reset_tensors()
    return out
