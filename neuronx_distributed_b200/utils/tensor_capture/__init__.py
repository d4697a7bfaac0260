"""Capture named intermediate tensors from a model run (reference ``utils/tensor_capture/api.py:16-95``,
``model_modification.py:12-93``): forward hooks on the requested sub-modules stash their outputs in a registry;
``get_captured_tensors`` returns and clears them.  Works in eager and under CUDA-graph capture (the stashed
tensors are graph outputs that stay valid until the next replay)."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import nn


class CapturedTensorRegistry:
    _inst: Optional["CapturedTensorRegistry"] = None

    def __init__(self):
        self.tensors: Dict[str, torch.Tensor] = {}
        self.manual: Dict[str, torch.Tensor] = {}
        self.handles: List = []
        self.max_tensors: Optional[int] = None

    @classmethod
    def get(cls) -> "CapturedTensorRegistry":
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def clear(self) -> None:
        self.tensors.clear()
        self.manual.clear()


def enable_tensor_capture(model: nn.Module, modules_to_capture: List[str], max_tensors: Optional[int] = None,
                          capture_inputs: bool = False) -> nn.Module:
    reg = CapturedTensorRegistry.get()
    reg.max_tensors = max_tensors
    named = dict(model.named_modules())
    missing = [m for m in modules_to_capture if m not in named]
    if missing:
        raise ValueError(f"modules not found for capture: {missing}")
    for name in modules_to_capture:
        def hook(mod, inp, out, _name=name):
            t = out[0] if isinstance(out, (tuple, list)) else out
            if isinstance(t, torch.Tensor):
                reg.tensors[f"{_name}.outputs"] = t.detach()
            if capture_inputs and inp and isinstance(inp[0], torch.Tensor):
                reg.tensors[f"{_name}.inputs"] = inp[0].detach()
        reg.handles.append(named[name].register_forward_hook(hook))
    return model


def disable_tensor_capture(model: Optional[nn.Module] = None) -> None:
    reg = CapturedTensorRegistry.get()
    for h in reg.handles:
        h.remove()
    reg.handles.clear()
    reg.clear()


def register_tensor(name: str, tensor: torch.Tensor) -> None:
    """Manually register a tensor from inside model code."""
    reg = CapturedTensorRegistry.get()
    if reg.max_tensors is None or len(reg.manual) < reg.max_tensors:
        reg.manual[name] = tensor.detach()


def get_captured_tensors(clear: bool = True) -> Dict[str, torch.Tensor]:
    reg = CapturedTensorRegistry.get()
    out = {**reg.tensors, **reg.manual}
    if clear:
        reg.clear()
    return out


def get_available_modules(model: nn.Module) -> List[str]:
    return [n for n, _ in model.named_modules() if n]
