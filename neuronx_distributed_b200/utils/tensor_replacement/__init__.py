"""Overwrite a module's output with an injected tensor for numerical bisecting (reference
``utils/tensor_replacement/model_modification.py:14-79``): the hook returns
``torch.where(mask, injected, output)`` so the same program runs with replacement on or off."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import nn

from .model_modification import modify_model_for_tensor_replacement, patch_forward_with_additional_args  # noqa: F401
from .registry import RuntimeRegister  # noqa: F401


class TensorReplacementRegistry:
    _inst: Optional["TensorReplacementRegistry"] = None

    def __init__(self):
        self.replacements: Dict[str, torch.Tensor] = {}
        self.masks: Dict[str, torch.Tensor] = {}
        self.handles: List = []

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst


def enable_tensor_replacement(model: nn.Module, replacements: Dict[str, torch.Tensor],
                              masks: Optional[Dict[str, torch.Tensor]] = None) -> nn.Module:
    reg = TensorReplacementRegistry.get()
    named = dict(model.named_modules())
    for name, tensor in replacements.items():
        if name not in named:
            raise ValueError(f"module {name} not found")
        reg.replacements[name] = tensor
        if masks and name in masks:
            reg.masks[name] = masks[name]

        def hook(mod, inp, out, _name=name):
            rep = reg.replacements.get(_name)
            if rep is None:
                return out
            first = out[0] if isinstance(out, (tuple, list)) else out
            mask = reg.masks.get(_name)
            mask = torch.ones((), dtype=torch.bool, device=first.device) if mask is None else mask.to(first.device).bool()
            new = torch.where(mask, rep.to(first.device, first.dtype), first)
            if isinstance(out, tuple):
                return (new,) + tuple(out[1:])
            if isinstance(out, list):
                return [new] + list(out[1:])
            return new
        reg.handles.append(named[name].register_forward_hook(hook))
    return model


def disable_tensor_replacement() -> None:
    reg = TensorReplacementRegistry.get()
    for h in reg.handles:
        h.remove()
    reg.handles.clear()
    reg.replacements.clear()
    reg.masks.clear()
