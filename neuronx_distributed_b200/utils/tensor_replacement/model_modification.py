"""Prepare a model for tensor replacement (reference ``utils/tensor_replacement/model_modification.py:14-79``).

The prepared model's ``forward`` takes ``2·k`` extra trailing positional arguments — ``k`` replacement tensors followed by
``k`` boolean masks, in ``RuntimeRegister.module_superset`` order — and every listed module's output becomes
``torch.where(mask, replacement, output)``.  Because replacement is data (masks), not control flow, ONE captured CUDA graph
serves "replace nothing", "replace layer 3", … : numerical bisection without re-capturing.
"""
from __future__ import annotations

import inspect
import types
from functools import wraps
from typing import Dict, List, Tuple

import torch
from torch import nn

from .registry import RuntimeRegister


def patch_forward_with_additional_args(model: nn.Module, module_superset: List[str]) -> nn.Module:
    original_forward = model.forward
    params = [p for p in inspect.signature(original_forward).parameters.values() if p.name != "self"]
    variadic = any(p.kind == p.VAR_POSITIONAL for p in params)
    n_orig = len([p for p in params if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)])
    k = len(module_superset)

    @wraps(original_forward)
    def patched_forward(self, *args):
        if variadic:                                      # cannot count the original args: the last 2k are ours
            orig, extra = args[: len(args) - 2 * k], args[len(args) - 2 * k:]
        else:
            orig, extra = args[:n_orig], args[n_orig:]
        if len(extra) != 2 * k:
            raise ValueError(f"[Tensor replacement] expected {2 * k} trailing replacement args (tensors[{k}] + masks[{k}]), "
                             f"got {len(extra)}.")
        RuntimeRegister.register_runtime_args(tr_args=list(extra[:k]), mask_args=list(extra[k:]))
        try:
            return original_forward(*orig)
        finally:
            RuntimeRegister.clear_runtime_args()          # drop the references so the buffers can be recycled

    model._nxd_tr_original_forward = original_forward
    model.forward = types.MethodType(patched_forward, model)
    return model


def modify_model_for_tensor_replacement(model: nn.Module) -> Tuple[nn.Module, Dict[str, object]]:
    targets = list(RuntimeRegister.module_superset)
    named = dict(model.named_modules())
    missing = [t for t in targets if t not in named]
    if missing:
        raise ValueError(f"[Tensor replacement] modules not found in the model: {missing}")
    model = patch_forward_with_additional_args(model, targets)

    def make_hook(name: str):
        def hook(module, inputs, output):
            t, m = RuntimeRegister._tr_runtime_list.get(name), RuntimeRegister._tr_mask_list.get(name)
            if t is None or m is None:
                raise ValueError(f"Need tensor and mask for tensor replacement. Got {t} and {m}")
            first = output[0] if isinstance(output, (tuple, list)) else output
            new = torch.where(m.to(first.device).bool(), t.to(first.device, first.dtype), first)
            if isinstance(output, tuple):
                return (new,) + tuple(output[1:])
            if isinstance(output, list):
                return [new] + list(output[1:])
            return new
        return hook

    hooks = {name: named[name].register_forward_hook(make_hook(name)) for name in targets}
    return model, hooks
