"""Attach / detach the capture hooks (reference ``utils/tensor_capture/model_modification.py:12-240``)."""
from __future__ import annotations

import dataclasses
import os
from typing import Any, Callable, List, Optional

import torch
from torch import nn

from ..logger import get_logger
from .registry import TensorRegistry

logger = get_logger()


def find_available_modules(model: nn.Module, prefix: str = "") -> List[str]:
    """Dotted names of every sub-module (the values ``modules_to_capture`` may contain)."""
    base = prefix + "." if prefix else ""
    return [base + n for n, _ in model.named_modules() if n]


def _walk(prefix: str, obj: Any, sink: Callable[[str, torch.Tensor], None]) -> None:
    """Register every tensor nested in ``obj``: tuples / lists by position, dicts by key, dataclass-like objects
    (e.g. HF ``ModelOutput``) by public attribute."""
    if isinstance(obj, torch.Tensor):
        sink(prefix, obj)
    elif isinstance(obj, (tuple, list)):
        for i, item in enumerate(obj):
            _walk(f"{prefix}.{i}", item, sink)
    elif isinstance(obj, dict):
        for k, v in obj.items():
            _walk(f"{prefix}.{k}", v, sink)
    elif dataclasses.is_dataclass(obj) or (hasattr(obj, "__dict__") and not isinstance(obj, (type, nn.Module))):
        for k, v in vars(obj).items():
            if not k.startswith("_") and isinstance(v, torch.Tensor):
                sink(f"{prefix}.{k}", v)


def _validate(model: nn.Module, modules_to_capture: List[str], strip_prefix: str = "") -> None:
    avail = set(find_available_modules(model))
    if strip_prefix:
        avail |= {n[len(strip_prefix):] for n in avail if n.startswith(strip_prefix)}
    bad = [m for m in modules_to_capture if m not in avail]
    if bad:
        raise ValueError(f"The following modules were not found in the model: {bad}")


def modify_model_for_tensor_capture(model: nn.Module, modules_to_capture: Optional[List[str]] = None,
                                    max_tensors: Optional[int] = None, capture_inputs: bool = False) -> nn.Module:
    modules_to_capture = list(modules_to_capture or [])
    _validate(model, modules_to_capture)
    registry = TensorRegistry.get_instance()
    registry.remove_hooks()
    registry.configure(enabled=True, modules=modules_to_capture, max_tensors=max_tensors, capture_inputs=capture_inputs)
    named = dict(model.named_modules())

    def make_hook(name: str, module: nn.Module):
        def hook(mod, args, kwargs, output):
            sink = lambda key, t: registry.register_tensor(key, t, module)  # noqa: E731
            if capture_inputs:
                if args:
                    _walk(f"{name}.inputs", tuple(args), sink)
                if kwargs:
                    _walk(f"{name}.inputs.kwargs", kwargs, sink)
            _walk(f"{name}.outputs", output, sink)
        return hook

    for name in modules_to_capture:
        registry.model_info.hooks.append(named[name].register_forward_hook(make_hook(name, named[name]), with_kwargs=True))
        logger.info("Registered forward hook for module %s for tensor capture", name)
    model._nxd_tensor_capture = True
    return model


def modify_hf_eager_model_for_tensor_capture(model: nn.Module, modules_to_capture: Optional[List[str]] = None,
                                             tensor_capture_save_dir: str = "", capture_inputs: bool = False) -> nn.Module:
    """Golden-side capture for an eager (HF) model: every forward call is one step — step 1 is the prompt (``cte``), later
    steps are decode (``tkg``) — and each monitored tensor is written to
    ``captured_tensors_<phase>_step_<n>_module_<name>.pt`` for comparison with the device run."""
    modules_to_capture = list(modules_to_capture or [])
    _validate(model, modules_to_capture, strip_prefix="model.")
    os.makedirs(tensor_capture_save_dir or ".", exist_ok=True)
    step = {"n": 1}
    original_forward = model.forward

    def patched_forward(*args, **kwargs):
        out = original_forward(*args, **kwargs)
        step["n"] += 1
        return out

    model.forward = patched_forward
    model._nxd_original_forward = original_forward

    def make_hook(name: str):
        def save(key: str, t: torch.Tensor) -> None:
            phase = "cte" if step["n"] == 1 else "tkg"
            path = os.path.join(tensor_capture_save_dir, f"captured_tensors_{phase}_step_{step['n']}_module_{key}.pt")
            torch.save(t.detach().cpu(), path)

        def hook(mod, args, kwargs, output):
            if capture_inputs and args:
                _walk(f"{name}.inputs", tuple(args), save)
            _walk(f"{name}.outputs", output, save)
        return hook

    named = dict(model.named_modules())
    hooks = []
    for name in modules_to_capture:
        target = named.get(name, named.get("model." + name))
        hooks.append(target.register_forward_hook(make_hook(name), with_kwargs=True))
    model._nxd_capture_hooks = hooks
    return model


def restore_model(model: nn.Module) -> nn.Module:
    TensorRegistry.get_instance().remove_hooks()
    for h in getattr(model, "_nxd_capture_hooks", []):
        h.remove()
    if hasattr(model, "_nxd_original_forward"):
        model.forward = model._nxd_original_forward
        del model._nxd_original_forward
    for attr in ("_nxd_capture_hooks", "_nxd_tensor_capture"):
        if hasattr(model, attr):
            delattr(model, attr)
    return model
