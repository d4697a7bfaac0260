"""Model construction helpers: meta-device init, staggered materialisation, attribute
preservation (reference ``utils/model_utils.py:147-386``)."""
from __future__ import annotations

import contextlib
import math
from typing import Callable, Dict, Optional

import torch
import torch.distributed as dist
from torch import nn

from ..utils import cpu_mode, get_device
from ..utils.logger import get_logger


def move_model_to_device(model: nn.Module, device: Optional[torch.device] = None) -> nn.Module:
    """Reference ``utils/model_utils.py`` re-exports this helper; the implementation lives with the TP utilities."""
    from ..parallel_layers.utils import move_model_to_device as _move

    return _move(model, device)

logger = get_logger()

_PARALLEL_ATTRS = ("tensor_model_parallel", "partition_dim", "partition_stride", "num_partitions", "rank_ordering",
                   "sequence_parallel_enabled", "shared", "expert_model_parallel", "fused_qkv", "qkv_sections")


def is_hf_pretrained_model(model) -> bool:
    try:
        from transformers import PreTrainedModel

        return isinstance(model, PreTrainedModel)
    except Exception:
        return False


def is_nxd_pipeline_model(model) -> bool:
    try:
        from ..pipeline.model import NxDPPModel

        return isinstance(model, NxDPPModel)
    except Exception:
        return False


@contextlib.contextmanager
def init_on_device(device: torch.device, include_buffers: bool = False):
    """Create parameters (and optionally buffers) directly on ``device`` (typically ``meta``) while
    keeping the parallel attributes set by the TP layers (reference :147-213)."""
    old_register_parameter = nn.Module.register_parameter
    old_register_buffer = nn.Module.register_buffer

    def register_parameter(module, name, param):
        old_register_parameter(module, name, param)
        if param is not None:
            cur = module._parameters[name]
            attrs = dict(cur.__dict__)
            new = nn.Parameter(cur.to(device), requires_grad=cur.requires_grad)
            new.__dict__.update(attrs)
            module._parameters[name] = new

    def register_buffer(module, name, buf, persistent=True):
        old_register_buffer(module, name, buf, persistent=persistent)
        if buf is not None:
            module._buffers[name] = module._buffers[name].to(device)

    try:
        nn.Module.register_parameter = register_parameter
        if include_buffers:
            nn.Module.register_buffer = register_buffer
        yield
    finally:
        nn.Module.register_parameter = old_register_parameter
        nn.Module.register_buffer = old_register_buffer


def preserve_parallel_attributes(model: nn.Module) -> Dict[str, Dict]:
    return {n: {k: v for k, v in p.__dict__.items() if k in _PARALLEL_ATTRS} for n, p in model.named_parameters()}


def restore_parallel_attributes(model: nn.Module, saved: Dict[str, Dict]) -> None:
    for n, p in model.named_parameters():
        for k, v in saved.get(n, {}).items():
            setattr(p, k, v)


def get_tied_parameters(model: nn.Module):
    seen: Dict[int, str] = {}
    tied = []
    for name, p in model.named_parameters(remove_duplicate=False):
        if id(p) in seen:
            tied.append((seen[id(p)], name))
        else:
            seen[id(p)] = name
    return tied


def reinit_model(model: nn.Module, device: torch.device, param_init_fn: Optional[Callable]) -> None:
    """Materialise a meta-device model on ``device`` and (re-)initialise every module with
    ``param_init_fn(module, device)``; tied parameters stay tied (reference :276-332)."""
    saved = preserve_parallel_attributes(model)
    tied = get_tied_parameters(model)
    for module in model.modules():
        has_meta = any(p.device.type == "meta" for p in module.parameters(recurse=False)) or any(
            b.device.type == "meta" for b in module.buffers(recurse=False))
        if not has_meta:
            continue
        module.to_empty(device=device, recurse=False)
        if param_init_fn is not None:
            param_init_fn(module, device)
        elif hasattr(module, "reset_parameters"):
            module.reset_parameters()
    restore_parallel_attributes(model, saved)
    mods = dict(model.named_modules())
    for a, b in tied:
        pa, na = a.rsplit(".", 1) if "." in a else ("", a)
        pb, nb = b.rsplit(".", 1) if "." in b else ("", b)
        setattr(mods[pb], nb, getattr(mods[pa], na))


def get_model_sequential(model_fn: Callable[[], nn.Module], sequential_move_factor: int = 11,
                         move_to_device: bool = True) -> nn.Module:
    """Build (and optionally move) the model ``sequential_move_factor`` local ranks at a time so host
    memory holds at most that many full CPU copies at once (reference :335-356)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        model = model_fn()
        if move_to_device:
            model.to(get_device())
        return model
    local_rank = dist.get_rank() % max(1, torch.cuda.device_count() if not cpu_mode() else dist.get_world_size())
    local_world = torch.cuda.device_count() if not cpu_mode() else dist.get_world_size()
    waves = max(1, math.ceil(local_world / max(1, sequential_move_factor)))
    model = None
    for w in range(waves):
        if local_rank % waves == w:
            model = model_fn()
            if move_to_device:
                model.to(get_device())
        if waves > 1:
            dist.barrier()
    return model


def maybe_materalize_model(model: nn.Module) -> None:
    """Reference name (sic). Materialise any remaining meta tensors as empty storage."""
    for module in model.modules():
        if any(p.device.type == "meta" for p in module.parameters(recurse=False)):
            module.to_empty(device=get_device(), recurse=False)


def get_delay_tracing(nxd_config) -> bool:
    return False


def check_delay_tracing(nxd_config) -> bool:
    return False
