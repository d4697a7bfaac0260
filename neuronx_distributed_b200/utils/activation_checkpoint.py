"""Activation checkpointing by module predicate (reference ``utils/activation_checkpoint.py:20-83``).

Matching sub-modules are wrapped so their forward re-runs in backward.  The wrapper registers
state-dict hooks that strip its ``_checkpoint_wrapped_module.`` prefix, so checkpoints written
with and without activation checkpointing are interchangeable.  Recompute re-issues the fused
TP kernels (their collectives included), exactly as the reference's recompute re-issues its
collectives."""
from __future__ import annotations

from typing import Callable, Optional

import torch
from torch import nn
from torch.utils.checkpoint import checkpoint

_PREFIX = "_checkpoint_wrapped_module."


class NxDCheckpointWrapper(nn.Module):
    def __init__(self, module: nn.Module, checkpoint_impl=None, checkpoint_fn: Optional[Callable] = None, **checkpoint_fn_kwargs):
        super().__init__()
        self._checkpoint_wrapped_module = module
        self._checkpoint_fn, self._checkpoint_fn_kwargs = checkpoint_fn, checkpoint_fn_kwargs
        self._register_state_dict_hook(self._strip_prefix)
        self.register_load_state_dict_pre_hook(self._add_prefix)

    @staticmethod
    def _strip_prefix(module, state_dict, prefix, local_metadata):
        for k in list(state_dict.keys()):
            if k.startswith(prefix + _PREFIX):
                state_dict[prefix + k[len(prefix) + len(_PREFIX):]] = state_dict.pop(k)
        return state_dict

    @staticmethod
    def _add_prefix(module, state_dict, prefix, *args):
        for k in list(state_dict.keys()):
            if k.startswith(prefix) and not k.startswith(prefix + _PREFIX):
                state_dict[prefix + _PREFIX + k[len(prefix):]] = state_dict.pop(k)

    def forward(self, *args, **kwargs):
        if not torch.is_grad_enabled():
            return self._checkpoint_wrapped_module(*args, **kwargs)
        if self._checkpoint_fn is not None:
            return self._checkpoint_fn(self._checkpoint_wrapped_module, *args, **self._checkpoint_fn_kwargs, **kwargs)
        return checkpoint(self._checkpoint_wrapped_module, *args, use_reentrant=False, **kwargs)

    def named_modules(self, *args, **kwargs):
        """Module names without the wrapper's prefix, so name-based tools (LoRA targets, tensor capture, pipeline cuts) see
        the same names with and without activation checkpointing."""
        for name, m in super().named_modules(*args, **kwargs):
            yield name.replace(_PREFIX, ""), m              # children lose the prefix; the wrapped module keeps its own name

    def named_parameters(self, *args, **kwargs):
        for name, p in super().named_parameters(*args, **kwargs):
            yield name.replace(_PREFIX, ""), p

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self._checkpoint_wrapped_module, name)


def checkpoint_wrapper(module: nn.Module, checkpoint_impl=None, checkpoint_fn: Optional[Callable] = None,
                       **checkpoint_fn_kwargs) -> nn.Module:
    """Wrap ONE module (reference :31-52).  ``checkpoint_fn(module, *args, **kwargs)`` replaces the default
    non-reentrant ``torch.utils.checkpoint.checkpoint``; ``checkpoint_impl`` is accepted for signature compatibility
    (there is no XLA reentrancy restriction here)."""
    return NxDCheckpointWrapper(module, checkpoint_impl, checkpoint_fn, **checkpoint_fn_kwargs)


def apply_activation_checkpointing(model: nn.Module, checkpoint_wrapper_fn: Optional[Callable] = None,
                                   check_fn: Callable[[nn.Module], bool] = lambda _: True) -> None:
    """Wrap (in place) every sub-module for which ``check_fn`` is true."""
    wrap = checkpoint_wrapper_fn or checkpoint_wrapper

    def _recurse(parent: nn.Module):
        for name, child in list(parent.named_children()):
            if isinstance(child, NxDCheckpointWrapper):
                continue
            if check_fn(child):
                setattr(parent, name, wrap(child))
            else:
                _recurse(child)

    _recurse(model)
