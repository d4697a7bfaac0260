"""Uniform model wrapper (reference ``trainer/model.py:8-116``): ``run_train`` / ``run_eval`` work
the same with and without pipeline parallelism; ``local_*`` accessors return only what this
rank owns (all of it when pp == 1)."""
from __future__ import annotations

from typing import Any, Iterator, Tuple

import torch
from torch import nn


class NxDModel(nn.Module):
    def __init__(self, module: nn.Module, nxd_config: dict):
        super().__init__()
        self.module = module
        self.nxd_config = nxd_config
        from ..pipeline.model import NxDPPModel

        self.pp_enabled = isinstance(module, NxDPPModel)

    def __repr__(self) -> str:
        return "NxDModel(" + repr(self.module) + ")"

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def original_module(self) -> nn.Module:
        return self.module.original_torch_module if self.pp_enabled else self.module

    def run_train(self, *args, **kwargs):
        if self.pp_enabled:
            return self.module.run_train(*args, **kwargs)
        self.module.train()
        out = self.module(*args, **kwargs)
        loss = out[0] if isinstance(out, (tuple, list)) else (out.loss if hasattr(out, "loss") else out)
        loss.backward()
        return loss.detach()

    def run_eval(self, *args, **kwargs):
        if self.pp_enabled:
            return self.module.run_eval(*args, **kwargs)
        self.module.eval()
        with torch.no_grad():
            out = self.module(*args, **kwargs)
        return out[0] if isinstance(out, (tuple, list)) else (out.loss if hasattr(out, "loss") else out)

    # ---- local views -----------------------------------------------------------------
    def local_parameters(self, *a, **k) -> Iterator[nn.Parameter]:
        return self.module.local_parameters(*a, **k) if self.pp_enabled else self.module.parameters(*a, **k)

    def local_named_parameters(self, *a, **k) -> Iterator[Tuple[str, nn.Parameter]]:
        return self.module.local_named_parameters(*a, **k) if self.pp_enabled else self.module.named_parameters(*a, **k)

    def local_modules(self, *a, **k):
        return self.module.local_modules(*a, **k) if self.pp_enabled else self.module.modules(*a, **k)

    def local_named_modules(self, *a, **k):
        return self.module.local_named_modules(*a, **k) if self.pp_enabled else self.module.named_modules(*a, **k)

    def local_state_dict(self, *a, **k):
        return self.module.local_state_dict(*a, **k) if self.pp_enabled else self.module.state_dict(*a, **k)

    # nn.Module API scoped to the local partition
    def parameters(self, *a, **k):
        return self.local_parameters(*a, **k)

    def named_parameters(self, *a, **k):
        return self.local_named_parameters(*a, **k)

    def state_dict(self, *a, **k):
        return self.local_state_dict(*a, **k)

    def load_state_dict(self, state_dict, strict: bool = True):
        return self.module.load_state_dict(state_dict, strict=strict)

    # ---- Hugging Face conveniences, always answered by the user's module (also under pipeline parallelism; reference
    # trainer/model.py:106-116) ------------------------------------------------------------------------------------------
    @property
    def dtype(self):
        m = self.original_module()
        d = getattr(m, "dtype", None)
        if d is None:
            p = next(iter(m.parameters()), None)
            d = None if p is None else p.dtype
        return d

    @property
    def config(self):
        return self.original_module().config

    @property
    def name_or_path(self):
        return self.original_module().name_or_path

    def __getattr__(self, name: str) -> Any:
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("module"), name)
