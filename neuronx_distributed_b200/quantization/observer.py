"""Calibration observers (reference ``quantization/observer.py:12-166``)."""
from __future__ import annotations

from typing import Tuple

import torch


class PerChannelAbsMaxObserver(torch.nn.Module):
    """Running per-channel ``max |x|`` → symmetric scales ``absmax / quant_max`` and zero zero-points.

    ``with_args`` mirrors the torch.ao observer factory protocol so the class drops into ``QConfig(weight=…)``."""

    def __init__(self, ch_axis: int = 0, dtype=torch.qint8, qscheme=torch.per_channel_symmetric, quant_min=None,
                 quant_max=None, eps: float = torch.finfo(torch.float32).eps, **_unused):
        super().__init__()
        self.ch_axis, self.dtype, self.qscheme, self.eps = ch_axis, dtype, qscheme, eps
        if qscheme not in (torch.per_channel_symmetric, torch.per_tensor_symmetric):
            raise ValueError(f"Only support {torch.per_tensor_symmetric} and {torch.per_channel_symmetric}")
        default_max = {torch.qint8: 127, torch.int8: 127}.get(dtype)
        if default_max is None:
            default_max = torch.finfo(dtype).max
        self.quant_max = quant_max if quant_max is not None else default_max
        self.quant_min = quant_min if quant_min is not None else -self.quant_max
        self.register_buffer("max_val", torch.tensor([]))

    @classmethod
    def with_args(cls, **kwargs):
        from functools import partial

        factory = partial(cls, **kwargs)
        factory.with_args = lambda **kw: cls.with_args(**{**kwargs, **kw})  # type: ignore[attr-defined]
        return factory

    def forward(self, x_orig: torch.Tensor) -> torch.Tensor:
        if x_orig.numel() == 0:
            return x_orig
        x = x_orig.detach().float()
        dims = [d for d in range(x.dim()) if d != self.ch_axis % x.dim()]
        cur = x.abs().amax(dim=dims) if dims else x.abs()
        self.max_val = cur if self.max_val.numel() == 0 else torch.maximum(self.max_val, cur)
        return x_orig

    @property
    def abs_max(self) -> torch.Tensor:
        return self.max_val

    def calculate_qparams(self) -> Tuple[torch.Tensor, torch.Tensor]:
        scale = (self.max_val / float(self.quant_max)).clamp(min=self.eps).reshape(-1)
        return scale, torch.zeros_like(scale, dtype=torch.int64)

    def reset_min_max_vals(self) -> None:
        self.max_val = torch.tensor([], device=self.max_val.device)

    def extra_repr(self) -> str:
        return f"abs_max_val={self.max_val}"
