"""Weight / activation quantisation math (symmetric absmax; reference ``quantization_utils.py`` + ``observer.py``)."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch


_QMAX = {torch.int8: 127.0, torch.float8_e4m3fn: 448.0, torch.float8_e5m2: 57344.0}


def qmax(dtype: torch.dtype) -> float:
    return _QMAX[dtype]


def _cast(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    if dtype == torch.int8:
        return x.round().clamp(-127, 127).to(torch.int8)
    return x.clamp(-_QMAX[dtype], _QMAX[dtype]).to(dtype)


def quantize_per_tensor(w: torch.Tensor, dtype: torch.dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    scale = (w.abs().max().float() / _QMAX[dtype]).clamp(min=1e-12)
    return _cast(w.float() / scale, dtype), scale.reshape(1)


def quantize_per_channel(w: torch.Tensor, dtype: torch.dtype, axis: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """One scale per index of ``axis`` (scale keeps the weight's rank with size 1 on reduced dims)."""
    dims = [d for d in range(w.dim()) if d != axis % w.dim()]
    scale = (w.abs().amax(dim=dims, keepdim=True).float() / _QMAX[dtype]).clamp(min=1e-12)
    return _cast(w.float() / scale, dtype), scale


def quantize_blockwise(w: torch.Tensor, dtype: torch.dtype, block_axis: Sequence[int], block_size: Sequence[int]
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Independent scale per block of ``block_size[i]`` elements along ``block_axis[i]``."""
    shape = list(w.shape)
    view, red = [], []
    for d, n in enumerate(shape):
        if d in block_axis:
            b = block_size[list(block_axis).index(d)]
            assert n % b == 0, f"dim {d} ({n}) not divisible by block {b}"
            view += [n // b, b]
            red.append(len(view) - 1)
        else:
            view.append(n)
    wv = w.float().reshape(view)
    scale = (wv.abs().amax(dim=red, keepdim=True) / _QMAX[dtype]).clamp(min=1e-12)
    q = _cast(wv / scale, dtype).reshape(shape)
    return q, scale.squeeze(red) if red else scale


def dequantize_blockwise(q: torch.Tensor, scale: torch.Tensor, block_axis: Sequence[int], block_size: Sequence[int],
                         dtype: torch.dtype) -> torch.Tensor:
    s = scale
    for d in sorted(block_axis):
        s = s.repeat_interleave(block_size[list(block_axis).index(d)], dim=d)
    return (q.float() * s).to(dtype)


def quantize_activation_dynamic(x: torch.Tensor, dtype: torch.dtype = torch.float8_e4m3fn, clamp_bound: Optional[float] = None
                                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-row (last dim) dynamic quantisation: ``x ≈ q · scale[..., None]``."""
    xf = x.float()
    if clamp_bound is not None:
        xf = xf.clamp(-clamp_bound, clamp_bound)
    scale = (xf.abs().amax(dim=-1, keepdim=True) / _QMAX[dtype]).clamp(min=1e-12)
    return _cast(xf / scale, dtype), scale


class PerChannelAbsMaxObserver(torch.nn.Module):
    """Running per-channel absmax for static calibration (reference ``observer.py:12``)."""

    def __init__(self, ch_axis: int = 0, dtype: torch.dtype = torch.int8):
        super().__init__()
        self.ch_axis, self.dtype = ch_axis, dtype
        self.register_buffer("abs_max", torch.tensor([]))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        dims = [d for d in range(x.dim()) if d != self.ch_axis % x.dim()]
        cur = x.detach().abs().amax(dim=dims).float()
        self.abs_max = cur if self.abs_max.numel() == 0 else torch.maximum(self.abs_max, cur)
        return x

    def calculate_qparams(self) -> torch.Tensor:
        return (self.abs_max / _QMAX[self.dtype]).clamp(min=1e-12)
