"""RoPE utilities (reference ``modules/attention/utils.py:20-80``): Llama-3 frequency scaling,
``precompute_freqs_cis`` and a polar-compatible (interleaved-pair) apply."""
from __future__ import annotations

import math
from typing import Tuple

import torch


def apply_scaling(freqs: torch.Tensor, scale_factor: float = 8.0, low_freq_factor: float = 1.0,
                  high_freq_factor: float = 4.0, old_context_len: int = 8192) -> torch.Tensor:
    """Llama-3.1 long-context frequency rescaling."""
    low_wl = old_context_len / low_freq_factor
    high_wl = old_context_len / high_freq_factor
    wl = 2 * math.pi / freqs
    smooth = (old_context_len / wl - low_freq_factor) / (high_freq_factor - low_freq_factor)
    scaled = torch.where(wl > low_wl, freqs / scale_factor,
                         torch.where(wl < high_wl, freqs, (1 - smooth) * freqs / scale_factor + smooth * freqs))
    return scaled


def precompute_freqs_cis(dim: int, end: int, theta: float = 10000.0, use_scaled: bool = False,
                         device=None) -> torch.Tensor:
    """``[end, dim/2, 2]`` (cos, sin) table."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, device=device)[: dim // 2].float() / dim))
    if use_scaled:
        freqs = apply_scaling(freqs)
    t = torch.arange(end, device=device, dtype=torch.float32)
    f = torch.outer(t, freqs)
    return torch.stack([f.cos(), f.sin()], dim=-1)


def apply_rotary_polar_compatible(xq: torch.Tensor, xk: torch.Tensor, freqs_cis: torch.Tensor
                                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Interleaved-pair rotation (Meta checkpoint convention): x[..., 2i], x[..., 2i+1] rotate together.
    ``xq/xk``: [B, S, H, D]; ``freqs_cis``: [S, D/2, 2]."""
    def rot(x):
        xf = x.float().reshape(*x.shape[:-1], -1, 2)
        c = freqs_cis[None, : x.shape[1], None, :, 0]
        s = freqs_cis[None, : x.shape[1], None, :, 1]
        o = torch.stack([xf[..., 0] * c - xf[..., 1] * s, xf[..., 0] * s + xf[..., 1] * c], dim=-1)
        return o.flatten(-2).to(x.dtype)

    return rot(xq), rot(xk)
