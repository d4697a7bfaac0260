"""RoPE utilities (reference ``modules/attention/utils.py:20-80``): Llama-3 frequency scaling,
``precompute_freqs_cis`` and a polar-compatible (interleaved-pair) apply."""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch


# Llama-3.1 defaults, keyed like the ``rope_scaling`` dict of a Hugging Face config (reference utils.py:13-18)
ROPE_DEFAULTS = {"factor": 8, "low_freq_factor": 1, "high_freq_factor": 4, "original_max_position_embeddings": 8192}


def apply_scaling(freqs: torch.Tensor, scale_factor: Optional[float] = None, low_freq_factor: Optional[float] = None,
                  high_freq_factor: Optional[float] = None, old_context_len: Optional[int] = None, **rope_scaling) -> torch.Tensor:
    """Llama-3.1 long-context frequency rescaling.  Parameters come from (in this order) the explicit arguments, the
    ``rope_scaling`` keys of a Hugging Face config passed as keywords (``factor``, ``low_freq_factor``, ``high_freq_factor``,
    ``original_max_position_embeddings``), ``ROPE_DEFAULTS``."""
    p = {**ROPE_DEFAULTS, **{k: v for k, v in rope_scaling.items() if k in ROPE_DEFAULTS}}
    scale_factor = p["factor"] if scale_factor is None else scale_factor
    low_freq_factor = p["low_freq_factor"] if low_freq_factor is None else low_freq_factor
    high_freq_factor = p["high_freq_factor"] if high_freq_factor is None else high_freq_factor
    old_context_len = p["original_max_position_embeddings"] if old_context_len is None else old_context_len
    low_wl = old_context_len / low_freq_factor
    high_wl = old_context_len / high_freq_factor
    wl = 2 * math.pi / freqs
    smooth = (old_context_len / wl - low_freq_factor) / (high_freq_factor - low_freq_factor)
    scaled = torch.where(wl > low_wl, freqs / scale_factor,
                         torch.where(wl < high_wl, freqs, (1 - smooth) * freqs / scale_factor + smooth * freqs))
    return scaled


def precompute_freqs_cis(dim: int, end: int, theta: float = 10000.0, use_scaled: bool = False,
                         device=None, **rope_scaling) -> torch.Tensor:
    """``[end, dim/2, 2]`` (cos, sin) table; ``rope_scaling`` keywords as for :func:`apply_scaling`."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, device=device)[: dim // 2].float() / dim))
    if use_scaled:
        freqs = apply_scaling(freqs, **rope_scaling)
    t = torch.arange(end, device=device, dtype=torch.float32)
    f = torch.outer(t, freqs)
    return torch.stack([f.cos(), f.sin()], dim=-1)


def apply_rotary_polar_compatible(query: torch.Tensor, key: torch.Tensor, freqs_cis: torch.Tensor
                                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Interleaved-pair rotation (Meta checkpoint convention): x[..., 2i], x[..., 2i+1] rotate together.
    ``xq/xk``: [B, S, H, D]; ``freqs_cis``: [S, D/2, 2]."""
    xq, xk = query, key      # reference parameter names in the signature
    def rot(x):
        xf = x.float().reshape(*x.shape[:-1], -1, 2)
        c = freqs_cis[None, : x.shape[1], None, :, 0]
        s = freqs_cis[None, : x.shape[1], None, :, 1]
        o = torch.stack([xf[..., 0] * c - xf[..., 1] * s, xf[..., 0] * s + xf[..., 1] * c], dim=-1)
        return o.flatten(-2).to(x.dtype)

    return rot(xq), rot(xk)
