"""Rotary position embedding, HF "rotate-half" convention, applied in place to q and k
(kernel: ``csrc/elementwise.cu`` rope_apply).  Backward is the same rotation with −sin.
Reference: ``overrides/transformer_overrides.py:20-32``, ``modules/attention/utils.py:50-80``.

Layout: q/k ``[B, S, H, D]`` (any strides on B/S/H as long as D is contiguous); cos/sin ``[S, D/2]``
fp32 (position ``s`` of the *global* sequence; callers slice for CP/SP offsets).
"""
from __future__ import annotations

import torch

from . import _ext


def _rope_ref(x, cos, sin, sign):
    # x [B,S,H,D]; cos/sin [S, D/2]
    d2 = x.shape[-1] // 2
    xf = x.float()
    x1, x2 = xf[..., :d2], xf[..., d2:]
    c = cos[None, :, None, :].float()
    s = sin[None, :, None, :].float() * sign
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1).to(x.dtype)


class _Rope(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cos, sin):
        ctx.save_for_backward(cos, sin)
        if _ext.use_cuda(x, cos, sin) and x.dtype in (torch.bfloat16, torch.float16, torch.float32) \
                and x.stride(-1) == 1 and x.shape[-1] % 2 == 0:
            ctx.cuda = True
            _ext.count_launch()
            return _ext.ext().rope_apply(x, cos.contiguous().float(), sin.contiguous().float(), 1.0)
        ctx.cuda = False
        return _rope_ref(x, cos, sin, 1.0)

    @staticmethod
    def backward(ctx, g):
        cos, sin = ctx.saved_tensors
        if ctx.cuda:
            _ext.count_launch()
            return _ext.ext().rope_apply(g, cos.contiguous().float(), sin.contiguous().float(), -1.0), None, None
        return _rope_ref(g, cos, sin, -1.0), None, None


def apply_rotary(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    return _Rope.apply(x, cos, sin)


def rope_tables(seq_len: int, dim: int, base: float = 10000.0, device=None, offset: int = 0,
                scaling: float = 1.0):
    inv = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.float32, device=device) / dim))
    t = (torch.arange(offset, offset + seq_len, dtype=torch.float32, device=device)) / scaling
    fr = torch.outer(t, inv)
    return fr.cos(), fr.sin()
