"""Causal self-attention (flash style).  Kernel: ``csrc/attention_sm100.cu`` (tcgen05 QK^T / PV with
TMEM accumulators, online softmax) when built; otherwise the PyTorch SDPA library call, which is
flagged as a *library* path in DESIGN.md.  Reference call sites: ``kernels/flash_attn.py:162-212``
(``nki_flash_attn_func``), ``modeling_llama_nxd.py:468-475``.

Layout: q ``[B, S, Hq, D]``, k/v ``[B, S, Hkv, D]`` → out ``[B, S, Hq, D]``.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn.functional as F

from . import _ext


def _sdpa(q, k, v, causal: bool, scale: Optional[float]):
    hq, hkv = q.shape[2], k.shape[2]
    qt, kt, vt = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    if hq != hkv:
        kt = kt.repeat_interleave(hq // hkv, dim=1)
        vt = vt.repeat_interleave(hq // hkv, dim=1)
    o = F.scaled_dot_product_attention(qt, kt, vt, is_causal=causal, scale=scale)
    return o.transpose(1, 2)


def _own_kernel_ok(q, k, v) -> bool:
    if os.environ.get("NXD_DISABLE_OWN_ATTENTION", "0") == "1":
        return False
    if not (q.is_cuda and q.dtype == torch.bfloat16 and q.shape[-1] == 128):
        return False
    e = _ext.ext()
    return e is not None and hasattr(e, "flash_attn_fwd") and q.shape[1] % 128 == 0 and k.shape[1] % 128 == 0


class _FlashAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, causal, scale):
        e = _ext.ext()
        _ext.count_launch()
        o, lse = e.flash_attn_fwd(q.contiguous(), k.contiguous(), v.contiguous(), bool(causal), float(scale))
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.causal, ctx.scale = causal, scale
        return o

    @staticmethod
    def backward(ctx, go):
        q, k, v, o, lse = ctx.saved_tensors
        e = _ext.ext()
        if hasattr(e, "flash_attn_bwd"):
            _ext.count_launch(2)
            dq, dk, dv = e.flash_attn_bwd(go.contiguous(), q, k, v, o, lse, bool(ctx.causal), float(ctx.scale))
            return dq, dk, dv, None, None
        # recompute through the library for the gradient until the bwd kernel lands
        with torch.enable_grad():
            qq, kk, vv = (t.detach().requires_grad_(True) for t in (q, k, v))
            oo = _sdpa(qq, kk, vv, ctx.causal, ctx.scale)
            dq, dk, dv = torch.autograd.grad(oo, (qq, kk, vv), go)
        return dq, dk, dv, None, None


def flash_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = True,
                    scale: Optional[float] = None) -> torch.Tensor:
    scale = scale if scale is not None else 1.0 / math.sqrt(q.shape[-1])
    if _own_kernel_ok(q, k, v):
        return _FlashAttn.apply(q, k, v, causal, scale)
    return _sdpa(q, k, v, causal, scale)


# reference-compatible names (kernels/flash_attn.py:162, kernels/ring_attention_kernel.py:118)
def nki_flash_attn_func(q, k, v, lnc: int = 1, dropout_p: float = 0.0, softmax_scale=None, causal: bool = True,
                        transpose_nki_inputs: bool = True):
    """Reference layout is ``[B, H, S, D]``; returns the same layout."""
    del lnc, transpose_nki_inputs
    assert dropout_p == 0.0, "attention dropout is not supported by the fused kernel"
    o = flash_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), causal, softmax_scale)
    return o.transpose(1, 2)
