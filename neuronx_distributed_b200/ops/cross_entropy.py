"""Vocab-parallel cross-entropy passes (kernels: ``csrc/elementwise.cu`` ce_stats/ce_backward).

``ce_stats``   : one read of the local logits → per-row (max, Σexp(x-max), target logit, Σx) fp32
``ce_backward``: grad = (softmax - onehot·(1-s) - s/V)·g, softmax recomputed from the saved row
                 log-normaliser; written in the logits dtype.
"""
from __future__ import annotations

import torch

from . import _ext


def ce_stats(logits2d: torch.Tensor, target: torch.Tensor, vocab_start: int) -> torch.Tensor:
    if _ext.use_cuda(logits2d) and logits2d.dtype in (torch.bfloat16, torch.float32, torch.float16):
        _ext.count_launch()
        return _ext.ext().ce_stats(logits2d.contiguous(), target.contiguous().long(), int(vocab_start))
    x = logits2d.float()
    vp = x.shape[-1]
    mx = x.max(dim=-1).values
    se = torch.exp(x - mx.unsqueeze(-1)).sum(-1)
    local = target.long() - vocab_start
    owned = (local >= 0) & (local < vp)
    idx = torch.where(owned, local, torch.zeros_like(local))
    tl = x.gather(1, idx.unsqueeze(1)).squeeze(1) * owned.float()
    return torch.stack([mx, se, tl, x.sum(-1)], dim=1)


def ce_backward(logits2d, target, lse, gout, vocab_start: int, smoothing: float, vocab: int) -> torch.Tensor:
    if _ext.use_cuda(logits2d) and logits2d.dtype in (torch.bfloat16, torch.float32, torch.float16):
        _ext.count_launch()
        return _ext.ext().ce_backward(
            logits2d.contiguous(), target.contiguous().long(), lse.contiguous().float(), gout.contiguous().float(),
            int(vocab_start), float(smoothing), int(vocab),
        )
    x = logits2d.float()
    vp = x.shape[-1]
    soft = torch.exp(x - lse.unsqueeze(-1))
    local = target.long() - vocab_start
    owned = (local >= 0) & (local < vp)
    idx = torch.where(owned, local, torch.zeros_like(local))
    onehot = torch.zeros_like(soft)
    onehot.scatter_(1, idx.unsqueeze(1), owned.float().unsqueeze(1))
    grad = soft - (1.0 - smoothing) * onehot - (smoothing / vocab if smoothing > 0 else 0.0)
    return (grad * gout.unsqueeze(-1)).to(logits2d.dtype)
