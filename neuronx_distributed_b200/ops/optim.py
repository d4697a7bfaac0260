"""Multi-tensor optimizer kernels (``csrc/optim.cu``): squared-norm, scale, fused AdamW.

Each op takes python lists of tensors and launches ONE kernel over a device-side table of
(pointer, numel) chunks — no per-tensor launches, no host sync.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import torch

from . import _ext


def _all_cuda(ts: Sequence[torch.Tensor]) -> bool:
    return len(ts) > 0 and all(t.is_cuda for t in ts)


def multi_tensor_sq_norm(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    """Σ_t ‖t‖² as a 0-d fp32 tensor on the tensors' device."""
    tensors = [t for t in tensors if t.numel() > 0]
    if not tensors:
        return torch.zeros((), dtype=torch.float32)
    if _all_cuda(tensors) and _ext.use_cuda(*tensors):
        out = torch.zeros((), dtype=torch.float32, device=tensors[0].device)
        by_dtype = {}
        for t in tensors:
            by_dtype.setdefault(t.dtype, []).append(t if t.is_contiguous() else t.contiguous())
        for dt, ts in by_dtype.items():
            _ext.count_launch()
            _ext.ext().multi_tensor_sq_norm(ts, out)
        return out
    total = torch.zeros((), dtype=torch.float32, device=tensors[0].device)
    for t in tensors:
        total = total + t.float().pow(2).sum()
    return total


def multi_tensor_scale_(tensors: Sequence[torch.Tensor], scale: Union[float, torch.Tensor]) -> None:
    """In-place ``t *= scale`` for every tensor; ``scale`` may be a device scalar."""
    tensors = [t for t in tensors if t.numel() > 0]
    if not tensors:
        return
    if _all_cuda(tensors) and all(t.is_contiguous() for t in tensors) and _ext.use_cuda(*tensors):
        dev = tensors[0].device
        s = scale if isinstance(scale, torch.Tensor) else torch.full((), float(scale), dtype=torch.float32, device=dev)
        s = s.to(device=dev, dtype=torch.float32).reshape(())
        by_dtype = {}
        for t in tensors:
            by_dtype.setdefault(t.dtype, []).append(t)
        for ts in by_dtype.values():
            _ext.count_launch()
            _ext.ext().multi_tensor_scale(ts, s)
        return
    for t in tensors:
        t.mul_(scale.to(t.dtype) if isinstance(scale, torch.Tensor) else scale)


def fused_adamw_(
    params: List[torch.Tensor],      # fp32 master (or the params themselves if fp32)
    grads: List[torch.Tensor],       # fp32 or bf16
    exp_avgs: List[torch.Tensor],
    exp_avg_sqs: List[torch.Tensor],
    lr: float,
    beta1: float,
    beta2: float,
    eps: float,
    weight_decay: float,
    step: int,
    grad_scale: Optional[torch.Tensor] = None,   # device scalar multiplied into grads (clip coeff)
    model_params: Optional[List[torch.Tensor]] = None,  # optional low-precision copies to refresh
    hf_form: bool = True,
    correct_bias: bool = True,
) -> None:
    """One launch: AdamW update on fp32 state for all tensors; optionally also writes the bf16
    model copy (``model_params``) — semantic match for reference
    ``utils/adamw_fp32_optim_params.py:91-155``.

    ``hf_form=True`` (default) reproduces that file's update bit-for-bit in structure: ``denom = sqrt(v) + eps``,
    ``step = lr·sqrt(bc2)/bc1`` and the decoupled decay applied *after* the Adam step; ``hf_form=False`` is
    ``torch.optim.AdamW`` (decay first, eps after bias correction)."""
    if not params:
        return
    bc1 = 1.0 - beta1 ** step if correct_bias else 1.0
    bc2 = 1.0 - beta2 ** step if correct_bias else 1.0
    if _all_cuda(params) and _ext.use_cuda(*params):
        dev = params[0].device
        gs = grad_scale if grad_scale is not None else torch.ones((), dtype=torch.float32, device=dev)
        gs = gs.to(device=dev, dtype=torch.float32).reshape(())
        by = {}
        for i, (p, g) in enumerate(zip(params, grads)):
            mp = model_params[i] if model_params is not None else None
            key = (g.dtype, None if mp is None else mp.dtype)
            by.setdefault(key, []).append(i)
        for key, idxs in by.items():
            _ext.count_launch()
            _ext.ext().fused_adamw(
                [params[i] for i in idxs], [grads[i] for i in idxs], [exp_avgs[i] for i in idxs],
                [exp_avg_sqs[i] for i in idxs],
                [model_params[i] for i in idxs] if model_params is not None else [],
                float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), float(bc1), float(bc2), gs, bool(hf_form),
            )
        return
    for i, (p, g, m, v) in enumerate(zip(params, grads, exp_avgs, exp_avg_sqs)):
        g32 = g.float()
        if grad_scale is not None:
            g32 = g32 * grad_scale.to(g32.device)
        m.mul_(beta1).add_(g32, alpha=1.0 - beta1)
        v.mul_(beta2).addcmul_(g32, g32, value=1.0 - beta2)
        if hf_form:
            p.addcdiv_(m, v.sqrt().add_(eps), value=-lr * (bc2 ** 0.5) / bc1)
            if weight_decay > 0.0:
                p.mul_(1.0 - lr * weight_decay)
        else:
            p.mul_(1.0 - lr * weight_decay)
            p.addcdiv_(m / bc1, (v / bc2).sqrt_().add_(eps), value=-lr)
        if model_params is not None:
            model_params[i].copy_(p)
