"""RMSNorm forward/backward (kernels: ``csrc/elementwise.cu`` rmsnorm_fwd / rmsnorm_bwd).

fp32 statistics regardless of I/O dtype (reference ``modules/rms_norm.py:10-33``).  The backward
kernel emits dX and per-CTA partial dW rows that a tiny second kernel reduces.
"""
from __future__ import annotations

import torch

from . import _ext


def _rms_ref(x, w, eps):
    xf = x.float()
    rstd = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return (xf * rstd * w.float()).to(x.dtype), rstd


class _RMSNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, eps):
        ctx.eps = eps
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        if _ext.use_cuda(x2, weight) and x2.dtype in (torch.bfloat16, torch.float32, torch.float16):
            _ext.count_launch()
            y, rstd = _ext.ext().rmsnorm_fwd(x2.contiguous(), weight.contiguous(), float(eps))
            ctx.cuda = True
        else:
            y, rstd = _rms_ref(x2, weight, eps)
            ctx.cuda = False
        ctx.save_for_backward(x2, weight, rstd)
        return y.view(shape)

    @staticmethod
    def backward(ctx, gy):
        x2, weight, rstd = ctx.saved_tensors
        g2 = gy.reshape(-1, gy.shape[-1])
        if ctx.cuda:
            _ext.count_launch(2)
            gx, gw = _ext.ext().rmsnorm_bwd(g2.contiguous(), x2.contiguous(), weight.contiguous(), rstd)
            return gx.view(gy.shape), gw.to(weight.dtype), None
        xf, gf, wf = x2.float(), g2.float(), weight.float()
        xhat = xf * rstd
        gw = (gf * xhat).sum(0)
        gxh = gf * wf
        gx = rstd * (gxh - xhat * (gxh * xhat).mean(-1, keepdim=True))
        return gx.to(x2.dtype).view(gy.shape), gw.to(weight.dtype), None


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    return _RMSNorm.apply(x, weight, eps)
