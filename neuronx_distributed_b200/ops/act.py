"""SwiGLU: ``silu(gate) * up`` on a fused ``[…, 2*I]`` gate|up tensor (kernels:
``csrc/elementwise.cu`` swiglu_fwd / swiglu_bwd).  Reference call site:
``examples/training/llama/modeling_llama_nxd.py:208-219``; MoE GLU ``modules/moe/experts.py:219-235``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _ext


class _SwiGLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate_up):
        ctx.save_for_backward(gate_up)
        shape = gate_up.shape
        gu2 = gate_up.reshape(-1, shape[-1])
        if _ext.use_cuda(gu2) and gu2.dtype in (torch.bfloat16, torch.float32, torch.float16) and shape[-1] % 16 == 0:
            _ext.count_launch()
            ctx.cuda = True
            out = _ext.ext().swiglu_fwd(gu2.contiguous())
        else:
            ctx.cuda = False
            g, u = gu2.float().chunk(2, dim=-1)
            out = (F.silu(g) * u).to(gate_up.dtype)
        return out.view(*shape[:-1], shape[-1] // 2)

    @staticmethod
    def backward(ctx, go):
        (gate_up,) = ctx.saved_tensors
        shape = gate_up.shape
        gu2 = gate_up.reshape(-1, shape[-1])
        go2 = go.reshape(-1, go.shape[-1])
        if ctx.cuda:
            _ext.count_launch()
            return _ext.ext().swiglu_bwd(go2.contiguous(), gu2.contiguous()).view(shape)
        g, u = gu2.float().chunk(2, dim=-1)
        gof = go2.float()
        sig = torch.sigmoid(g)
        silu = g * sig
        dg = gof * u * (sig + silu * (1 - sig))
        du = gof * silu
        return torch.cat([dg, du], dim=-1).to(gate_up.dtype).view(shape)


def swiglu(gate_up: torch.Tensor) -> torch.Tensor:
    """``gate_up[..., :I]`` is gate, ``gate_up[..., I:]`` is up."""
    return _SwiGLU.apply(gate_up)
