"""Linear layers on microscaling (OCP MX) weights.

Weights are the K-contiguous byte stream the checkpoints ship (``quantization/microscaling``): MXFP4 = e2m1 codes two per
byte (low nibble first; ``uint16`` x4 words are a view of the same bytes), MXFP8 = one e4m3 / e5m2 byte per element
(``uint32`` x4 words likewise); one E8M0 scale per 32 elements of a row.

* ``rows <= 8`` on CUDA (token generation): ``csrc/gemv_mx.cu`` — the codes are decoded in registers, nothing is expanded in
  memory, so a decode step reads 4.25 / 8.25 bits per weight instead of 16;
* otherwise: de-quantise to the activation dtype and run the dense GEMM (tcgen05 bf16 kernel on CUDA) — numerically the
  oracle ``experimental…mx_torch.mx_matmul``."""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _ext

_FMT = {"mxfp4": 0, "mxfp8": 1, "mxfp8_e4m3": 1, "mxfp8_e5m2": 2}


def kind_of(weight: torch.Tensor, fp8_dtype: torch.dtype = torch.float8_e4m3fn) -> str:
    if weight.dtype in (torch.uint16, torch.float16, torch.int16):
        return "mxfp4"
    if weight.dtype in (torch.uint32, torch.int32):
        return "mxfp8_e5m2" if fp8_dtype == torch.float8_e5m2 else "mxfp8"
    raise ValueError(f"not an x4-packed MX weight: {weight.dtype}")


def dequantize(weight: torch.Tensor, scale: torch.Tensor, kind: str, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """``[N, K/4]`` x4 words (or the raw byte stream) + ``[N, K/32]`` E8M0 → ``[N, K]``."""
    b = weight.contiguous().view(torch.uint8)
    N = scale.shape[0]
    if _FMT[kind] == 0:
        lut = torch.tensor([0, .5, 1, 1.5, 2, 3, 4, 6, -0., -.5, -1, -1.5, -2, -3, -4, -6], dtype=torch.float32, device=b.device)
        vals = torch.stack((lut[(b & 0xF).long()], lut[(b >> 4).long()]), dim=-1).reshape(N, -1)
    else:
        vals = b.view(torch.float8_e4m3fn if _FMT[kind] == 1 else torch.float8_e5m2).float().reshape(N, -1)
    return torch.ldexp(vals.reshape(N, -1, 32), (scale.to(torch.int32) - 127).unsqueeze(-1)).reshape(N, -1).to(dtype)


def gemv_eligible(x2d: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor) -> bool:
    if os.environ.get("NXD_GEMV_MX", "0") != "1":                 # opt-in until the kernel has run on hardware
        return False
    return (x2d.is_cuda and x2d.dtype == torch.bfloat16 and x2d.dim() == 2 and 1 <= x2d.shape[0] <= 8 and x2d.shape[1] % 32 == 0
            and x2d.is_contiguous() and weight.is_contiguous() and scale.is_contiguous() and scale.dtype == torch.uint8
            and scale.dim() == 2 and _ext.use_cuda(x2d, weight, scale) and hasattr(_ext.ext(), "gemv_mx"))


def linear_mx(x: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor, kind: Optional[str] = None,
              residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``x [..., K] @ dequant(weight)[N, K]ᵀ (+ residual)`` → ``[..., N]`` in ``x.dtype``."""
    kind = kind or kind_of(weight)
    x2 = x.reshape(-1, x.shape[-1])
    if gemv_eligible(x2.contiguous(), weight, scale):
        _ext.count_launch()
        r2 = None if residual is None else residual.reshape(-1, residual.shape[-1]).contiguous()
        y = _ext.ext().gemv_mx(x2.contiguous(), weight, scale, _FMT[kind], r2)
        return y.view(*x.shape[:-1], y.shape[-1])
    w = dequantize(weight, scale, kind, x.dtype if x.dtype in (torch.bfloat16, torch.float16) else torch.float32)
    y = torch.nn.functional.linear(x.to(w.dtype), w)
    if residual is not None:
        y = y + residual.to(y.dtype)
    return y.to(x.dtype)
