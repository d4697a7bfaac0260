"""PyTorch reference for OCP microscaling formats (role of reference
``experimental/quantization/microscaling/mx_torch.py:65-253``): MXFP8 (e4m3) and MXFP4 (e2m1) with one shared
power-of-two E8M0 scale per block of 32 along the last dim; packing 4 values per word
(mxfp4 → uint16, mxfp8 → uint32) as in ``quantization_config.QuantizedDtype``."""
from __future__ import annotations

from typing import Tuple

import torch

BLOCK = 32
_FP4_VALUES = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])
_FP4_MAX_EXP = 2      # 6 = 1.5·2^2
_FP8_MAX_EXP = 8      # 448 = 1.75·2^8


def float_to_e8m0(scale: torch.Tensor) -> torch.Tensor:
    return (torch.log2(scale.float()).round() + 127).clamp(0, 254).to(torch.uint8)


def e8m0_to_float(e: torch.Tensor) -> torch.Tensor:
    return torch.pow(2.0, e.float() - 127.0)


def _block_scale(x: torch.Tensor, elem_max_exp: int) -> torch.Tensor:
    xb = x.float().reshape(*x.shape[:-1], -1, BLOCK)
    amax = xb.abs().amax(-1).clamp(min=2.0 ** -100)
    return torch.pow(2.0, torch.floor(torch.log2(amax)) - elem_max_exp)     # [..., nblk]


def _to_fp4_code(v: torch.Tensor) -> torch.Tensor:
    """Nearest e2m1 code (sign bit 3, magnitude index 0-7)."""
    mag = v.abs().clamp(max=6.0)
    idx = (mag.unsqueeze(-1) - _FP4_VALUES.to(v.device)).abs().argmin(-1)
    return (idx | ((v < 0).long() << 3)).to(torch.uint8)


def _from_fp4_code(c: torch.Tensor) -> torch.Tensor:
    mag = _FP4_VALUES.to(c.device)[(c & 7).long()]
    return torch.where((c & 8) != 0, -mag, mag)


def quantize_mx(x: torch.Tensor, kind: str = "mxfp8") -> Tuple[torch.Tensor, torch.Tensor]:
    """Return ``(packed, scale_e8m0)``; last dim must be a multiple of 32."""
    assert x.shape[-1] % BLOCK == 0
    scale = _block_scale(x, _FP4_MAX_EXP if kind == "mxfp4" else _FP8_MAX_EXP)
    xs = x.float() / scale.repeat_interleave(BLOCK, dim=-1)
    if kind == "mxfp4":
        codes = _to_fp4_code(xs).reshape(*x.shape[:-1], -1, 4).to(torch.int32)
        packed = (codes[..., 0] | (codes[..., 1] << 4) | (codes[..., 2] << 8) | (codes[..., 3] << 12)).to(torch.uint16)
    else:
        b = xs.clamp(-448, 448).to(torch.float8_e4m3fn).view(torch.uint8).reshape(*x.shape[:-1], -1, 4).to(torch.int64)
        packed = (b[..., 0] | (b[..., 1] << 8) | (b[..., 2] << 16) | (b[..., 3] << 24)).to(torch.uint32)
    return packed, float_to_e8m0(scale)


def dequantize_mxfp4_packed(packed: torch.Tensor) -> torch.Tensor:
    p = packed.to(torch.int32)
    codes = torch.stack([(p >> s) & 0xF for s in (0, 4, 8, 12)], dim=-1).reshape(*packed.shape[:-1], -1)
    return _from_fp4_code(codes.to(torch.uint8))


def dequantize_mxfp8_packed(packed: torch.Tensor) -> torch.Tensor:
    p = packed.to(torch.int64)
    b = torch.stack([(p >> s) & 0xFF for s in (0, 8, 16, 24)], dim=-1).reshape(*packed.shape[:-1], -1).to(torch.uint8)
    return b.view(torch.float8_e4m3fn).float()


def mx_matmul(a: torch.Tensor, b_packed: torch.Tensor, b_scale: torch.Tensor, kind: str = "mxfp8",
              out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """``a [M,K] @ dequant(b)[N,K]ᵀ`` — numerics oracle for the block-scaled tensor-core path."""
    vals = dequantize_mxfp4_packed(b_packed) if kind == "mxfp4" else dequantize_mxfp8_packed(b_packed)
    w = vals * e8m0_to_float(b_scale).repeat_interleave(BLOCK, dim=-1)
    return (a.float() @ w.t()).to(out_dtype)
