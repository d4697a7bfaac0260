"""De-quantisation (reference ``quantization/dequantize.py:11-121``) incl. OCP microscaling (block 32, E8M0 scales)."""
from __future__ import annotations

import torch

from .microscaling.mx_torch import dequantize_mxfp4_packed, dequantize_mxfp8_packed, e8m0_to_float


def direct_cast_dequantize(tensor: torch.Tensor, upcast_dtype: torch.dtype) -> torch.Tensor:
    return tensor.to(upcast_dtype)


def scale_dequantize(tensor: torch.Tensor, scale: torch.Tensor, upcast_dtype: torch.dtype) -> torch.Tensor:
    return (tensor.to(torch.float32) * scale.to(torch.float32)).to(upcast_dtype)


def dequantize(tensor: torch.Tensor, scale: torch.Tensor, upcast_dtype: torch.dtype) -> torch.Tensor:
    return scale_dequantize(tensor, scale, upcast_dtype)


def mx_dequantize(packed: torch.Tensor, scale_e8m0: torch.Tensor, kind: str, upcast_dtype: torch.dtype) -> torch.Tensor:
    """``kind`` ∈ {"mxfp4", "mxfp8"}; blocks of 32 elements along the last dim share one E8M0 scale."""
    vals = dequantize_mxfp4_packed(packed) if kind == "mxfp4" else dequantize_mxfp8_packed(packed)
    s = e8m0_to_float(scale_e8m0).repeat_interleave(32, dim=-1)
    return (vals * s).to(upcast_dtype)
