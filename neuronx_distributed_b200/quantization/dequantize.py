"""De-quantisation helpers (reference ``quantization/dequantize.py:11-121``).

``scale_dequantize`` is the epilogue form (multiply a matmul output by a broadcastable scale);
``blockwise_scale_dequantize`` expands block scales — fp32 or E8M0, any subset of axes blocked — over a weight, and
understands x4-packed MXFP4 / MXFP8 storage.  On CUDA these run as a handful of elementwise kernels when a quantised
layer materialises its bf16 weight once per forward; the fp8 GEMM path (``ops.gemm_fp8``) never calls them."""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from .microscaling.mx_torch import dequantize_mxfp4_packed, dequantize_mxfp8_packed, e8m0_to_float
from .quantization_config import QuantizedDtype, ScaleDtype

DEFAULT_BLOCK_SIZE = 128
OCP_MX_BLOCK_SIZE = 32


def direct_cast_dequantize(tensor: torch.Tensor, upcast_dtype: torch.dtype) -> torch.Tensor:
    """Value-preserving upcast (int8 and fp8 values are exact in bf16)."""
    return tensor.to(upcast_dtype)


def scale_dequantize(tensor: torch.Tensor, scale: torch.Tensor, upcast_dtype: torch.dtype) -> torch.Tensor:
    """``tensor · scale`` in fp32, result in ``upcast_dtype``.  ``scale`` broadcasts from the right (a ``[1, N]`` scale
    against ``[B, S, N]``); a trailing singleton column (``[N, 1]`` weight-style scale) is accepted for a ``[..., N]``
    tensor as well."""
    s = scale.to(torch.float32)
    if s.dim() >= 2 and s.shape[-1] == 1 and s.shape[-2] == tensor.shape[-1] and tensor.shape[-1] != 1:
        s = s.transpose(-1, -2)
    return (tensor.to(torch.float32) * s).to(upcast_dtype)


def dequantize(tensor: torch.Tensor, scale: torch.Tensor, upcast_dtype: torch.dtype) -> torch.Tensor:
    return scale_dequantize(tensor, scale, upcast_dtype)


def get_broadcastable_shapes_for_blockwise_scale_dequantize(tensor_shape: Sequence[int], scale_shape: Sequence[int]
                                                            ) -> Tuple[Tuple[int, ...], Tuple[int, ...]]:
    """Views under which ``tensor.reshape(a) * scale.reshape(b)`` applies one scale per block: every blocked axis
    ``n`` with ``k`` scales becomes ``(k, n/k)`` on the tensor and ``(k, 1)`` on the scale; missing trailing scale axes
    count as one block."""
    t, s = list(tensor_shape), list(scale_shape)
    assert len(s) <= len(t), f"scale has more dims than the tensor: tensor {t}, scale {s}"
    s += [1] * (len(t) - len(s))
    tv, sv, blocked = [], [], False
    for n, k in zip(t, s):
        assert k <= n and n % k == 0, f"tensor dim {n} is not a multiple of scale dim {k}"
        if k == n:
            tv.append(n); sv.append(k)
        else:
            tv += [k, n // k]; sv += [k, 1]
            blocked = True
    assert blocked, "scale and tensor shapes are identical: no blocked dimension"
    return tuple(tv), tuple(sv)


def blockwise_scale_dequantize(tensor: torch.Tensor, scale: torch.Tensor, upcast_dtype: torch.dtype,
                               mx_swizzle: bool = False) -> torch.Tensor:
    if mx_swizzle:
        raise ValueError("tile-interleaved MX weights are consumed by the block-scaled GEMM only; de-swizzle first")
    if tensor.dtype == QuantizedDtype.F4E2M1FN_X4.value:
        tensor = dequantize_mxfp4_packed(tensor.view(torch.uint16) if tensor.dtype != torch.uint16 else tensor)
    elif tensor.dtype == torch.uint32:
        tensor = dequantize_mxfp8_packed(tensor)
    if scale.dtype == ScaleDtype.F8E8M0.value:
        scale = e8m0_to_float(scale)
    tv, sv = get_broadcastable_shapes_for_blockwise_scale_dequantize(tensor.shape, scale.shape)
    out = tensor.to(torch.float32).reshape(tv) * scale.to(torch.float32).reshape(sv)
    return out.reshape(tensor.shape).to(upcast_dtype)


def mx_dequantize(packed: torch.Tensor, scale_e8m0: torch.Tensor, kind: str, upcast_dtype: torch.dtype) -> torch.Tensor:
    """``kind`` ∈ {"mxfp4", "mxfp8"}; blocks of 32 elements along the last dim share one E8M0 scale."""
    vals = dequantize_mxfp4_packed(packed) if kind == "mxfp4" else dequantize_mxfp8_packed(packed)
    s = e8m0_to_float(scale_e8m0).repeat_interleave(OCP_MX_BLOCK_SIZE, dim=-1)
    return (vals * s).to(upcast_dtype)
