"""Weight / activation quantisation math (symmetric absmax; reference ``quantization/quantization_utils.py:12-250``).

Two families live here: the tensor-level primitives the quantised layers use (``quantize_per_tensor / per_channel /
blockwise``, ``quantize_activation_dynamic``) and the reference's checkpoint-side utilities (``quantize_pytorch_model_*``
to produce a quantised state dict from a float ``nn.Module``, ``convert_qint8_to_int8_state_dict`` and
``extract_q_scale*`` to read torch ``qint8`` packed checkpoints)."""
from __future__ import annotations

import copy
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn

from .observer import PerChannelAbsMaxObserver  # noqa: F401  (re-export, reference quantization_utils.py:9)
from .quantization_config import DtypeBound


_QMAX = {torch.int8: 127.0, torch.float8_e4m3fn: 448.0, torch.float8_e5m2: 57344.0}


def qmax(dtype: torch.dtype) -> float:
    return _QMAX[dtype]


def _cast(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    if dtype == torch.int8:
        return x.round().clamp(-127, 127).to(torch.int8)
    return x.clamp(-_QMAX[dtype], _QMAX[dtype]).to(dtype)


def quantize_per_tensor(w: torch.Tensor, dtype: torch.dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    scale = (w.abs().max().float() / _QMAX[dtype]).clamp(min=1e-12)
    return _cast(w.float() / scale, dtype), scale.reshape(1)


def quantize_per_channel(w: torch.Tensor, dtype: torch.dtype, axis: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """One scale per index of ``axis`` (scale keeps the weight's rank with size 1 on reduced dims)."""
    dims = [d for d in range(w.dim()) if d != axis % w.dim()]
    scale = (w.abs().amax(dim=dims, keepdim=True).float() / _QMAX[dtype]).clamp(min=1e-12)
    return _cast(w.float() / scale, dtype), scale


def quantize_blockwise(w: torch.Tensor, dtype: torch.dtype, block_axis: Sequence[int], block_size: Sequence[int]
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Independent scale per block of ``block_size[i]`` elements along ``block_axis[i]``."""
    shape = list(w.shape)
    view, red = [], []
    for d, n in enumerate(shape):
        if d in block_axis:
            b = block_size[list(block_axis).index(d)]
            assert n % b == 0, f"dim {d} ({n}) not divisible by block {b}"
            view += [n // b, b]
            red.append(len(view) - 1)
        else:
            view.append(n)
    wv = w.float().reshape(view)
    scale = (wv.abs().amax(dim=red, keepdim=True) / _QMAX[dtype]).clamp(min=1e-12)
    q = _cast(wv / scale, dtype).reshape(shape)
    return q, scale.squeeze(red) if red else scale


def dequantize_blockwise(q: torch.Tensor, scale: torch.Tensor, block_axis: Sequence[int], block_size: Sequence[int],
                         dtype: torch.dtype) -> torch.Tensor:
    s = scale
    for d in sorted(block_axis):
        s = s.repeat_interleave(block_size[list(block_axis).index(d)], dim=d)
    return (q.float() * s).to(dtype)


def quantize_activation_dynamic(x: torch.Tensor, dtype: torch.dtype = torch.float8_e4m3fn, clamp_bound: Optional[float] = None
                                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-row (last dim) dynamic quantisation: ``x ≈ q · scale[..., None]``."""
    xf = x.float()
    if clamp_bound is not None:
        xf = xf.clamp(-clamp_bound, clamp_bound)
    scale = (xf.abs().amax(dim=-1, keepdim=True) / _QMAX[dtype]).clamp(min=1e-12)
    return _cast(xf / scale, dtype), scale



# ---------------------------------------------------------------------------------------------------------------------
# reference-named utilities
# ---------------------------------------------------------------------------------------------------------------------
def _normalize_modules_to_not_convert_paths(float_model: nn.Module, modules_to_not_convert: List[str]) -> List[str]:
    """Every module path that contains one of the given fragments (checkpoint names and model names differ by prefixes)."""
    return [name for name, _ in float_model.named_modules() if any(frag in name for frag in modules_to_not_convert)]


def extract_q_scale_per_tensor(q_tensor: torch.Tensor) -> torch.Tensor:
    assert q_tensor.qscheme() == torch.per_tensor_affine
    return torch.tensor([q_tensor.q_scale()])


def extract_q_scale_per_channel(q_tensor: torch.Tensor) -> torch.Tensor:
    """Scales of a torch per-channel quantised tensor, shaped to broadcast against the weight (``[C, 1, …]``)."""
    assert q_tensor.qscheme() == torch.per_channel_affine
    axis = q_tensor.q_per_channel_axis()
    shape = [1] * q_tensor.dim()
    shape[axis] = q_tensor.shape[axis]
    return q_tensor.q_per_channel_scales().to(torch.float32).view(shape)


def extract_q_scale(q_tensor: torch.Tensor) -> torch.Tensor:
    if q_tensor.qscheme() == torch.per_tensor_affine:
        return extract_q_scale_per_tensor(q_tensor)
    if q_tensor.qscheme() == torch.per_channel_affine:
        return extract_q_scale_per_channel(q_tensor)
    raise ValueError(f"qscheme: {q_tensor.qscheme()} is not supported")


def quantize_static_quant_activations(input: torch.Tensor, input_scale: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """``clamp(x / input_scale)`` cast to ``dtype`` (rounded for int8) with a calibrated per-tensor scale."""
    q_max, q_min = DtypeBound.from_torch_dtype(dtype)
    x = (input.float() / input_scale.float()).clamp(q_min, q_max)
    return x.round().to(dtype) if dtype == torch.int8 else x.to(dtype)


def convert_qint8_to_int8_state_dict(state_dict: Dict[str, Any]) -> None:
    """In place: torch dynamic-quantised ``…_packed_params`` entries → plain ``weight`` (int8) / ``scale`` / ``bias``."""
    prefixes = {k.split("_packed_params.dtype")[0] for k in state_dict if "_packed_params.dtype" in k}
    for prefix in prefixes:
        packed = state_dict.pop(prefix + "_packed_params._packed_params")
        state_dict.pop(prefix + "_packed_params.dtype")
        state_dict.pop(prefix + "zero_point", None)
        state_dict[prefix + "weight"] = torch.int_repr(packed[0])
        state_dict[prefix + "scale"] = extract_q_scale(packed[0])
        bias = packed[1] if len(packed) == 2 else None
        state_dict[prefix + "bias"] = bias.data if isinstance(bias, nn.Parameter) else bias


def quantize_fp8_per_channel(tensor: torch.Tensor, dtype: torch.dtype, channel_axis: int, clamp_bound: float = float("inf")
                             ) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp8 weight + fp32 scale shaped ``[1, …, C, …, 1]`` (one scale per index of ``channel_axis``)."""
    assert dtype in (torch.float8_e4m3fn, torch.float8_e5m2)
    fmax, fmin = DtypeBound.from_torch_dtype(dtype)
    dims = tuple(d for d in range(tensor.dim()) if d != channel_axis % tensor.dim())
    t = tensor.to(torch.float32)
    amax = t.abs().amax(dim=dims, keepdim=True)
    if clamp_bound != float("inf"):
        amax = amax.clamp(max=clamp_bound)
        t = t.clamp(-clamp_bound, clamp_bound)
    scales = (amax / fmax).clamp(min=1e-5)
    return (t / scales).clamp(fmin, fmax).to(dtype), scales


def quantize_fp8_per_tensor(tensor: torch.Tensor, dtype: torch.dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    assert dtype in (torch.float8_e4m3fn, torch.float8_e5m2)
    fmax, fmin = DtypeBound.from_torch_dtype(dtype)
    scale = (tensor.float().abs().max() / fmax).clamp(min=1e-12)
    return (tensor.float() / scale).clamp(fmin, fmax).to(dtype), scale


def _int8_per_tensor(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    q, s = quantize_per_tensor(w, torch.int8)
    return q, s.reshape(1)


class QuantizedLinear(nn.Module):
    """``nn.Linear`` with int8 / fp8 weight and fp32 ``scale`` (state-dict keys ``weight``, ``scale``, ``bias``) — the
    checkpoint producer for the quantised parallel layers.  Unlike the reference's stand-in it also runs (weight-only
    de-quantised matmul), which makes CPU-side accuracy checks of a quantised checkpoint possible."""

    def __init__(self) -> None:
        super().__init__()
        self.register_parameter("bias", None)

    def set_weight_and_scale(self, weight: torch.Tensor, scale: torch.Tensor, bias: Optional[torch.Tensor] = None) -> None:
        self.weight = nn.Parameter(weight, requires_grad=False)
        self.scale = nn.Parameter(scale, requires_grad=False)
        self.bias = None if bias is None else nn.Parameter(bias.detach().clone(), requires_grad=False)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        y = torch.nn.functional.linear(input.float(), self.weight.float()) * self.scale.float().reshape(1, -1)
        y = y if self.bias is None else y + self.bias.float()
        return y.to(input.dtype)

    @classmethod
    def from_float(cls, mod: nn.Linear, dtype: torch.dtype = torch.int8, per_channel: bool = False) -> "QuantizedLinear":
        dtype = torch.int8 if dtype == torch.qint8 else dtype
        if per_channel:
            weight, scale = quantize_per_channel(mod.weight.data, dtype, 0)
        else:
            weight, scale = quantize_per_tensor(mod.weight.data, dtype)
        q = cls()
        q.set_weight_and_scale(weight, scale, mod.bias)
        return q


def _swap_linears(model: nn.Module, dtype: torch.dtype, per_channel: bool, skip: Sequence[str], prefix: str = "") -> None:
    for name, child in list(model.named_children()):
        full = f"{prefix}.{name}" if prefix else name
        if isinstance(child, nn.Linear) and full not in skip:
            setattr(model, name, QuantizedLinear.from_float(child, dtype, per_channel))
        else:
            _swap_linears(child, dtype, per_channel, skip, full)


def _check_cpu_quant_dtype(dtype) -> None:
    if dtype not in (torch.qint8, torch.int8, torch.float8_e4m3fn, torch.float8_e5m2):
        raise ValueError(f"dtype: {dtype} is not supported to quantize model on CPU")


def quantize_pytorch_model_per_tensor_symmetric(float_model: nn.Module, inplace: bool = False, dtype=torch.qint8) -> nn.Module:
    """Every ``nn.Linear`` → :class:`QuantizedLinear` with one scale per weight."""
    _check_cpu_quant_dtype(dtype)
    model = float_model if inplace else copy.deepcopy(float_model)
    _swap_linears(model, dtype, False, ())
    return model


def quantize_pytorch_model_per_channel_symmetric(float_model: nn.Module, inplace: bool = False, dtype=torch.qint8,
                                                 modules_to_not_convert: Optional[List[str]] = None) -> nn.Module:
    """Every ``nn.Linear`` → :class:`QuantizedLinear` with one scale per output feature (``scale [out, 1]``)."""
    _check_cpu_quant_dtype(dtype)
    skip = _normalize_modules_to_not_convert_paths(float_model, modules_to_not_convert) if modules_to_not_convert else []
    model = float_model if inplace else copy.deepcopy(float_model)
    _swap_linears(model, dtype, True, skip)
    return model


def quantize_per_tensor_symmetric(tensor: torch.Tensor) -> torch.Tensor:
    """torch ``qint8`` tensor (per-tensor affine, zero-point 0) — the format ``extract_q_scale`` reads."""
    scale = float((tensor.detach().float().abs().max() / 127.0).clamp(min=torch.finfo(torch.float32).eps))
    return torch.quantize_per_tensor(tensor.detach().float(), scale, 0, torch.qint8)


def quantize_per_channel_symmetric(tensor: torch.Tensor, channel_axis: int) -> torch.Tensor:
    obs = PerChannelAbsMaxObserver(ch_axis=channel_axis)
    obs(tensor)
    scales, zeros = obs.calculate_qparams()
    return torch.quantize_per_channel(tensor.detach().float(), scales.double(), zeros, channel_axis, torch.qint8)
