"""Quantization configuration (reference ``quantization/quantization_config.py:65-256``)."""
from __future__ import annotations

import enum
from dataclasses import dataclass
from typing import Any, TypedDict

import torch


class QuantizationType(enum.Enum):
    PER_TENSOR_SYMMETRIC = "per_tensor_symmetric"
    PER_CHANNEL_SYMMETRIC = "per_channel_symmetric"
    PER_KEY_SYMMETRIC = "per_key_symmetric"
    BLOCKWISE_SYMMETRIC = "blockwise_symmetric"
    EXPERT_WISE_PER_CHANNEL_SYMMETRIC = "expert_wise_per_channel_symmetric"


class ActivationQuantizationType(enum.Enum):
    NONE = "none"
    DYNAMIC = "dynamic"      # per-row absmax computed on the fly
    STATIC = "static"        # fixed calibrated scale


class QuantizedDtype(enum.Enum):
    INT8 = torch.int8
    F8E4M3 = torch.float8_e4m3fn
    F8E5M2 = torch.float8_e5m2
    F4E2M1FN_X4 = "mxfp4_x4"      # 4 e2m1 values packed in a uint16
    F8E4M3FN_X4 = "mxfp8_x4"      # 4 e4m3 values packed in a uint32

    def storage_dtype(self) -> torch.dtype:
        return {QuantizedDtype.F4E2M1FN_X4: torch.uint16, QuantizedDtype.F8E4M3FN_X4: torch.uint32}.get(self, self.value)

    def get_packed_count(self) -> int:
        return 4 if self in (QuantizedDtype.F4E2M1FN_X4, QuantizedDtype.F8E4M3FN_X4) else 1

    @staticmethod
    def get_dtype(name) -> "QuantizedDtype":
        if isinstance(name, QuantizedDtype):
            return name
        table = {"int8": QuantizedDtype.INT8, "f8e4m3": QuantizedDtype.F8E4M3, "f8e5m2": QuantizedDtype.F8E5M2,
                 "mxfp4": QuantizedDtype.F4E2M1FN_X4, "mxfp8": QuantizedDtype.F8E4M3FN_X4}
        return table[str(name).lower()]


class ScaleDtype(enum.Enum):
    F32 = torch.float32
    E8M0 = "e8m0"                 # power-of-two scale stored as a uint8 exponent (OCP microscaling)


class BASE_QCONFIG_DICT_TYPE(TypedDict, total=False):
    quantization_type: QuantizationType
    quantized_dtype: QuantizedDtype
    quantization_per_channel_axis: int
    block_axis: Any
    block_size: Any
    activation_quantization_type: ActivationQuantizationType
    clamp_bound: float
    scale_dtype: ScaleDtype


def get_default_per_tensor_custom_qconfig_dict() -> BASE_QCONFIG_DICT_TYPE:
    return {"quantization_type": QuantizationType.PER_TENSOR_SYMMETRIC, "quantized_dtype": QuantizedDtype.INT8,
            "activation_quantization_type": ActivationQuantizationType.NONE, "scale_dtype": ScaleDtype.F32}


def get_default_per_channel_custom_qconfig_dict() -> BASE_QCONFIG_DICT_TYPE:
    return {"quantization_type": QuantizationType.PER_CHANNEL_SYMMETRIC, "quantized_dtype": QuantizedDtype.INT8,
            "quantization_per_channel_axis": 0, "activation_quantization_type": ActivationQuantizationType.NONE,
            "scale_dtype": ScaleDtype.F32}


def get_default_blockwise_custom_qconfig_dict() -> BASE_QCONFIG_DICT_TYPE:
    return {"quantization_type": QuantizationType.BLOCKWISE_SYMMETRIC, "quantized_dtype": QuantizedDtype.F8E4M3,
            "block_axis": [1], "block_size": [128], "activation_quantization_type": ActivationQuantizationType.NONE,
            "scale_dtype": ScaleDtype.F32}


def get_default_expert_wise_per_channel_custom_qconfig_dict() -> BASE_QCONFIG_DICT_TYPE:
    return {"quantization_type": QuantizationType.EXPERT_WISE_PER_CHANNEL_SYMMETRIC, "quantized_dtype": QuantizedDtype.INT8,
            "quantization_per_channel_axis": 2, "activation_quantization_type": ActivationQuantizationType.NONE,
            "scale_dtype": ScaleDtype.F32}


@dataclass
class KVQuantizationConfig:
    quant_dtype: torch.dtype = torch.float8_e4m3fn
    dequant_dtype: torch.dtype = torch.bfloat16
    scale: float = 1.0
    per_key: bool = False
