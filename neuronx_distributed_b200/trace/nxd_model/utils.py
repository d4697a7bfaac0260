"""Helpers of the runtime model (reference ``trace/nxd_model/utils.py:12-157``)."""
from __future__ import annotations

from typing import Any, Dict, List, Sequence, Tuple

import torch

_DTYPES = [torch.float32, torch.float16, torch.bfloat16, torch.float64, torch.int8, torch.uint8, torch.int16, torch.int32,
           torch.int64, torch.bool, torch.float8_e4m3fn, torch.float8_e5m2, torch.uint16, torch.uint32]


TORCH_DTYPES = [getattr(torch, a) for a in dir(torch) if isinstance(getattr(torch, a), torch.dtype)]   # every dtype of this build


def get_dtype_enum(dtype: torch.dtype) -> int:
    """Stable integer code of a dtype (for metadata that must not pickle torch objects)."""
    return _DTYPES.index(dtype)


def get_dtype_from_enum(dtype_enum: int) -> torch.dtype:
    if not 0 <= dtype_enum < len(_DTYPES):
        raise ValueError(f"unknown dtype code {dtype_enum}")
    return _DTYPES[dtype_enum]


def retrieve_artifact_from_model(model, key: str, artifact: str):
    """``artifact`` ∈ {"hlo", "metaneff", "neff"} of bucket ``key``."""
    nxd_model = model      # reference parameter names in the signature
    return {"hlo": nxd_model.get_hlo, "metaneff": nxd_model.get_metaneff, "neff": nxd_model.get_neff}[artifact](key)


def generate_route_key_from_provided_args(provided_args: Sequence[Any]) -> str:
    """Routing key of a traced bucket: parameter names + shapes + dtypes of its example inputs."""
    from .nxd_model import NxDModel

    return NxDModel._route_key([a.param_name for a in provided_args], [a.tensor for a in provided_args])


def _ordered(model_params: List[Tuple[str, bool]], inputs: Dict[str, Any], num_pos_args: int):
    names = [p[0] for p in model_params]
    picked = [n for n in names[num_pos_args:] if n in inputs]
    return [inputs[n] for n in picked], names[:num_pos_args] + picked


def ts_convert_dict_to_ordered_list_type_tensor(model_params, inputs: Dict[str, torch.Tensor], num_pos_args: int):
    return _ordered(model_params, inputs, num_pos_args)


def ts_convert_dict_to_ordered_list_type_list_tensor(model_params, inputs: Dict[str, List[torch.Tensor]], num_pos_args: int):
    return _ordered(model_params, inputs, num_pos_args)
