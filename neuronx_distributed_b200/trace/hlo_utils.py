"""Reference ``trace/hlo_utils.py`` rewrites HLO protobufs so that every bucket consumes the weight layout the compiler
chose for the priority bucket (weight-layout optimisation, WLO) and builds a separate "layout transformer" program that
re-lays-out the checkpoint at load time.

None of that exists on B200 by construction:

* buckets are CUDA graphs captured over the SAME parameter tensors — there is one copy of every weight and therefore one
  layout;
* the tcgen05 GEMMs read weights through TMA tensor maps with 128-byte swizzle applied by the copy engine on the way into
  shared memory, so the *stored* layout is plain row-major ``[N, K]`` for every tile shape and bucket;
* the one offline re-layout that does exist — interleaving MX scale factors into 128×4 tiles — is a data transform
  (``experimental.quantization.microscaling.swizzle.swizzle_scale_factors``), not a program transform.

The functions below keep the names that user code may import and implement the identity behaviour of a world without
WLO; functions that only make sense on HLO protobufs raise with an explanation."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch


def _no_hlo(name: str):
    def f(*args, **kwargs):
        raise NotImplementedError(f"{name} operates on HLO protobufs; programs are captured CUDA graphs here "
                                  "(see neuronx_distributed_b200/trace/hlo_utils.py)")
    f.__name__ = name
    return f


read_hlo = _no_hlo("read_hlo")
write_hlo = _no_hlo("write_hlo")
add_weight_idx_attr_to_hlo = _no_hlo("add_weight_idx_attr_to_hlo")
update_computation_id_and_name = _no_hlo("update_computation_id_and_name")
append_layout_computation_to_hlo = _no_hlo("append_layout_computation_to_hlo")
extract_weight_layout_transform_hlo = _no_hlo("extract_weight_layout_transform_hlo")
prepare_metaneff_for_wlt_hlo = _no_hlo("prepare_metaneff_for_wlt_hlo")
read_metaneff = _no_hlo("read_metaneff")
traceback_instruction_to_parameter = _no_hlo("traceback_instruction_to_parameter")


def mark_weights_for_wlo(trace_artifacts: Any = None, weights_to_skip_layout_optimization: Optional[set] = None, **_) -> None:
    """No weight is re-laid-out: nothing to mark."""


def apply_layout_transformation(trace_artifacts: Any = None, priority_model_trace_artifacts: Any = None, wlo_artifacts: Any = None,
                                key: Optional[str] = None, **_) -> None:
    """All buckets already share the priority bucket's (only) layout."""


def get_layout_transform_map(*args, **kwargs) -> Dict[str, Any]:
    return {}


get_wlt_map = get_layout_transform_map


def get_wlt(*args, **kwargs):
    return None


def update_weight(weights: Dict[str, torch.Tensor], *args, **kwargs) -> Dict[str, torch.Tensor]:
    return weights


def transform_weight_layout_on_cpu(weights: Dict[str, torch.Tensor], *args, **kwargs) -> Dict[str, torch.Tensor]:
    return weights


def transform_weight_layout_on_device_and_save_to_disk(*args, **kwargs) -> None:
    return None


def get_input_order(trace_artifacts: Any) -> List[str]:
    """Names of a bucket's inputs in call order."""
    return [a.param_name for a in trace_artifacts.provided_args]


def convert_inputs_to_optimal_shape(inputs, *args, **kwargs):
    return inputs


def cleanup_after_layout_transformation(*args, **kwargs) -> None:
    return None


def get_compiler_package_dir() -> str:
    import os

    return os.environ.get("CUDA_HOME", "/usr/local/cuda")


def get_executable_full_qualified_path(name: str = "nvcc") -> Optional[str]:
    import shutil

    return shutil.which(name)


def is_nki_kernel_called(*args, **kwargs) -> bool:
    return False


def prepare_parameter_usage_map(*args, **kwargs) -> Dict[str, Any]:
    return {}


def get_nki_kernel_weight_names(*args, **kwargs) -> List[str]:
    return []
