"""Capture named intermediate tensors from a model run — module inputs/outputs via forward hooks and values registered
manually from model code (reference ``utils/tensor_capture``).  Works in eager mode and under CUDA-graph capture."""
from .api import (disable_tensor_capture, enable_tensor_capture, get_available_modules, get_captured_tensors,  # noqa: F401
                  get_captured_tensors_dict, register_tensor)
from .model_modification import (find_available_modules, modify_hf_eager_model_for_tensor_capture,  # noqa: F401
                                 modify_model_for_tensor_capture, restore_model)
from .registry import CapturedModelInfo, TensorRegistry  # noqa: F401

CapturedTensorRegistry = TensorRegistry   # earlier name in this package
