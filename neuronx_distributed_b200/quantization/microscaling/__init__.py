from . import mx_torch, transform_weights  # noqa: F401
from .mx_torch import (  # noqa: F401
    dequantize_mx_tensor, dequantize_mxfp4_packed, dequantize_mxfp8_packed, e8m0_to_float, float_to_e8m0, matmul_mx,
    mx_matmul, quantize_mx, quantize_mxfp8,
)
