from .mx_torch import (  # noqa: F401
    dequantize_mxfp4_packed, dequantize_mxfp8_packed, e8m0_to_float, float_to_e8m0, mx_matmul, quantize_mx,
)
