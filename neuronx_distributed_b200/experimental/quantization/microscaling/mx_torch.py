"""Torch oracle of MX quantisation and MX matmul (reference ``experimental/quantization/microscaling/mx_torch.py``).
The implementation lives with the production (de)quantisers in ``quantization.microscaling.mx_torch``."""
from ....quantization.microscaling.mx_torch import (  # noqa: F401
    VALID_MX_TYPES, VALID_QMX_INPUT_TYPE, VALID_QMX_OUTPUT_TYPE, dequantize_mx_tensor, matmul_mx, matmul_mx_single_tile, quantize_mxfp8,
)
