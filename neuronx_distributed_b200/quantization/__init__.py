from . import dequantize, microscaling, quantization_config, quantization_layers, quantization_utils, quantize  # noqa: F401
from .quantization_config import (ActivationQuantizationType, KVQuantizationConfig, QuantizationType, QuantizedDtype,  # noqa: F401
                                  ScaleDtype)
from .quantize import convert  # noqa: F401
