from . import (dequantize, microscaling, observer, quantization_config, quantization_layers, quantization_mappings,  # noqa: F401
               quantization_utils, quantize)
from .quantization_config import (ActivationQuantizationType, KVQuantizationConfig, QuantizationType, QuantizedDtype,  # noqa: F401
                                  ScaleDtype)
from .quantize import convert  # noqa: F401
