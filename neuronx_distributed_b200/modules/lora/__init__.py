from .config import LoraConfig  # noqa: F401
from .layer import LoraConv2d, LoraEmbedding, LoraLayer, LoraLinear  # noqa: F401
from .model import LoraModel, get_lora_model  # noqa: F401
from .tp_layer import LoraGQAQKVParallelLinear, LoraParallelLinear  # noqa: F401
from .serving import LoraServingConfig, LoraServingModel, MultiLoraLinear, wrap_model_with_lora  # noqa: F401
