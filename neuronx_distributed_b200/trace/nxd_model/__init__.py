from .base_nxd_model import BaseNxDModel, StateInitializer  # noqa: F401
from .nxd_model import (BucketProgram, NxDModel, TorchScriptNxDModel,  # noqa: F401
                        convert_nxd_model_to_torchscript_model)
