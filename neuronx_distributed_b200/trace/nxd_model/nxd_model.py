"""Runtime model (reference ``trace/nxd_model/nxd_model.py``).  ``TorchScriptNxDModel`` /
``convert_nxd_model_to_torchscript_model`` ship a model-code-free artefact there (TorchScript around NEFFs); a captured
CUDA graph cannot be serialised, so here the artefact is one *launch plan* per bucket (``inference/launch_plan.py``) that is
re-captured into a CUDA graph at load time."""
from ...inference.nxd_model import (BaseNxDModel, NxDModel, StateInitializer, TorchScriptNxDModel,  # noqa: F401
                                    convert_nxd_model_to_torchscript_model)
from ..model_builder import JITWrapper  # noqa: F401
