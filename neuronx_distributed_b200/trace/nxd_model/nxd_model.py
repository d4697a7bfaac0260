"""Runtime model (reference ``trace/nxd_model/nxd_model.py``).  ``TorchScriptNxDModel`` /
``convert_nxd_model_to_torchscript_model`` exist there to ship a Python-free artefact; a captured CUDA graph cannot be
serialised, so here the portable artefact is ``NxDModel.save`` (signatures + weights, re-captured by ``NxDModel.load``) and
the conversion returns the model itself."""
from ...inference.nxd_model import BaseNxDModel, NxDModel, StateInitializer  # noqa: F401
from ..model_builder import JITWrapper  # noqa: F401


class TorchScriptNxDModel(NxDModel):
    pass


def convert_nxd_model_to_torchscript_model(nxd_model: NxDModel, save_weights: bool = False) -> NxDModel:
    return nxd_model
