from ...inference.nxd_model import BaseNxDModel, StateInitializer  # noqa: F401
