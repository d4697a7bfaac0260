from .base_nxd_model import BaseNxDModel, StateInitializer  # noqa: F401
from .nxd_model import NxDModel  # noqa: F401
