from .model import NxDPPModel  # noqa: F401
