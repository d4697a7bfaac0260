"""Experimental pieces (reference ``src/neuronx_distributed/experimental``): numerics oracles, not production paths."""
from . import quantization  # noqa: F401
