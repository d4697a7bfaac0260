"""``neuronx_distributed.trace`` import paths (reference ``trace/__init__.py``).  The implementation is the
:mod:`neuronx_distributed_b200.inference` package — "tracing/compiling" a bucket means capturing it into a CUDA graph —
and every reference module name resolves to the corresponding part of it."""
from ..inference.model_builder import ModelBuilder  # noqa: F401
from ..inference.sharding import shard_state_dict_for_rank  # noqa: F401
from ..inference.trace import parallel_model_load, parallel_model_save, parallel_model_trace  # noqa: F401
from .spmd import NxDModel, SPMDBucketModelScript  # noqa: F401


def __getattr__(name):
    import importlib

    inf = importlib.import_module("..inference", __name__)
    return getattr(inf, name)
