"""Inference runtime (B200 equivalent of the reference's ``trace/`` compile/serve layer)."""
from .sharding import shard_state_dict_for_rank  # noqa: F401


def __getattr__(name):
    import importlib

    table = {
        "ModelBuilder": ".model_builder", "NxDModel": ".nxd_model", "BaseNxDModel": ".nxd_model",
        "shard_checkpoint": ".model_builder", "NxDParallelState": ".parallel_context",
        "parallel_model_trace": ".trace", "parallel_model_save": ".trace", "parallel_model_load": ".trace",
    }
    if name in table:
        return getattr(importlib.import_module(table[name], __name__), name)
    raise AttributeError(name)
