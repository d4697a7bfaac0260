"""v1 builder (reference ``trace/model_builder.py``): ``ModelBuilder(router, tp_degree, checkpoint_loader).add(key, …).trace()``."""
from ..inference.functions import (append_default_compiler_flags, compile, compile_layout_transformer, compile_wlo,  # noqa: F401,A004
                                   trace)
from ..inference.model_builder import BaseModelInstance, ModelBuilder, shard_checkpoint  # noqa: F401
from ..inference.nxd_model import NxDModel  # noqa: F401
from .spmd import NxDModelExecutor, SPMDBucketModelScript, StateInitializer  # noqa: F401


class ModelContainer:
    """One registered key of the v1 builder: the model instance, its example inputs (one tuple per bucket) and the per-key
    options (reference :66-83)."""

    def __init__(self, model_instance, example_inputs, compiler_args=None, bucket_config=None, priority_model_idx=None):
        self.model_instance, self.example_inputs = model_instance, example_inputs
        self.compiler_args, self.bucket_config, self.priority_model_idx = compiler_args, bucket_config, priority_model_idx
        self.traced = None


class JITWrapper:
    """Reference :85-95 wraps a Python callable for TorchScript; callables need no wrapping to be captured here."""

    def __init__(self, func, is_torchscript: bool = False):
        self.func = func

    def __call__(self, inputs):
        return self.func(inputs)


def get_hash_module(module_description) -> str:
    import hashlib

    return hashlib.sha256(repr(module_description).encode()).hexdigest()


def init_process_wrapper(*args, **kwargs):
    """The reference spawns one worker process per rank for tracing; ranks are ``torchrun`` processes here."""
    raise RuntimeError("launch one process per GPU with torchrun; the builder does not spawn workers")
