from ..inference.parallel_context import NxDParallelState  # noqa: F401
