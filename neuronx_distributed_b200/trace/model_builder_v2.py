from ..inference.model_builder_v2 import ModelBuilder  # noqa: F401
