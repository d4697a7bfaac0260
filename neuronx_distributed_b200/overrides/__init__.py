from . import transformer_overrides  # noqa: F401
