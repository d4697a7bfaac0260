from . import microscaling  # noqa: F401
