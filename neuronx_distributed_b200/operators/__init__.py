from .argmax import argmax  # noqa: F401
from .topk import topk  # noqa: F401
