"""Deferred callbacks executed after the model has been partitioned / materialised
(reference ``trainer/post_partition_hooks.py:5-35``)."""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Sequence, Tuple


class PostPartitionHooks:
    def __init__(self) -> None:
        self._hooks: List[Tuple[Callable, Sequence[Any], Dict[str, Any]]] = []

    def register_post_partition_hook(self, fn: Callable, args: Sequence[Any] = (), kwargs: Dict[str, Any] | None = None):
        self._hooks.append((fn, tuple(args), dict(kwargs or {})))

    def execute_all_hooks(self, model) -> None:
        for fn, args, kwargs in self._hooks:
            fn(*args, **kwargs) if args or kwargs else fn(model)
        self._hooks.clear()


hooks = PostPartitionHooks()
register_post_partition_hook = hooks.register_post_partition_hook
execute_all_hooks = hooks.execute_all_hooks
