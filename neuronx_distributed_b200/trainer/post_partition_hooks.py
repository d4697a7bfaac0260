"""Deferred callbacks executed after the model has been partitioned / materialised
(reference ``trainer/post_partition_hooks.py:5-35``)."""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Sequence, Tuple


class PostPartitionHooks:
    def __init__(self) -> None:
        self._hooks: List[Tuple[Callable, Sequence[Any], Dict[str, Any]]] = []

    @property
    def hooks(self) -> List[Dict[str, Any]]:
        """The pending hooks as the reference stores them: ``{"function": partial, "name": str}``."""
        import functools

        return [{"function": functools.partial(fn, *a, **k), "name": getattr(fn, "__name__", type(fn).__name__)}
                for fn, a, k in self._hooks]

    def register_post_partition_hook(self, callable_function: Callable[..., Any] = None, func_args: Sequence[Any] = (),
                                     func_kwargs: Dict[str, Any] | None = None, **aliases):
        """Argument names of reference post_partition_hooks.py:11; ``fn`` / ``args`` / ``kwargs`` (this package's earlier
        names) are accepted as keywords."""
        fn = callable_function if callable_function is not None else aliases.pop("fn", None)
        func_args = aliases.pop("args", func_args)
        func_kwargs = aliases.pop("kwargs", func_kwargs)
        if aliases:
            raise TypeError(f"unexpected arguments {sorted(aliases)}")
        if not callable(fn):
            raise ValueError("callable_function must be a callable object")
        self._hooks.append((fn, tuple(func_args), dict(func_kwargs or {})))

    def execute_all_hooks(self, model=None) -> List[Any]:
        """Run and clear the hooks; returns their results.  A hook registered without arguments receives the model; a hook
        named ``filter_to_local_parameter_group`` receives it as ``model=`` next to its own arguments (reference :22-34)."""
        outputs = []
        for fn, args, kwargs in self._hooks:
            if getattr(fn, "__name__", "") == "filter_to_local_parameter_group":
                assert model is not None, "When executing filter_to_local_parameter_group hook, model object cannot be None"
                outputs.append(fn(*args, model=model, **kwargs))
            else:
                outputs.append(fn(*args, **kwargs) if args or kwargs else fn(model))
        self._hooks.clear()
        return outputs


hooks = PostPartitionHooks()
register_post_partition_hook = hooks.register_post_partition_hook
execute_all_hooks = hooks.execute_all_hooks
