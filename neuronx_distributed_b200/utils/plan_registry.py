"""Registry of framework functions that are ONE node of a launch plan (``inference/launch_plan.py``).  Kept free of imports
so that ``parallel_layers`` / ``ops`` can decorate their collectives without pulling in the inference package."""
from __future__ import annotations

import functools
from typing import Any, Callable, Dict, Optional, Tuple

PY_OPS: Dict[str, Tuple[Callable, bool]] = {}
_STATE: Dict[str, Optional[Any]] = {"recorder": None}


def active_recorder():
    return _STATE["recorder"]


def set_recorder(rec) -> None:
    _STATE["recorder"] = rec


def plan_op(name: str, pure: bool = False):
    """Mark a framework function as one plan node.  Its arguments must be tensors, scalars, dtypes, (nested) lists of
    those, process groups registered in ``parallel_state`` or framework enums; whatever it does inside (NCCL, symmetric
    memory, epochs) is re-done by calling the same function at replay.  ``pure``: no side effect besides its outputs."""

    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            rec = _STATE["recorder"]
            if rec is None or rec.paused:
                return fn(*args, **kwargs)
            return rec.record_call("py", name, fn, args, kwargs)

        PY_OPS[name] = (fn, pure)
        wrapper.__plan_op__ = name
        return wrapper

    return deco


def recording() -> bool:
    rec = _STATE["recorder"]
    return rec is not None and not rec.paused
