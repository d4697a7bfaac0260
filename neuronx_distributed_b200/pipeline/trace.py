"""FX tracing front-end of the pipeline partitioner (reference ``pipeline/trace.py:31-219``).

``trace_model`` itself lives in :mod:`.partition` (tracing and cutting are one step here); this module carries the tracer
classes and helpers user code may customise or import: leaf selection by class-name fragment or qualified name
(``NxDTracer``), the torch / HF tracer wrappers, ``get_concrete_args``, ``get_tracer_class`` and ``patch_obj_method`` (make
chosen bound methods opaque to the tracer)."""
from __future__ import annotations

import contextlib
import inspect
from collections import defaultdict
from typing import Any, Dict, List, Optional

import torch.fx as fx
from torch import nn

from ..utils.model_utils import is_hf_pretrained_model, is_hf_transformers_available
from .partition import trace_model as _trace_with_leaf_classes

try:
    from torch.fx._symbolic_trace import _create_wrapped_func
except ImportError:  # pragma: no cover
    _create_wrapped_func = None


class NxDTracer(fx.Tracer):
    """A module is a leaf when a ``leaf_modules`` entry is a fragment of its class name or equals its qualified name."""

    leaf_modules: List[str] = []

    def is_leaf_module(self, m: nn.Module, module_qualified_name: str) -> bool:
        leaves = getattr(self, "leaf_modules", ()) or ()
        if any(t in type(m).__name__ for t in leaves) or module_qualified_name in leaves:
            return True
        return super().is_leaf_module(m, module_qualified_name)


class TorchTracerWrapper(NxDTracer):
    """``torch.fx.Tracer`` with ``param_shapes_constant`` (shape accesses on parameters fold to constants, which keeps
    ``view(-1, weight.shape[0])``-style code traceable)."""

    def __init__(self, **config) -> None:
        super().__init__(autowrap_modules=tuple(config.get("autowrap_modules", ())),
                         autowrap_functions=tuple(config.get("autowrap_functions", ())), param_shapes_constant=True)
        self.leaf_modules = list(config.get("leaf_modules", []))
        self.name = "pytorch"


if is_hf_transformers_available():  # pragma: no branch
    try:
        from transformers.utils.fx import HFTracer

        class HFTracerWrapper(NxDTracer, HFTracer):
            def __init__(self, **config) -> None:
                HFTracer.__init__(self, autowrap_modules=tuple(config.get("autowrap_modules", ())),
                                  autowrap_functions=tuple(config.get("autowrap_functions", ())))
                self.leaf_modules = list(config.get("leaf_modules", ()))
                self.name = "HF"
    except Exception:  # noqa: BLE001  (transformers builds without the fx utilities)
        HFTracerWrapper = None  # type: ignore[assignment,misc]
else:
    HFTracerWrapper = None  # type: ignore[assignment,misc]


def get_concrete_args(model: nn.Module, input_names: Optional[List[str]] = None, args: Optional[List[Any]] = None,
                      kwargs: Optional[Dict[Any, Any]] = None) -> Dict[str, Any]:
    """``{parameter: default}`` for every ``forward`` parameter that is NOT a traced input — what ``fx`` must treat as a
    constant.  ``input_names`` may be given directly or inferred from example ``args`` / ``kwargs``."""
    sig = inspect.signature(model.forward)
    if input_names is None and (args is not None or kwargs is not None):
        input_names = list(sig.parameters.keys())[: len(args or [])] + list((kwargs or {}).keys())
    assert input_names is not None, "input_names, args or kwargs must be provided"
    unknown = [n for n in input_names if n not in sig.parameters]
    if unknown:
        raise ValueError(f"The model does not have input(s) named: {', '.join(unknown)}, expected a subset of the following: "
                         f"{', '.join(sig.parameters.keys())}")
    return {p.name: p.default for p in sig.parameters.values() if p.name not in input_names}


def get_tracer_class(model: nn.Module, tracer_cls=None):
    if tracer_cls is None:
        return HFTracerWrapper if (is_hf_pretrained_model(model) and HFTracerWrapper is not None) else TorchTracerWrapper
    if isinstance(tracer_cls, str):
        if tracer_cls == "torch":
            return TorchTracerWrapper
        if tracer_cls == "hf" and HFTracerWrapper is not None:
            return HFTracerWrapper
        raise ValueError(f"Unsupported tracer_cls {tracer_cls}")
    return tracer_cls


@contextlib.contextmanager
def patch_obj_method(autowrap_obj_methods: Optional[Dict[Any, List[str]]]):
    """``{obj: ["method", …]}`` → while inside the context those bound methods are wrapped so that ``fx`` records a single
    ``call_function`` node instead of tracing into them; restored on exit."""
    enabled = _create_wrapped_func is not None and bool(autowrap_obj_methods)
    saved: Dict[Any, Dict[str, Any]] = defaultdict(dict)
    if enabled:
        for obj, methods in autowrap_obj_methods.items():
            assert isinstance(methods, list), f"Expect autowrap_obj_methods has list as value but getting {type(methods)}"
            for name in methods:
                if not hasattr(obj, name):
                    raise ValueError(f"Inside autowrap_obj_methods obj type {type(obj)} does not have method {name}")
                own = name in getattr(obj, "__dict__", {})
                saved[obj][name] = (getattr(obj, name), own)
                setattr(obj, name, _create_wrapped_func(saved[obj][name][0]))
    try:
        yield
    finally:
        for obj, methods in saved.items():
            for name, (original, own) in methods.items():
                if own:
                    setattr(obj, name, original)
                else:                                   # was a class attribute: drop the shadowing instance attribute
                    try:
                        delattr(obj, name)
                    except AttributeError:
                        setattr(obj, name, original)


def trace_model(model: nn.Module, args: Optional[List[Any]] = None, kwargs: Optional[Dict[Any, Any]] = None,
                input_names: Optional[List[str]] = None, tracer_cls: Any = None, leaf_modules: Optional[List[Any]] = None,
                autowrap_functions: Optional[List[Any]] = None, autowrap_modules: Optional[List[Any]] = None,
                autowrap_obj_methods: Optional[Dict[Any, List[str]]] = None, leaf_module_cls=None) -> fx.GraphModule:
    """FX-trace ``model`` for pipeline partitioning with the reference's calling convention (``pipeline/trace.py:153-203``):
    the traced inputs come from ``input_names`` or from example ``args`` / ``kwargs`` (only their NAMES matter — nothing is
    executed); ``leaf_modules`` are classes or class names kept as single call nodes (the parallel layers and norms always
    are); ``autowrap_obj_methods`` = ``{obj: [method names]}`` recorded as opaque calls.  This package's earlier positional
    form ``trace_model(model, input_names, leaf_classes)`` is still accepted."""
    if isinstance(args, (list, tuple)) and args and all(isinstance(a, str) for a in args) and input_names is None:
        input_names, args = list(args), None                                  # (model, input_names, leaf classes, …)
        if isinstance(kwargs, (list, tuple)):
            leaf_modules, kwargs = list(kwargs), None
    if input_names is None and (args is not None or kwargs is not None):
        sig = list(inspect.signature(model.forward).parameters)
        input_names = sig[: len(args or [])] + [k for k in (kwargs or {}) if kwargs[k] is not None]
    wanted = list(leaf_modules or []) + list(leaf_module_cls or [])
    by_name = {type(m).__name__: type(m) for m in model.modules()}
    leaf_classes = []
    for item in wanted:
        cls = by_name.get(item) if isinstance(item, str) else item
        if cls is None:
            raise ValueError(f"leaf module {item!r} does not name the class of any submodule")
        leaf_classes.append(cls)
    if tracer_cls is not None and not (isinstance(tracer_cls, type) and issubclass(tracer_cls, fx.Tracer)):
        tracer_cls = None if tracer_cls in ("torch", "hf") else tracer_cls    # both names select the built-in leaf-aware tracer
    with patch_obj_method(autowrap_obj_methods):
        return _trace_with_leaf_classes(model, input_names, leaf_classes, tuple(autowrap_functions or ()),
                                        tuple(autowrap_modules or ()), None)

