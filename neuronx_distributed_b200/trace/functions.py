"""Fundamental units of the inference builder (reference ``trace/functions.py:48-908``): ``trace`` → ``compile`` →
``NxDModel.add``, plus ``shard_checkpoint``.

``trace`` validates the call (no ``*args/**kwargs`` in the signature, only ``None`` defaults, every required parameter
supplied, tensors only), binds the example inputs to parameter names and runs the bucket once in eager mode to record the
output structure.  ``compile`` warms the bucket up on a side stream and captures it into a CUDA graph with persistent
input / output buffers (eager on CPU).  ``compile_wlo`` splits the priority bucket's launch plan
(``launch_plan.py``) into a layout transformer — the launches that depend only on frozen weights — and the per-call plan;
``compile_layout_transformer`` packages the transformer for ``NxDModel``."""
from __future__ import annotations

import inspect
import os
from typing import Any, Callable, Dict, List, Optional, Set, Tuple, Union

import torch
from torch import nn

from ..utils.logger import get_logger
from .model_builder_utils import (CompilationArtifacts, LayoutTransformerArtifacts, ModelParamInfo, ProvidedArgInfo,
                                  TraceArtifacts, WLOArtifacts)

logger = get_logger()


def append_default_compiler_flags(compiler_args: Optional[str] = "") -> str:
    """There is no ahead-of-time compiler; recognised "flags" configure graph capture: ``--no-cuda-graph``,
    ``--warmup=N``.  Unknown flags (neuronx-cc options of ported scripts) are kept and ignored."""
    args = (compiler_args or "").strip()
    if "--warmup" not in args:
        args = (args + " --warmup=2").strip()
    return args


def _flag(args: Optional[str], name: str, default=None):
    for tok in (args or "").split():
        if tok == name:
            return True
        if tok.startswith(name + "="):
            return tok.split("=", 1)[1]
    return default


def _validate_model(model: Union[Callable, nn.Module]) -> inspect.Signature:
    if model is None:
        raise ValueError("Model cannot be None")
    sig = inspect.signature(model.forward if isinstance(model, nn.Module) else model)
    if not sig.parameters:
        raise ValueError("Model must have at least one parameter")
    if any(p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD) for p in sig.parameters.values()):
        raise NotImplementedError("Methods with *args or **kwargs are not supported. Please explicitly specify all parameters.")
    return sig


def _validate_model_params(params) -> List[ModelParamInfo]:
    out = []
    for name, p in params.items():
        if name == "self":
            continue
        positional = p.default is inspect.Parameter.empty
        if not positional and p.default is not None:
            raise ValueError(f"Parameter '{name}' has a non-None default value: {p.default}. Only None is allowed as a "
                             "default value for parameters.")
        out.append(ModelParamInfo(name, positional))
    return out


def _validate_args(args, model_params: List[ModelParamInfo]) -> Tuple[torch.Tensor, ...]:
    if args is None:
        return ()
    if isinstance(args, torch.Tensor):
        args = (args,)
    if not isinstance(args, tuple) or not all(isinstance(t, torch.Tensor) for t in args):
        raise ValueError("args must be either None, a single tensor, or a tuple of tensors")
    if len(args) > len(model_params):
        raise ValueError(f"Too many positional arguments. Model accepts {len(model_params)} but received {len(args)}")
    return args


def _validate_kwargs(kwargs, model_params: List[ModelParamInfo], provided: List[ProvidedArgInfo]):
    if kwargs is None:
        return None
    if not isinstance(kwargs, dict):
        raise ValueError("kwargs must be a dictionary")
    names = {p.param_name for p in model_params}
    bad = set(kwargs) - names
    if bad:
        raise ValueError(f"Found unexpected keys in kwargs: {bad}. Valid keys are: {sorted(names)}")
    dup = set(kwargs) & {a.param_name for a in provided}
    if dup:
        raise ValueError(f"Parameters {dup} were already provided as positional arguments and cannot be overridden by "
                         "keyword arguments")
    for k, v in kwargs.items():
        if not isinstance(v, torch.Tensor):
            raise ValueError(f"Value for key '{k}' must be a tensor")
    return kwargs


def _process_example_inputs(model, args, kwargs) -> Tuple[List[ProvidedArgInfo], List[ModelParamInfo]]:
    """Bind example tensors to signature parameters: positional ones first (signature order), then keyword ones in
    *signature* order — the order every later call is flattened to."""
    model_params = _validate_model_params(_validate_model(model).parameters)
    by_name = {p.param_name: p for p in model_params}
    provided = [ProvidedArgInfo(model_params[i].param_name, model_params[i].is_positional, t)
                for i, t in enumerate(_validate_args(args, model_params))]
    kw = _validate_kwargs(kwargs, model_params, provided)
    if kw:
        provided += [ProvidedArgInfo(p.param_name, by_name[p.param_name].is_positional, kw[p.param_name])
                     for p in model_params if p.param_name in kw]
    have = {a.param_name for a in provided}
    missing = [p.param_name for p in model_params if p.is_positional and p.param_name not in have]
    if missing:
        raise ValueError(f"Missing required parameters: {missing}. These must be provided either as positional arguments "
                         "or keyword arguments.")
    if not provided:
        raise ValueError("At least one input tensor must be provided via args or kwargs")
    return provided, model_params


def _spec(obj):
    if isinstance(obj, torch.Tensor):
        return (tuple(obj.shape), str(obj.dtype))
    if isinstance(obj, (list, tuple)):
        return type(obj)(_spec(o) for o in obj)
    if isinstance(obj, dict):
        return {k: _spec(v) for k, v in obj.items()}
    return type(obj).__name__


def trace(model: Union[Callable, nn.Module], args=None, kwargs: Optional[Dict[str, torch.Tensor]] = None, spmd: bool = True,
          preserve_parameters: bool = True, weights_to_skip_layout_optimization: Optional[Set] = None) -> TraceArtifacts:
    if not spmd:
        raise NotImplementedError("MPMD tracing is not currently supported")
    provided, model_params = _process_example_inputs(model, args, kwargs)
    weights, states = {}, []
    if isinstance(model, nn.Module):
        weights = {n: i for i, (n, _) in enumerate(model.named_parameters())}
        before = {n: b._version for n, b in model.named_buffers()}
    try:
        with torch.no_grad():
            out = model(**{a.param_name: a.tensor for a in provided})
    except Exception as e:  # noqa: BLE001
        raise RuntimeError(f"Tracing run failed: {e}") from e
    if isinstance(model, nn.Module):
        states = [n for n, b in model.named_buffers() if b._version != before.get(n)]       # mutated in place = state
    skip = set(weights_to_skip_layout_optimization or ())
    return TraceArtifacts(model=model, provided_args=provided, model_params=model_params, output_spec=_spec(out),
                          weight_name_to_idx=weights, weight_names_to_skip={w for w in weights if w in skip}, state_names=states)


class _Program:
    """One bucket: static input buffers + (optionally) a captured CUDA graph."""

    def __init__(self, ta: TraceArtifacts, use_graph: bool, warmup: int, runner: Optional[Callable] = None):
        self.ta, self.runner = ta, runner
        self.names = [a.param_name for a in ta.provided_args]
        self.static_in = [a.tensor.clone() for a in ta.provided_args]
        self.graph, self.static_out = None, None
        cuda = torch.cuda.is_available() and all(t.is_cuda for t in self.static_in)
        if use_graph and cuda:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(warmup):
                    self._run(self.static_in)
            torch.cuda.current_stream().wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(self.graph):
                self.static_out = self._run(self.static_in)

    def _run(self, tensors):
        if self.runner is not None:                              # a LaunchPlan: positional, in the traced order
            return self.runner(*tensors)
        return self.ta.model(**dict(zip(self.names, tensors)))

    def __call__(self, *tensors: torch.Tensor):
        if self.graph is None:
            with torch.no_grad():
                return self._run(tensors)
        for dst, src in zip(self.static_in, tensors):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out


def compile(trace_artifacts: TraceArtifacts, metaneff: Any = None, compiler_workdir: Optional[Union[str, os.PathLike]] = None,  # noqa: A001
            compiler_args: Optional[str] = None, key: Optional[str] = None) -> CompilationArtifacts:
    """Capture one traced bucket.  (Second positional parameter mirrors the reference's ``(hlo, metaneff, …)`` order and
    is ignored.)"""
    if isinstance(metaneff, (str, os.PathLike)) and compiler_workdir is None:
        compiler_workdir, metaneff = metaneff, None
    flags = append_default_compiler_flags(compiler_args)
    use_graph = not _flag(flags, "--no-cuda-graph", False)
    from ..inference.launch_plan import LaunchPlan

    plan = None
    if isinstance(trace_artifacts.model, LaunchPlan) or _flag(flags, "--plan", False):
        plan = trace_artifacts.record_plan()                     # the program interprets the recorded launches
    prog = _Program(trace_artifacts, use_graph, int(_flag(flags, "--warmup", 2)), runner=plan)
    if compiler_workdir is not None:
        os.makedirs(compiler_workdir, exist_ok=True)
        with open(os.path.join(compiler_workdir, f"{key or 'model'}.program.txt"), "w") as f:
            f.write(repr(trace_artifacts.describe()) + f"\ncaptured_cuda_graph={prog.graph is not None}\nflags={flags}\n")
    return CompilationArtifacts(program=prog, key=key or "", compiler_workdir=None if compiler_workdir is None else str(compiler_workdir),
                                compiler_args=flags, captured=prog.graph is not None, plan=plan)


def compile_wlo(trace_artifacts: TraceArtifacts, metaneff: Any = None, compiler_workdir=None, compiler_args: Optional[str] = None,
                key: Optional[str] = None) -> WLOArtifacts:
    """Compile the priority bucket with weight-layout optimisation (reference ``functions.py:compile_wlo``): record the
    bucket's launch plan, hoist every launch that depends only on frozen weights into a layout-transformer plan (run once
    per weight load), and capture the remaining per-call plan.  Weights in ``weights_to_skip_layout_optimization`` and
    state buffers are left alone."""
    if isinstance(metaneff, (str, os.PathLike)) and compiler_workdir is None:
        compiler_workdir, metaneff = metaneff, None
    flags = append_default_compiler_flags(compiler_args)
    use_graph = not _flag(flags, "--no-cuda-graph", False)
    plan = trace_artifacts.record_plan()
    transformer, main, tmap = plan.hoist_weight_only(skip=trace_artifacts.weight_names_to_skip)
    main.apply_transformer(transformer)
    prog = _Program(trace_artifacts, use_graph, int(_flag(flags, "--warmup", 2)), runner=main)
    if compiler_workdir is not None:
        os.makedirs(compiler_workdir, exist_ok=True)
        main.save(os.path.join(compiler_workdir, f"{key or 'model'}.plan.json"))
        transformer.save(os.path.join(compiler_workdir, f"{key or 'model'}.layout_transformer.plan.json"))
    logger.info("WLO %s: %d launches hoisted into the layout transformer, %d per call", key, len(transformer.nodes), len(main.nodes))
    return WLOArtifacts(program=prog, key=key or "", compiler_workdir=None if compiler_workdir is None else str(compiler_workdir),
                        compiler_args=flags, captured=prog.graph is not None, plan=main, transformer=transformer,
                        layout_transform_map=tmap)


def compile_layout_transformer(wlo_artifacts: Optional[WLOArtifacts] = None,
                               priority_model_weight_name_to_idx: Optional[Dict[str, int]] = None, compiler_workdir=None,
                               **_unused) -> LayoutTransformerArtifacts:
    """Package the transformer extracted by :func:`compile_wlo` (reference ``functions.py:compile_layout_transformer``)."""
    lt = LayoutTransformerArtifacts()
    if wlo_artifacts is not None and getattr(wlo_artifacts, "transformer", None) is not None:
        lt.transformers[wlo_artifacts.key] = (wlo_artifacts.transformer, wlo_artifacts.plan)
    return lt


def shard_checkpoint(checkpoint: Dict[str, torch.Tensor], model: nn.Module, start_rank: Optional[int] = None,
                     end_rank: Optional[int] = None, load_on_device: bool = False, serialize_path: Optional[str] = None,
                     tp_degree: Optional[int] = None) -> List[Dict[str, torch.Tensor]]:
    """Reference argument order (``functions.py:808-908``): ``(checkpoint, model, start_rank, end_rank, load_on_device,
    serialize_path)``.  ``load_on_device`` places the shards on this process's GPU."""
    from .model_builder import shard_checkpoint as _impl

    shards = _impl(checkpoint, model, tp_degree, 0 if start_rank is None else start_rank, end_rank, load_on_device, serialize_path)
    if load_on_device and torch.cuda.is_available():
        shards = [{k: v.cuda(non_blocking=True) for k, v in sd.items()} for sd in shards]
    return shards
