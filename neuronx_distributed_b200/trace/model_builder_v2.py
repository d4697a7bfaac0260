"""``ModelBuilder`` v2 — ``ModelBuilder(model).trace(args, kwargs, tag)…compile() → NxDModel``
(reference ``trace/model_builder_v2.py:33-319``).  One builder = one module; every ``trace`` call adds a bucket."""
from __future__ import annotations

import time
from typing import Any, Callable, Dict, Optional, Set, Union

import torch
from torch import nn

from ..utils.logger import get_logger
from .functions import compile as _compile
from .functions import compile_layout_transformer, compile_wlo, trace as _trace
from .model_builder_utils import ModelBuilderConstants, TraceArtifacts, generate_key
from .nxd_model import NxDModel

logger = get_logger()


class ModelBuilder:
    def __init__(self, model: Union[Callable, nn.Module], weights_to_skip_layout_optimization: Optional[Set] = None):
        self.model = model
        self.weights_to_skip_layout_optimization = weights_to_skip_layout_optimization
        self.trace_artifacts_collection: Dict[str, TraceArtifacts] = {}
        self.world_size = torch.distributed.get_world_size() if torch.distributed.is_initialized() else \
            ModelBuilderConstants.DEFAULT_WORLD_SIZE

    def trace(self, args=None, kwargs: Optional[Dict[str, torch.Tensor]] = None, tag: Optional[str] = None,
              spmd: bool = True) -> "ModelBuilder":
        t0 = time.time()
        ta = _trace(self.model, args, kwargs, spmd=spmd, preserve_parameters=True,
                    weights_to_skip_layout_optimization=self.weights_to_skip_layout_optimization)
        tag = generate_key(ta, tag)
        self.trace_artifacts_collection[tag] = ta
        logger.info("Finished tracing %s in %.3f seconds", tag, time.time() - t0)
        return self

    def compile(self, priority_model_key: Optional[str] = None, compiler_workdir=None,
                compiler_args: Optional[Union[str, Dict[str, str]]] = None, max_workers: Optional[int] = None) -> NxDModel:
        """Capture every traced bucket.  ``priority_model_key`` names the bucket that is compiled first and with
        weight-layout optimisation (``compile_wlo``: its weight-only launches are hoisted into a layout transformer that runs
        once per weight load); a bucket whose launches cannot be recorded (``PlanError``) is captured as is, with a
        warning.  Captures run sequentially — CUDA graph capture is a per-stream, per-process affair — so
        ``max_workers`` is accepted and ignored."""
        if not self.trace_artifacts_collection:
            raise ValueError("No traces available for compilation. Call trace() first.")
        if priority_model_key and priority_model_key not in self.trace_artifacts_collection:
            raise ValueError(f"Invalid priority_model_key: {priority_model_key}")
        if isinstance(compiler_args, dict):
            missing = set(self.trace_artifacts_collection) - set(compiler_args)
            if missing:
                raise ValueError(f"Missing compiler args for buckets: {missing}")
        elif isinstance(compiler_args, str):
            compiler_args = {k: compiler_args for k in self.trace_artifacts_collection}
        results: Dict[str, Any] = {}
        order = list(self.trace_artifacts_collection)
        if priority_model_key:
            order.remove(priority_model_key)
            order.insert(0, priority_model_key)
        try:
            from ..inference.launch_plan import PlanError

            for key in order:
                ta, flags = self.trace_artifacts_collection[key], (compiler_args[key] if compiler_args else None)
                if key == priority_model_key:
                    try:
                        results[key] = compile_wlo(ta, None, compiler_workdir, flags, key)
                        continue
                    except PlanError as e:
                        logger.warning("bucket %s: launch plan not recordable (%s); compiled without the layout pass", key, e)
                        ta._plan = None
                results[key] = _compile(ta, None, compiler_workdir, flags, key)
            if priority_model_key:
                results[ModelBuilderConstants.LAYOUT_TRANSFORMER_KEY] = compile_layout_transformer(results[priority_model_key])
        except Exception as e:  # noqa: BLE001
            raise RuntimeError("Compilation process failed") from e
        return self._build_nxd_model(results)

    def _build_nxd_model(self, compilation_results: Dict[str, Any]) -> NxDModel:
        nxd = NxDModel(world_size=self.world_size,
                       layout_transformer=compilation_results.get(ModelBuilderConstants.LAYOUT_TRANSFORMER_KEY))
        for key, ta in self.trace_artifacts_collection.items():
            nxd.add(key=key, trace_artifacts=ta, compilation_artifacts=compilation_results[key])
        return nxd
