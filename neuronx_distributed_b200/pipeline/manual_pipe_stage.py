"""Manual pipeline partition: the user supplies an ordered list of layers instead of relying on FX
tracing (reference ``pipeline/manual_pipe_stage.py:14-295``)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
from torch import nn

from .partition import create_partitions


WEIGHT_SHARING_ATTR_NAME = "weight_sharing"          # layer attribute: {sharing group name: path of the shared weight}


class _InstanceOrStatic:
    """``PipelineStageModule.mark_weight_sharing(...)`` (the reference's static form, called before the stage module exists)
    and ``stage.mark_weight_sharing(...)`` resolve to the same function; the instance is passed when there is one."""

    def __init__(self, fn):
        self.fn = fn
        self.__doc__ = fn.__doc__

    def __get__(self, obj, cls):
        import functools

        return functools.partial(self.fn, obj)


class _LocalStage(nn.Module):
    def __init__(self, layers: Sequence[nn.Module], names: Sequence[str]):
        super().__init__()
        self.layers = nn.ModuleDict({n.replace(".", "_"): l for n, l in zip(names, layers)})
        self._order = [n.replace(".", "_") for n in names]

    def forward(self, *args, **kwargs):
        x = None
        for i, n in enumerate(self._order):
            layer = self.layers[n]
            if i == 0:
                x = layer(*args, **kwargs)
            else:
                x = layer(*x) if isinstance(x, tuple) else layer(x)
        return x


class PipelineStageModule(nn.Module):
    """Holds the full ordered layer list; ``build_stage`` returns the sub-sequence of one stage.

    ``partition_fn(num_layers, num_stages) -> List[int]`` returns the layer index after which each cut
    happens (default: even split, remainder to later stages).

    Weights shared by layers on different stages (tied embeddings) are found three ways: the SAME ``Parameter`` object used by
    two layers, layers marked with :meth:`mark_weight_sharing` before construction (reference :209-262 — the weights may be
    separate tensors of equal shape), or qualified names declared on the instance.  The pipeline engine keeps each group in
    sync: broadcast from the first holder at start, gradient all-reduce over the holders every step."""

    def __init__(self, layers: Sequence[nn.Module], num_stages: int = 1, stage_index: int = 0,
                 partition_fn: Optional[Callable[[int, int], List[int]]] = None, layer_names: Optional[Sequence[str]] = None):
        super().__init__()
        self.all_layers = nn.ModuleList(layers)
        self.layer_names = list(layer_names) if layer_names else [f"layer_{i}" for i in range(len(layers))]
        self.partition_fn = partition_fn or create_partitions
        self.num_stages, self.stage_index = num_stages, stage_index
        self._tied: List[List[str]] = []

    def stage_layer_range(self, num_stages: int, stage: int):
        cuts = self.partition_fn(len(self.all_layers), num_stages) if num_stages > 1 else []
        bounds = [0] + [c + 1 for c in cuts] + [len(self.all_layers)]
        return bounds[stage], bounds[stage + 1]

    def build_stage(self, num_stages: int, stage: int) -> nn.Module:
        lo, hi = self.stage_layer_range(num_stages, stage)
        return _LocalStage(list(self.all_layers[lo:hi]), self.layer_names[lo:hi])

    @_InstanceOrStatic
    def mark_weight_sharing(self, layer_and_weight_path, sharing_group_name: Optional[str] = "default") -> None:
        """``[(layer, "path.to.weight"), …]``: record on every layer that this weight belongs to ``sharing_group_name`` (all
        members must have one shape; a name can be used once per layer).  On an instance, a list of qualified parameter names
        inside ``all_layers`` (``"0.weight"``) is accepted as well."""
        import operator

        items = list(layer_and_weight_path)
        if items and all(isinstance(i, str) for i in items):
            assert self is not None, "qualified names need a PipelineStageModule instance"
            self._tied.append(items)
            return
        shape = None
        for layer, path in items:
            marks = getattr(layer, WEIGHT_SHARING_ATTR_NAME, None)
            if not isinstance(marks, dict):
                marks = {}
            if sharing_group_name in marks:
                raise RuntimeError(f'weight sharing group named "{sharing_group_name}" already exists on {type(layer).__name__}')
            w = operator.attrgetter(path)(layer)
            assert isinstance(w, torch.Tensor), f"Expected to get a torch.Tensor, but got {type(w)}"
            shape = w.shape if shape is None else shape
            assert w.shape == shape, "All shared weights must have the same shape"
            marks[sharing_group_name] = path
            setattr(layer, WEIGHT_SHARING_ATTR_NAME, marks)

    def weight_sharing_groups(self) -> List[List[str]]:
        """Groups of qualified parameter names (``"<layer index>.<path>"``) that must stay equal."""
        groups: List[List[str]] = []
        seen = set()

        def add(names):
            key = frozenset(names)
            if len(key) > 1 and key not in seen:
                seen.add(key)
                groups.append(sorted(key, key=lambda n: (int(n.split(".", 1)[0]), n)))

        marked = {}
        for i, layer in enumerate(self.all_layers):
            marks = getattr(layer, WEIGHT_SHARING_ATTR_NAME, None)
            if isinstance(marks, dict):
                for g, path in marks.items():
                    marked.setdefault(g, []).append(f"{i}.{path}")
        for names in marked.values():
            add(names)
        for names in self._tied:
            add(names)
        by_id = {}
        for n, p in self.all_layers.named_parameters(remove_duplicate=False):
            by_id.setdefault(id(p), []).append(n)
        for names in by_id.values():
            add(names)
        return groups

    def shared_across_stages(self, num_stages: int) -> List[List[tuple]]:
        """The sharing groups as ``[(stage, parameter name inside that stage's module), …]`` (what the engine registers)."""
        stage_of = {}
        for s in range(num_stages):
            lo, hi = self.stage_layer_range(num_stages, s)
            for i in range(lo, hi):
                stage_of[i] = s
        out = []
        for names in self.weight_sharing_groups():
            group = []
            for n in names:
                idx, path = n.split(".", 1)
                group.append((stage_of[int(idx)], f"layers.{self.layer_names[int(idx)].replace('.', '_')}.{path}"))
            out.append(group)
        return out

    def forward(self, *args, **kwargs):
        x = None
        for i, layer in enumerate(self.all_layers):
            x = layer(*args, **kwargs) if i == 0 else (layer(*x) if isinstance(x, tuple) else layer(x))
        return x
