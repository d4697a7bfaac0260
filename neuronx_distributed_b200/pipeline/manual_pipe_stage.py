"""Manual pipeline partition: the user supplies an ordered list of layers instead of relying on FX
tracing (reference ``pipeline/manual_pipe_stage.py:14-295``)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

from torch import nn

from .partition import create_partitions


WEIGHT_SHARING_ATTR_NAME = "weight_sharing"          # parameter attribute carrying the tie-group index


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
    happens (default: even split, remainder to later stages)."""

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

    def mark_weight_sharing(self, names: Sequence[str]) -> None:
        """Declare parameters (qualified names inside ``all_layers``) that are tied across stages.  The parameters are also
        tagged with ``WEIGHT_SHARING_ATTR_NAME`` = index of their tie group (how the reference marks them)."""
        group = len(self._tied)
        self._tied.append(list(names))
        params = dict(self.all_layers.named_parameters(remove_duplicate=False))
        for n in names:
            if n in params:
                setattr(params[n], WEIGHT_SHARING_ATTR_NAME, group)

    def forward(self, *args, **kwargs):
        x = None
        for i, layer in enumerate(self.all_layers):
            x = layer(*args, **kwargs) if i == 0 else (layer(*x) if isinstance(x, tuple) else layer(x))
        return x
