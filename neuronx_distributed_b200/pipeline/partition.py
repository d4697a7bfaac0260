"""Model → pipeline stages.

Role parity with reference ``pipeline/trace.py:31-219`` + ``pipeline/partition.py:18-303``: trace
with ``torch.fx`` keeping transformer layers and TP layers as leaves, cut after the requested layer
calls, split into ``submod_i`` stage modules, analyse stage IO (values live across each boundary,
including *pass-along* values that skip stages, are forwarded hop by hop), and detect parameters
shared by several stages (tied embeddings).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Set, Tuple

import torch.fx as fx
from torch import nn
from torch.fx.passes.split_module import split_module


def create_partitions(num_layers, num_stages):
    """Even split of the transformer layers into stages, remainder to the *later* stages (reference partition.py:268-303).
    Two calling conventions: ``create_partitions(num_layers, num_stages)`` → the layer INDICES after which to cut;
    the reference's ``create_partitions(pipeline_parallel_size, model_layer_names)`` → the layer NAMES after which to cut
    (what ``pipeline_cuts`` takes)."""
    names = None
    if not isinstance(num_stages, int):                                  # (pipeline_parallel_size, model_layer_names)
        names = list(num_stages)
        num_layers, num_stages = len(names), int(num_layers)
    if num_stages > num_layers:
        raise ValueError(f"cannot split {num_layers} layers into {num_stages} stages")
    base, rem = divmod(num_layers, num_stages)
    sizes = [base + (1 if s >= num_stages - rem else 0) for s in range(num_stages)]
    cuts, acc = [], 0
    for s in sizes[:-1]:
        acc += s
        cuts.append(acc - 1)
    return cuts if names is None else [names[c] for c in cuts]


def stage_to_pipeline_parallel_rank(stage: int, pp_size: int) -> int:
    return stage % pp_size


class _LeafTracer(fx.Tracer):
    def __init__(self, leaf_types: Tuple[type, ...], autowrap_functions=(), autowrap_modules=()):
        super().__init__(autowrap_functions=tuple(autowrap_functions), autowrap_modules=tuple(autowrap_modules),
                         param_shapes_constant=True)
        self.leaf_types = leaf_types

    def is_leaf_module(self, m: nn.Module, qualname: str) -> bool:
        return isinstance(m, self.leaf_types) or super().is_leaf_module(m, qualname)


def trace_model(model: nn.Module, input_names: Optional[Sequence[str]], leaf_module_cls: Sequence[type],
                autowrap_functions=(), autowrap_modules=(), tracer_cls=None) -> fx.GraphModule:
    import inspect

    from ..parallel_layers import PARALLEL_FUNCTIONS, PARALLEL_MODULES

    from ..modules.rms_norm import RMSNorm
    from ..parallel_layers.layer_norm import LayerNorm

    leaf = tuple(leaf_module_cls) + tuple(PARALLEL_MODULES) + (RMSNorm, LayerNorm)
    sig = inspect.signature(model.forward)
    concrete = {}
    if input_names is not None:
        for name, p in sig.parameters.items():
            if name not in input_names and p.default is not inspect.Parameter.empty:
                concrete[name] = p.default
    tracer = (tracer_cls or _LeafTracer)(leaf, tuple(autowrap_functions) + tuple(PARALLEL_FUNCTIONS), autowrap_modules)
    graph = tracer.trace(model, concrete_args=concrete or None)
    _drop_specialised_placeholders(graph, concrete)
    return fx.GraphModule(model, graph)


def _drop_specialised_placeholders(graph: fx.Graph, concrete: Dict[str, Any]) -> None:
    """``concrete_args`` with a non-``None`` default leave a placeholder plus ``torch._assert(arg == default)`` in the graph;
    the pipeline never feeds those arguments (they are constants of the traced program), so remove the check and the
    placeholder instead of making every caller pass the default again."""
    for node in list(graph.nodes):
        if node.op != "placeholder" or not any(node.target == n or str(node.target).startswith(n + "_") for n in concrete):
            continue
        chain, frontier = [], [node]
        while frontier:
            cur = frontier.pop()
            for u in cur.users:
                if u not in chain:
                    chain.append(u)
                    frontier.append(u)
        if any(u.op == "output" for u in chain):
            continue                                   # really used by the program: leave it alone
        order = {n: i for i, n in enumerate(graph.nodes)}
        for u in sorted(chain, key=order.__getitem__, reverse=True):
            graph.erase_node(u)
        graph.erase_node(node)


@dataclass
class StageIO:
    inputs_from_model: List[str] = field(default_factory=list)      # placeholder names consumed directly
    inputs_from_prev: List[str] = field(default_factory=list)       # values received from stage-1 (ordered)
    outputs_to_next: List[str] = field(default_factory=list)        # values sent to stage+1 (ordered)
    call_args: List[str] = field(default_factory=list)              # argument names of the stage module, in order
    produces: List[Tuple[str, Optional[int]]] = field(default_factory=list)  # (name, index into the stage result or None)
    attr_args: Dict[str, str] = field(default_factory=dict)         # call arg name → qualified attribute (get_attr) on the root


def partition_traced_model(traced: fx.GraphModule, cut_after: Sequence[str], num_stages: int):
    """Split ``traced`` after each call_module node whose target is in ``cut_after``."""
    cut_set = set(cut_after)
    stage_of: Dict[fx.Node, int] = {}
    cur = 0
    for node in traced.graph.nodes:
        stage_of[node] = cur
        if node.op == "call_module" and node.target in cut_set:
            cur += 1
    assert cur + 1 == num_stages, f"cuts produce {cur + 1} stages, expected {num_stages}"
    split = split_module(traced, traced, lambda n: stage_of[n])
    return split


def analyze_pipeline_module(top_mod: fx.GraphModule) -> Tuple[List[StageIO], List[str]]:
    """Per-stage IO of a ``split_module`` result.  Returns (stage_ios, names of the final outputs)."""
    split = top_mod      # reference parameter names in the signature
    placeholders: Set[str] = set()
    produced_by: Dict[str, int] = {}
    stage_nodes: List[fx.Node] = []
    final_outputs: List[str] = []
    out_spec = None
    attrs: Dict[str, str] = {}
    for node in split.graph.nodes:
        if node.op == "placeholder":
            placeholders.add(node.name)
        elif node.op == "get_attr":
            attrs[node.name] = node.target   # parameters used outside leaf modules stay on the root (split_module)
        elif node.op == "call_module":
            stage_nodes.append(node)
        elif node.op == "output":
            out_spec = node.args[0]
    n = len(stage_nodes)
    ios = [StageIO() for _ in range(n)]
    # names: a stage result used via getitem has per-item names
    alias: Dict[str, Tuple[int, Optional[int]]] = {}
    for s, node in enumerate(stage_nodes):
        alias[node.name] = (s, None)
    for node in split.graph.nodes:
        if node.op == "call_function" and getattr(node.target, "__name__", "") == "getitem":
            src = node.args[0]
            if isinstance(src, fx.Node) and src.name in alias:
                alias[node.name] = (alias[src.name][0], node.args[1])
    for name, (s, _) in alias.items():
        produced_by[name] = s
    consumers: Dict[str, Set[int]] = {}
    for s, node in enumerate(stage_nodes):
        for a in list(node.args) + list(node.kwargs.values()):
            if isinstance(a, fx.Node):
                ios[s].call_args.append(a.name)
                if a.name in placeholders:
                    ios[s].inputs_from_model.append(a.name)
                elif a.name in attrs:
                    ios[s].attr_args[a.name] = attrs[a.name]
                else:
                    consumers.setdefault(a.name, set()).add(s)
    def _names(x):
        if isinstance(x, fx.Node):
            return [x.name]
        if isinstance(x, (tuple, list)):
            return [n for y in x for n in _names(y)]
        if isinstance(x, dict):
            return [n for y in x.values() for n in _names(y)]
        return []
    final_outputs = _names(out_spec)
    for name in final_outputs:
        if name in produced_by:
            consumers.setdefault(name, set()).add(n)  # virtual consumer after the last stage
    # a value produced at stage p and consumed at stage c > p crosses every boundary p..c-1
    for name, cons in consumers.items():
        p = produced_by.get(name)
        if p is None:
            continue
        last = max(cons)
        for b in range(p, min(last, n - 1)):
            if name not in ios[b].outputs_to_next:
                ios[b].outputs_to_next.append(name)
            if name not in ios[b + 1].inputs_from_prev:
                ios[b + 1].inputs_from_prev.append(name)
    for name, (s, idx) in alias.items():
        ios[s].produces.append((name, idx))
    return ios, final_outputs


def analyze_shared_weights_across_stages(top_module: fx.GraphModule, partitions: List[nn.Module]) -> List[List[Tuple[int, str]]]:
    """Groups of (stage, local parameter name) that are the same Parameter object in several stages."""
    split, stage_modules = top_module, partitions      # reference parameter names in the signature
    seen: Dict[int, List[Tuple[int, str]]] = {}
    for s, m in enumerate(stage_modules):
        for name, p in m.named_parameters(remove_duplicate=False):
            seen.setdefault(id(p), [])
            if not any(st == s for st, _ in seen[id(p)]):
                seen[id(p)].append((s, name))
    return [v for v in seen.values() if len(v) > 1]


# ---------------------------------------------------------------------------------------------------------------------
# IO annotation helpers with the reference's names (pipeline/partition.py:46-130)
# ---------------------------------------------------------------------------------------------------------------------
class PipelineIO:
    """Annotation of one value crossing a stage boundary: FX node ``name``; ``input_idx`` / ``output_idx`` = its position
    among the stage's inputs / outputs (``None`` = not used that way, i.e. passed along); ``metadata`` = tensor metas when it
    carries tensors; ``obj`` = the python object when it is not a tensor."""

    def __init__(self, name: str, input_idx: Optional[int] = None, output_idx: Optional[int] = None, metadata=None, obj=None):
        self.name, self.input_idx, self.output_idx, self.metadata, self.obj = name, input_idx, output_idx, metadata, obj

    def __repr__(self) -> str:
        return f"PipelineIO_{self.name}_input_idx_{self.input_idx}_output_idx_{self.output_idx}_metadata_{self.metadata}"


def adding_live_obj_for_previous_stages(stage_id_to_IO_input_names, stage_id_to_IO_output_names, obj_name: str,  # noqa: N803
                                        current_stage: int) -> None:
    """``obj_name`` is needed after ``current_stage`` but produced earlier: walk back to the producing stage and mark the
    value as an input AND output (pass-along) of every stage in between, so it is forwarded hop by hop."""
    if current_stage < 0:
        raise RuntimeError(f"{obj_name} is missing from all previous stages, stage_id_to_IO_output_names {stage_id_to_IO_output_names}")
    ins, outs = stage_id_to_IO_input_names[current_stage], stage_id_to_IO_output_names[current_stage]
    if obj_name not in ins and obj_name not in outs:
        adding_live_obj_for_previous_stages(stage_id_to_IO_input_names, stage_id_to_IO_output_names, obj_name, current_stage - 1)
        ins[obj_name] = PipelineIO(obj_name)
        outs[obj_name] = PipelineIO(obj_name)
    elif obj_name not in outs:                       # consumed here and needed later: forward it as well
        outs[obj_name] = PipelineIO(obj_name)


def iterate_graph_model_outputs(output_node_args):
    """Yield the values of an FX output node's ``args`` (always a 1-tuple holding either one node or a tuple of them)."""
    assert isinstance(output_node_args, tuple) and len(output_node_args) == 1, f"Unsupported output args found {output_node_args}"
    first = output_node_args[0]
    yield from (first if isinstance(first, (tuple, list)) else output_node_args)
