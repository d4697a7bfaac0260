"""Autograd collectives for tensor / sequence / expert parallel regions.

Capability parity with reference ``parallel_layers/mappings.py`` (8 autograd Functions
:175-403, public helpers :406-478, EP enter/exit :481-555, SPMD scatter :558-678).

Each mapping is a (forward collective, backward collective) pair.  Rather than eight
hand-written Function classes, the pairs are generated from one table by
:func:`_make_mapping`; the primitive ops (identity / all-reduce / split / all-gather /
reduce-scatter / all-to-all) are defined once.  On CUDA with the ``fused`` backend the TP
linear layers bypass these for their hot paths (``ops/tp_fused.py``); the functions here
remain the semantic reference and serve the cold paths.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist

from . import comm
from . import parallel_state as ps


def _tp_group(group):
    return group if group is not None else ps.get_tensor_model_parallel_group()


def _size(group) -> int:
    return dist.get_world_size(group)


# --------------------------------------------------------------------------------------
# primitives: f(x, dim, group, dtype) -> tensor
# --------------------------------------------------------------------------------------
def _identity(x, dim, group, dtype):
    return x


def _all_reduce(x, dim, group, dtype):
    if _size(group) == 1:
        return x
    y = x.contiguous().clone() if x.requires_grad or not x.is_contiguous() else x.clone()
    comm.all_reduce(y, group=group)
    return y


def _split(x, dim, group, dtype):
    n = _size(group)
    if n == 1:
        return x
    assert x.shape[dim] % n == 0, f"cannot split dim {dim} of {tuple(x.shape)} into {n} parts"
    return x.chunk(n, dim=dim)[dist.get_rank(group)].contiguous()


def _all_gather(x, dim, group, dtype):
    return comm.all_gather(x, dim=dim, group=group)


def _reduce_scatter(x, dim, group, dtype):
    if _size(group) == 1:
        return x
    if dtype is not None and dtype != x.dtype:
        return comm.reduce_scatter(x.to(dtype), dim=dim, group=group).to(x.dtype)
    return comm.reduce_scatter(x, dim=dim, group=group)


def _make_mapping(name: str, fwd: Callable, bwd: Callable):
    class _Mapping(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, dim, group, dtype):
            ctx.dim, ctx.group, ctx.dtype = dim, group, dtype
            return fwd(x, dim, group, dtype)

        @staticmethod
        def backward(ctx, g):
            return bwd(g, ctx.dim, ctx.group, ctx.dtype), None, None, None

    _Mapping.__name__ = _Mapping.__qualname__ = name
    return _Mapping


# forward / backward pairs (reference mappings.py:175-352)
_CopyToModelParallelRegion = _make_mapping("_CopyToModelParallelRegion", _identity, _all_reduce)
_ReduceFromModelParallelRegion = _make_mapping("_ReduceFromModelParallelRegion", _all_reduce, _identity)
_ScatterToModelParallelRegion = _make_mapping("_ScatterToModelParallelRegion", _split, _all_gather)
_GatherFromModelParallelRegion = _make_mapping("_GatherFromModelParallelRegion", _all_gather, _split)
_ScatterToSequenceParallelRegion = _make_mapping("_ScatterToSequenceParallelRegion", _split, _all_gather)
_GatherFromSequenceParallelRegionRS = _make_mapping("_GatherFromSequenceParallelRegion", _all_gather, _reduce_scatter)
_GatherFromSequenceParallelRegionSplit = _make_mapping("_GatherFromSequenceParallelRegionNoTP", _all_gather, _split)
_ReduceScatterToSequenceParallelRegion = _make_mapping(
    "_ReduceScatterToSequenceParallelRegion", _reduce_scatter, _all_gather
)


class _AllToAllInExpertParallelRegion(torch.autograd.Function):
    """all-to-all over the EP group; backward swaps split/concat dims (reference :355-382)."""

    @staticmethod
    def forward(ctx, x, split_dim, concat_dim, group=None):
        ctx.split_dim, ctx.concat_dim, ctx.group = split_dim, concat_dim, group
        return comm.all_to_all(x, split_dim, concat_dim, group=group)

    @staticmethod
    def backward(ctx, g):
        return comm.all_to_all(g.contiguous(), ctx.concat_dim, ctx.split_dim, group=ctx.group), None, None, None


# --------------------------------------------------------------------------------------
# public helpers
# --------------------------------------------------------------------------------------
def copy_to_tensor_model_parallel_region(input_, process_group=None):
    return _CopyToModelParallelRegion.apply(input_, -1, _tp_group(process_group), None)


def reduce_from_tensor_model_parallel_region(input_, process_group=None):
    return _ReduceFromModelParallelRegion.apply(input_, -1, _tp_group(process_group), None)


def reduce_from_context_model_parallel_region(input_, process_group=None):
    group = process_group if process_group is not None else ps.get_context_model_parallel_group()
    return _ReduceFromModelParallelRegion.apply(input_, -1, group, None)


def scatter_to_tensor_model_parallel_region(input_, process_group=None):
    return _ScatterToModelParallelRegion.apply(input_, input_.dim() - 1, _tp_group(process_group), None)


def gather_from_tensor_model_parallel_region(input_, process_group=None):
    return _GatherFromModelParallelRegion.apply(input_, input_.dim() - 1, _tp_group(process_group), None)


def scatter_input_channels_to_tensor_model_parallel_region(input_, process_group=None):
    """Split NCHW activations along C (reference :385-403)."""
    return _ScatterToModelParallelRegion.apply(input_, 1, _tp_group(process_group), None)


def scatter_to_sequence_parallel_region(input_, sequence_dimension: int = 0, process_group=None):
    return _ScatterToSequenceParallelRegion.apply(input_, sequence_dimension, _tp_group(process_group), None)


def gather_from_sequence_parallel_region(
    input_,
    sequence_dimension: int = 0,
    to_model_parallel: bool = True,
    process_group=None,
    tile_cc: bool = False,
):
    """all-gather along the sequence dim.  ``to_model_parallel=True`` means the consumer is a
    TP region, so the gradient must be *reduce-scattered*; otherwise it is merely split
    (reference :280-319).  ``tile_cc`` (collective tiling hint for the Neuron compiler) has no
    CUDA meaning and is ignored."""
    del tile_cc
    fn = _GatherFromSequenceParallelRegionRS if to_model_parallel else _GatherFromSequenceParallelRegionSplit
    return fn.apply(input_, sequence_dimension, _tp_group(process_group), None)


def reduce_scatter_to_sequence_parallel_region(
    input_, sequence_dimension: int = 0, process_group=None, dtype: Optional[torch.dtype] = None
):
    return _ReduceScatterToSequenceParallelRegion.apply(input_, sequence_dimension, _tp_group(process_group), dtype)


def reduce_scatter_to_tensor_model_parallel_region_with_dim(input_, partition_dim: int, process_group=None):
    return _ReduceScatterToSequenceParallelRegion.apply(input_, partition_dim, _tp_group(process_group), None)


def gather_from_tensor_model_parallel_region_with_dim(input_, gather_dim: int, process_group=None):
    return _GatherFromSequenceParallelRegionSplit.apply(input_, gather_dim, _tp_group(process_group), None)


def all_to_all_in_expert_parallel_region(x, split_dim: int, concat_dim: int, process_group=None):
    group = process_group if process_group is not None else ps.get_expert_model_parallel_group()
    return _AllToAllInExpertParallelRegion.apply(x, split_dim, concat_dim, group)


def enter_expert_parallel_region(x: torch.Tensor, scatter_gather: bool) -> torch.Tensor:
    """(e, c, h) routed activations → (e/ep, ep, c, h): each EP rank receives, from every EP
    peer, the tokens for the experts it owns.  ``scatter_gather`` drops TP-duplicated tokens
    before the exchange and re-gathers after, cutting all-to-all volume by tp
    (reference :481-519)."""
    e, c, h = x.shape
    x = x.view(e, 1, c, h)
    if scatter_gather:
        x = scatter_to_sequence_parallel_region(x, 2)
    x = all_to_all_in_expert_parallel_region(x, 0, 1)
    if scatter_gather:
        x = gather_from_sequence_parallel_region(x, 2, to_model_parallel=False)
    return x


def exit_expert_parallel_region(x: torch.Tensor, scatter_gather: bool) -> torch.Tensor:
    """(e/ep, ep, c, h) → (e, c, h); inverse of :func:`enter_expert_parallel_region`
    (reference :522-555)."""
    if scatter_gather:
        x = scatter_to_sequence_parallel_region(x, 2)
    x = all_to_all_in_expert_parallel_region(x, 1, 0)
    if scatter_gather:
        x = gather_from_sequence_parallel_region(x, 2, to_model_parallel=False)
    return x.squeeze(1)


# --------------------------------------------------------------------------------------
# inference-only scatter with the rank supplied as a tensor (so one captured CUDA graph /
# traced program serves all ranks) — reference :558-678
# --------------------------------------------------------------------------------------
def scatter_to_process_group_spmd(input_: torch.Tensor, partition_dim: int, rank: torch.Tensor, process_group=None):
    """Return the ``rank``-th of n contiguous slices of ``x`` along ``partition_dim`` where
    ``rank`` is a 0-d/1-elem integer *tensor* (e.g. :class:`SPMDRank`'s weight)."""
    x = input_      # reference parameter names in the signature
    group = _tp_group(process_group)
    n = _size(group)
    if n == 1:
        return x
    size = x.shape[partition_dim] // n
    idx = rank.reshape(-1)[:1].to(torch.long) * size + torch.arange(size, device=x.device)
    return torch.index_select(x, partition_dim, idx)


def round_robin_scatter_to_process_group_spmd(
    input_: torch.Tensor, partition_dim: int, rank: torch.Tensor, process_group=None
):
    """Like :func:`scatter_to_process_group_spmd` but rank r takes elements r, r+n, r+2n, …"""
    x = input_      # reference parameter names in the signature
    group = _tp_group(process_group)
    n = _size(group)
    if n == 1:
        return x
    size = x.shape[partition_dim] // n
    idx = rank.reshape(-1)[:1].to(torch.long) + torch.arange(size, device=x.device) * n
    return torch.index_select(x, partition_dim, idx)


def nonzero_partition_dim_swap(func):
    """Decorator for ``f(x, partition_dim, …)`` collectives that are written for dim 0 only: transpose the partition dim to
    the front, call with dim 0, transpose back (reference mappings.py:27-40)."""
    import functools

    @functools.wraps(func)
    def wrapped(x, partition_dim: int, *args, **kwargs):
        if partition_dim % x.dim() == 0:
            return func(x, 0, *args, **kwargs)
        return func(x.transpose(0, partition_dim), 0, *args, **kwargs).transpose(0, partition_dim)

    return wrapped
