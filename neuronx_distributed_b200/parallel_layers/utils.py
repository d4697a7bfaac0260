"""Parameter tagging + small tensor helpers (reference ``parallel_layers/utils.py:22-335``).

Every sharded parameter carries ``tensor_model_parallel, partition_dim, partition_stride,
num_partitions, rank_ordering``; the checkpoint sharder, the grad-norm code and the inference
weight sharder key off those attributes, so the names are part of the on-disk/API contract.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from ..utils import get_device
from . import parallel_state as ps

_MODEL_PARALLEL_ATTRIBUTE_DEFAULTS: Dict[str, Any] = {
    "tensor_model_parallel": False,
    "partition_dim": -1,
    "partition_stride": 1,
    "num_partitions": 1,
}


def ensure_divisibility(numerator: int, denominator: int) -> None:
    assert numerator % denominator == 0, f"{numerator} is not divisible by {denominator}"


def divide(numerator: int, denominator: int) -> int:
    ensure_divisibility(numerator, denominator)
    return numerator // denominator


def set_tensor_model_parallel_attributes(
    tensor: torch.Tensor,
    is_parallel: bool,
    dim: int,
    stride: int,
    num_partitions: int = -1,
    rank_ordering: Optional[Sequence[int]] = None,
) -> None:
    for attribute in _MODEL_PARALLEL_ATTRIBUTE_DEFAULTS:
        assert not hasattr(tensor, attribute), f"{attribute} already set"
    if num_partitions == -1:
        num_partitions = ps.get_tensor_model_parallel_size()
    if rank_ordering is not None:
        tp = ps.get_tensor_model_parallel_size()
        if sorted(rank_ordering) != list(range(tp)):
            raise ValueError(f"rank_ordering {rank_ordering} must be a permutation of 0..{tp - 1}")
    tensor.tensor_model_parallel = is_parallel
    tensor.partition_dim = dim
    tensor.partition_stride = stride
    tensor.num_partitions = num_partitions
    tensor.rank_ordering = list(rank_ordering) if rank_ordering is not None else None


def set_defaults_if_not_set_tensor_model_parallel_attributes(tensor: torch.Tensor) -> None:
    for attribute, value in _MODEL_PARALLEL_ATTRIBUTE_DEFAULTS.items():
        if not hasattr(tensor, attribute):
            setattr(tensor, attribute, value)


def copy_tensor_model_parallel_attributes(destination_tensor: torch.Tensor, source_tensor: torch.Tensor) -> None:
    for attribute in list(_MODEL_PARALLEL_ATTRIBUTE_DEFAULTS) + ["rank_ordering", "sequence_parallel_enabled", "shared"]:
        if hasattr(source_tensor, attribute):
            setattr(destination_tensor, attribute, getattr(source_tensor, attribute))


def param_is_not_tensor_parallel_duplicate(param: torch.Tensor) -> bool:
    return bool(getattr(param, "tensor_model_parallel", False)) or ps.get_tensor_model_parallel_rank() == 0


def param_is_not_shared(param: torch.Tensor) -> bool:
    return not getattr(param, "shared", False)


def split_tensor_along_dim(
    tensor: torch.Tensor, dim: int, num_partitions: int, contiguous_split_chunks: bool = False
) -> List[torch.Tensor]:
    size = divide(tensor.size(dim), num_partitions)
    chunks = list(torch.split(tensor, size, dim=dim))
    return [c.contiguous() for c in chunks] if contiguous_split_chunks else chunks


def split_tensor_along_last_dim(tensor, num_partitions, contiguous_split_chunks: bool = False):
    return split_tensor_along_dim(tensor, tensor.dim() - 1, num_partitions, contiguous_split_chunks)


def split_tensor_along_second_dim(tensor, num_partitions, contiguous_split_chunks: bool = False):
    return split_tensor_along_dim(tensor, 1, num_partitions, contiguous_split_chunks)


class EmbeddingUtility:
    """[first, last) vocabulary index range owned by a rank."""

    @staticmethod
    def range_from_per_partition_vocab_size(per_partition_vocab_size: int, rank: int, world_size: int) -> Tuple[int, int]:
        first = rank * per_partition_vocab_size
        return first, first + per_partition_vocab_size

    @staticmethod
    def range_from_global_vocab_size(global_vocab_size: int, rank: int, world_size: int) -> Tuple[int, int]:
        return EmbeddingUtility.range_from_per_partition_vocab_size(
            divide(global_vocab_size, world_size), rank, world_size
        )


def shard_slices(full_size: int, num_partitions: int, stride: int, rank: int) -> List[slice]:
    """Index ranges of the *full* dimension that belong to ``rank`` under strided sharding.

    The full dim is cut into ``num_partitions*stride`` equal chunks and rank r owns chunks
    ``r, r+P, r+2P, …`` — with stride 2 a fused [gate; up] weight shards so that every rank
    gets matching gate and up slices (reference layers.py:87-106)."""
    chunk = divide(full_size, num_partitions * stride)
    return [slice((rank + s * num_partitions) * chunk, (rank + s * num_partitions + 1) * chunk) for s in range(stride)]


def create_local_weight(
    full_weight: torch.Tensor,
    partition_dim: int,
    per_partition_size: int,
    stride: int,
    out_weight: Optional[torch.Tensor] = None,
    rank: Optional[int] = None,
    world_size: Optional[int] = None,
) -> torch.Tensor:
    """This rank's shard of ``full_weight`` (strided; see :func:`shard_slices`)."""
    rank = ps.get_tensor_model_parallel_rank() if rank is None else rank
    world_size = ps.get_tensor_model_parallel_size() if world_size is None else world_size
    assert per_partition_size * world_size == full_weight.shape[partition_dim]
    pieces = [
        full_weight.narrow(partition_dim, s.start, s.stop - s.start)
        for s in shard_slices(full_weight.shape[partition_dim], world_size, stride, rank)
    ]
    with torch.no_grad():
        local = pieces[0] if len(pieces) == 1 else torch.cat(pieces, dim=partition_dim)
        if out_weight is not None:
            out_weight.copy_(local)
            return out_weight
        return local.contiguous()


def gather_full_weight(shards: Sequence[torch.Tensor], partition_dim: int, stride: int) -> torch.Tensor:
    """Inverse of :func:`create_local_weight` given all ranks' shards in rank order."""
    n = len(shards)
    if stride == 1:
        return torch.cat(list(shards), dim=partition_dim)
    split = [torch.chunk(s, stride, dim=partition_dim) for s in shards]
    return torch.cat([split[r][s] for s in range(stride) for r in range(n)], dim=partition_dim)


def move_model_to_device(model: torch.nn.Module, device: Optional[torch.device] = None) -> torch.nn.Module:
    """``model.to(device)`` that preserves the parallel attributes on parameters
    (``Module.to`` re-creates Parameters on some paths; reference utils.py:236-259)."""
    device = device if device is not None else get_device()
    saved = {n: {k: v for k, v in p.__dict__.items()} for n, p in model.named_parameters()}
    model.to(device)
    for n, p in model.named_parameters():
        for k, v in saved.get(n, {}).items():
            if not hasattr(p, k):
                setattr(p, k, v)
    return model


def is_torch_version_greater_than_2() -> bool:
    return int(torch.__version__.split(".")[0]) >= 2


def cast_tensor(tensor, from_dtype=torch.float32, to_dtype=torch.bfloat16):
    t = tensor      # reference parameter names in the signature
    return t.to(to_dtype) if isinstance(t, torch.Tensor) and t.dtype == from_dtype else t


def cast_all(state, from_dtype=torch.float32, to_dtype=torch.bfloat16):
    """Recursively cast tensors in nested containers (reference utils.py:262-290)."""
    if isinstance(state, torch.Tensor):
        return cast_tensor(state, from_dtype, to_dtype)
    if isinstance(state, dict):
        return type(state)((k, cast_all(v, from_dtype, to_dtype)) for k, v in state.items())
    if isinstance(state, (list, tuple)):
        return type(state)(cast_all(v, from_dtype, to_dtype) for v in state)
    return state


def get_local_world_size() -> int:
    import os

    return int(os.environ.get("LOCAL_WORLD_SIZE", "1"))


def requires_init_pg_override() -> bool:
    return False


def get_padding_length(numerator: int, denominator: int) -> int:
    """Elements to append so that ``numerator`` becomes a multiple of ``denominator`` (reference utils.py:128-135)."""
    return (-numerator) % denominator


def is_pjrt_device() -> bool:
    """There is no XLA runtime here (reference utils.py:171); kept so that ported scripts can branch on it."""
    return False


def _autocast_dtype() -> torch.dtype:
    return torch.get_autocast_dtype("cuda" if torch.cuda.is_available() else "cpu")


def _autocast_on() -> bool:
    return torch.is_autocast_enabled("cuda") if torch.cuda.is_available() else torch.is_autocast_enabled("cpu")


def _cast_nested(value, dtype):
    if isinstance(value, torch.Tensor):
        return value.to(dtype) if value.is_floating_point() and value.dtype is not torch.float64 else value
    if isinstance(value, (str, bytes)) or type(value).__module__ == "numpy":
        return value
    if isinstance(value, dict):
        return {k: _cast_nested(v, dtype) for k, v in value.items()}
    if isinstance(value, (list, tuple)):
        return type(value)(_cast_nested(v, dtype) for v in value)
    return value


def cast_if_autocast_enabled(*args):
    """Cast floating tensors nested in ``args`` to the active autocast dtype (reference utils.py:202-207)."""
    return args if not _autocast_on() else _cast_nested(args, _autocast_dtype())


def verify_casted_dtype(value) -> None:
    """Assert that every tensor nested in ``value`` already has the autocast dtype (reference utils.py:269-286)."""
    if not _autocast_on():
        return
    if isinstance(value, torch.Tensor):
        assert value.dtype == _autocast_dtype(), f"Datatype of tensor is expected to be {_autocast_dtype()}, got {value.dtype} instead"
    elif isinstance(value, dict):
        for v in value.values():
            verify_casted_dtype(v)
    elif isinstance(value, (list, tuple)):
        for v in value:
            verify_casted_dtype(v)


def move_all_tensor_to_cpu(data, convert: bool = True):
    """Synchronise and (optionally) copy every device tensor nested in ``data`` to host memory — one stream sync for the
    whole structure instead of one per tensor (role of reference utils.py:230-243)."""
    if torch.cuda.is_available():
        torch.cuda.current_stream().synchronize()
    if not convert:
        return data

    def walk(v):
        if isinstance(v, torch.Tensor):
            return v.to("cpu") if v.device.type != "cpu" else v
        if isinstance(v, dict):
            return type(v)((k, walk(x)) for k, x in v.items())
        if isinstance(v, (list, tuple)):
            return type(v)(walk(x) for x in v)
        return v

    return walk(data)


def indices_split_along_dim(tensor: Optional[torch.Tensor], dim: int, rank: int, num_partitions: int) -> Optional[torch.Tensor]:
    """Index vector selecting partition ``rank`` of ``num_partitions`` contiguous slices of ``tensor`` along ``dim``
    (reference utils.py:288-316) — used with ``index_select`` where the rank is a device tensor (SPMDRank)."""
    if tensor is None:
        return None
    per = divide(tensor.size(dim), num_partitions)
    return torch.arange(per, device=tensor.device) + rank * per


def initialize_fallback_parallel_state(device: Optional[torch.device] = None) -> None:
    """Single-process world for layers constructed without ``initialize_model_parallel`` (reference utils.py:318-331)."""
    from . import parallel_state

    parallel_state.initialize_fallback_parallel_state()
