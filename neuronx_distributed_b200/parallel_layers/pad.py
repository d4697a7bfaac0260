"""Pad attention heads so they divide the TP degree (reference ``parallel_layers/pad.py:14-186``).

``pad_model`` walks the model and grows every Column/Row parallel projection that is sized in heads:
extra heads are zero weights, so the function computed is unchanged while shapes become shardable.
``generate_padding_mask`` marks the padded head slots so they can be zeroed after QK^T / before o_proj.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch
from torch import nn

from . import parallel_state as ps
from .layers import ColumnParallelLinear, RowParallelLinear
from .utils import divide


def get_number_of_extra_heads(n_head: int, tp_degree: int) -> int:
    """Heads to add so that ``n_head`` divides ``tp_degree`` (reference :14-27)."""
    return (-n_head) % tp_degree


def generate_padding_mask(num_heads: int, num_heads_with_pad: int, num_kv_heads: int, tp_degree: int, tp_rank: int,
                          hardware_type=None, kv_layout: Optional[str] = None) -> torch.Tensor:
    """1-D mask over the query heads held by ``tp_rank`` (length ``num_heads_with_pad / tp_degree``): True for heads of the
    original model, False for heads that exist only because the query heads were padded (reference :114-185).

    With GQA and ``tp_degree > num_kv_heads`` the KV heads are replicated and the query heads of one KV head are spread over the
    ranks holding its replicas, so WHICH local heads are padding depends on the replication layout (``modules.qkv_linear``):
    ``"tile"`` (K0..Kn, K0..Kn, … — replica index = ``tp_rank // num_kv_heads``; the reference's first-generation layout) or
    ``"adjacent"`` (K0,K0,…,K1,K1,… — replica index = ``tp_rank % replicas``).  ``hardware_type`` is accepted for source
    compatibility: the strings ``"trn1"`` / ``"trn2"`` select ``"tile"`` / ``"adjacent"``; otherwise ``kv_layout`` (default
    ``"tile"``, the default of ``GQAQKVColumnParallelLinear``) decides."""
    if kv_layout is None:
        name = str(getattr(hardware_type, "value", hardware_type)).lower() if isinstance(hardware_type, str) else ""
        kv_layout = "adjacent" if name == "trn2" else "tile"
    per_rank = num_heads_with_pad // tp_degree
    per_kv = num_heads // num_kv_heads
    if kv_layout == "tile":
        replica = tp_rank // num_kv_heads
        limit = per_kv
    elif kv_layout == "adjacent":
        replicas = max(tp_degree // num_kv_heads, 1)
        replica = tp_rank % replicas
        limit = per_kv * ((num_kv_heads * replicas) // tp_degree)         # local KV heads × query heads per KV head
    else:
        raise RuntimeError(f"Unexpected KV replication layout {kv_layout!r} in padding mask generation.")
    mask = torch.arange(per_rank * replica, per_rank * (replica + 1)) < limit
    mask.requires_grad = False
    return mask


def generate_global_padding_masks(num_heads: int, num_heads_with_pad: int, num_kv_heads: int, num_kv_heads_with_pad: int,
                                  tp_degree: int):
    """Masks over ALL padded query / KV head slots of the unsharded model (True = real head; replicated KV heads are real)."""
    q_mask = torch.arange(num_heads_with_pad) < num_heads
    kv_mask = torch.arange(num_kv_heads_with_pad) < num_kv_heads
    if num_kv_heads_with_pad > num_kv_heads and num_kv_heads_with_pad % num_kv_heads == 0 and num_kv_heads < tp_degree:
        kv_mask = torch.ones(num_kv_heads_with_pad, dtype=torch.bool)
    return q_mask, kv_mask


def _pad_dim(t: torch.Tensor, dim: int, new_size: int) -> torch.Tensor:
    if t.shape[dim] == new_size:
        return t
    shape = list(t.shape)
    shape[dim] = new_size - t.shape[dim]
    return torch.cat([t, torch.zeros(shape, dtype=t.dtype, device=t.device)], dim=dim)


def pad_model(model: nn.Module, tp_degree: int, n_heads: int, wrapped_classes: Sequence[type] = (),
              pad_hook_fn: Optional[Callable[..., None]] = None) -> nn.Module:
    """Pad every head-sized TP projection from ``n_heads`` to the next multiple of ``tp_degree``.

    A Column layer whose *output* size is a multiple of ``n_heads`` (q/k/v/fused-qkv) grows along the
    output dim; a Row layer whose *input* size is a multiple of ``n_heads`` (o_proj) grows along the
    input dim.  Module attributes holding the head count (``num_heads``, ``num_attention_heads``…) on
    the parent are updated.

    ``wrapped_classes``: restrict the padding to instances of these classes AND everything below them (reference :79-83).
    ``pad_hook_fn(module, tgt_src_ratio)`` is called for every module in that scope with ``padded heads / heads`` (reference
    :59-63, e.g. to rescale a ``split_size`` attribute); a three-parameter hook receives ``(module, tp_degree, n_heads)``."""
    import inspect

    extra = get_number_of_extra_heads(n_heads, tp_degree)
    if extra == 0:
        return model
    tgt_heads = n_heads + extra
    tgt_src_ratio = tgt_heads / n_heads
    tp_rank, tp = ps.get_tensor_model_parallel_rank(), ps.get_tensor_model_parallel_size()
    wrapped = tuple(wrapped_classes)
    in_scope = set()

    def mark(mod, inside):
        inside = inside or not wrapped or isinstance(mod, wrapped)
        if inside:
            in_scope.add(id(mod))
        for child in mod.children():
            mark(child, inside)

    mark(model, False)
    hook_arity = None if pad_hook_fn is None else len([p for p in inspect.signature(pad_hook_fn).parameters.values()
                                                       if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)])
    for parent in model.modules():
        if id(parent) not in in_scope:
            continue
        if pad_hook_fn is not None and hook_arity == 2:
            pad_hook_fn(parent, tgt_src_ratio)
        touched = False
        for name, child in list(parent.named_children()):
            if isinstance(child, ColumnParallelLinear) and child.output_size % n_heads == 0 and not child.gather_output:
                head_dim = child.output_size // n_heads
                new_out = tgt_heads * head_dim
                per = divide(new_out, tp)
                # rebuild the local shard from a padded *full* view: gather is avoided by padding at the end of
                # the full dim, which lands entirely on the last ranks' shards
                full_lo = tp_rank * per
                w = child.weight.data
                old_per = w.shape[0]
                new_w = torch.zeros(per, w.shape[1], dtype=w.dtype, device=w.device)
                # rows of the old full matrix that fall in [full_lo, full_lo+per)
                old_full = old_per * tp
                take_lo, take_hi = min(full_lo, old_full), min(full_lo + per, old_full)
                if take_hi > take_lo and tp == 1:
                    new_w[: take_hi - take_lo] = w[take_lo:take_hi]
                elif tp > 1:
                    # with tp>1 the old sharding (old_per rows per rank) differs from the new one; keep local rows
                    # and pad locally — valid because extra heads are all-zero and head order is not semantic
                    new_w[:old_per] = w
                child.weight = nn.Parameter(new_w)
                _copy_attrs(w_from=w, child=child)
                child.output_size, child.output_size_per_partition = new_out, per
                if child.bias is not None:
                    child.bias = nn.Parameter(_pad_dim(child.bias.data, 0, per))
                touched = True
            elif isinstance(child, RowParallelLinear) and child.input_size % n_heads == 0 and child.input_is_parallel:
                head_dim = child.input_size // n_heads
                new_in = tgt_heads * head_dim
                per = divide(new_in, tp)
                w = child.weight.data
                child.weight = nn.Parameter(_pad_dim(w, 1, per))
                _copy_attrs(w_from=w, child=child, dim=1)
                child.input_size, child.input_size_per_partition = new_in, per
                touched = True
        if touched:
            for attr in ("num_heads", "num_attention_heads", "n_head", "n_heads"):
                if hasattr(parent, attr) and getattr(parent, attr) == n_heads:
                    setattr(parent, attr, tgt_heads)
            for attr in ("num_heads_local",):
                if hasattr(parent, attr):
                    setattr(parent, attr, tgt_heads // tp)
            if pad_hook_fn is not None and hook_arity != 2:
                pad_hook_fn(parent, tp_degree, n_heads)
    return model


def _copy_attrs(w_from: torch.Tensor, child: nn.Module, dim: int = 0) -> None:
    from .utils import set_tensor_model_parallel_attributes

    tp = ps.get_tensor_model_parallel_size()
    if not hasattr(child.weight, "tensor_model_parallel"):
        set_tensor_model_parallel_attributes(child.weight, True, dim, getattr(child, "stride", 1), num_partitions=tp)
