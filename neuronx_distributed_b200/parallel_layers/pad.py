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


def get_number_of_extra_heads(num_heads: int, tp_degree: int) -> int:
    return (-num_heads) % tp_degree


def generate_padding_mask(num_heads: int, num_heads_with_pad: int, num_kv_heads: int, num_kv_heads_with_pad: int,
                          tp_degree: int, kv_layout: str = "tile"):
    """Boolean masks (True = real head) over the padded Q and KV head slots as laid out on this TP group.
    With ``kv_layout='tile'`` KV heads repeat K0..Kn,K0..Kn…; with ``'adjacent'`` each head repeats
    consecutively (the two replication layouts of :mod:`modules.qkv_linear`)."""
    q_mask = torch.arange(num_heads_with_pad) < num_heads
    kv_mask = torch.arange(num_kv_heads_with_pad) < num_kv_heads
    if num_kv_heads_with_pad > num_kv_heads and num_kv_heads_with_pad % num_kv_heads == 0 and num_kv_heads < tp_degree:
        rep = num_kv_heads_with_pad // num_kv_heads
        kv_mask = torch.ones(num_kv_heads_with_pad, dtype=torch.bool)  # replicated real heads, none padded
        del rep
    return q_mask, kv_mask


def _pad_dim(t: torch.Tensor, dim: int, new_size: int) -> torch.Tensor:
    if t.shape[dim] == new_size:
        return t
    shape = list(t.shape)
    shape[dim] = new_size - t.shape[dim]
    return torch.cat([t, torch.zeros(shape, dtype=t.dtype, device=t.device)], dim=dim)


def pad_model(model: nn.Module, tp_degree: int, n_heads: int, wrapped_classes: Sequence[type] = (),
              pad_hook_fn: Optional[Callable[[nn.Module, int, int], None]] = None) -> nn.Module:
    """Pad every head-sized TP projection from ``n_heads`` to the next multiple of ``tp_degree``.

    A Column layer whose *output* size is a multiple of ``n_heads`` (q/k/v/fused-qkv) grows along the
    output dim; a Row layer whose *input* size is a multiple of ``n_heads`` (o_proj) grows along the
    input dim.  Module attributes holding the head count (``num_heads``, ``num_attention_heads``…) on
    the parent are updated.  ``pad_hook_fn(module, tp_degree, n_heads)`` can customise further."""
    extra = get_number_of_extra_heads(n_heads, tp_degree)
    if extra == 0:
        return model
    tgt_heads = n_heads + extra
    tp_rank, tp = ps.get_tensor_model_parallel_rank(), ps.get_tensor_model_parallel_size()
    for parent in model.modules():
        if wrapped_classes and not isinstance(parent, tuple(wrapped_classes)):
            continue
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
            if pad_hook_fn is not None:
                pad_hook_fn(parent, tp_degree, n_heads)
    return model


def _copy_attrs(w_from: torch.Tensor, child: nn.Module, dim: int = 0) -> None:
    from .utils import set_tensor_model_parallel_attributes

    tp = ps.get_tensor_model_parallel_size()
    if not hasattr(child.weight, "tensor_model_parallel"):
        set_tensor_model_parallel_attributes(child.weight, True, dim, getattr(child, "stride", 1), num_partitions=tp)
