"""Shard a full (unsharded) state dict for a given TP rank from the parameters' parallel attributes.

Role parity with reference ``trace/trace.py:628-825`` (``get_sharded_checkpoint`` →
``preprocess_checkpoint`` (preshard hooks) → ``shard_children``) and
``parallel_layers/checkpointing.py:48-67``.  Rules: a parameter tagged ``tensor_model_parallel`` is cut
along ``partition_dim`` into ``num_partitions * partition_stride`` chunks and rank r takes chunks
``r, r+P, …`` (``create_local_weight``); fused QKV parameters are sharded per (q, k, v) section;
``rank_ordering`` permutes which logical shard a rank receives; everything else is replicated.
"""
from __future__ import annotations

from typing import Any, Dict

import torch
from torch import nn

from ..parallel_layers.utils import create_local_weight


def _run_preshard_hooks(model: nn.Module, sd: Dict[str, Any]) -> None:
    for name, module in model.named_modules():
        hook = getattr(module, "preshard_hook", None)
        if hook is None:
            continue
        prefix = (name + "." if name else "")
        for pname, _ in list(module.named_parameters(recurse=False)) or [("weight", None)]:
            key = prefix + pname
            try:
                hook(sd, key)
            except KeyError:
                pass
            break


def _select_local_experts(full: torch.Tensor, param: torch.Tensor, global_rank: int, world: int) -> torch.Tensor:
    """Expert-parallel parameters hold ``E/ep`` experts on dim 0: keep the experts owned by this rank's EP coordinate
    (reference trace/trace.py:762-776; rank layout ``[ep, tp]`` with TP fastest — ``parallel_state`` grid order)."""
    if not getattr(param, "expert_model_parallel", False) or full.shape[0] == param.shape[0]:
        return full
    from ..parallel_layers import parallel_state as ps

    ep = full.shape[0] // param.shape[0]
    tp = max(1, world // ep)
    ep_rank = (global_rank // tp) % ep
    dist_spec = getattr(param, "expert_distribution", None)
    ids = list(dist_spec[ep_rank]) if dist_spec is not None else ps.get_experts_for_expert_parallel_rank(ep_rank, full.shape[0], ep)
    idx = torch.as_tensor(ids, dtype=torch.long)
    if full.dtype in (torch.float8_e4m3fn, torch.float8_e5m2):                 # no fp8 gather on CPU: index the byte view
        return full.view(torch.int8)[idx].view(full.dtype)
    if full.dtype in (torch.uint16, torch.uint32):
        alias = torch.int16 if full.dtype == torch.uint16 else torch.int32
        return full.view(alias)[idx].view(full.dtype)
    return full[idx]


def shard_tensor(full: torch.Tensor, param: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    full = _select_local_experts(full, param, rank, world)
    if not getattr(param, "tensor_model_parallel", False):
        return full
    dim = param.partition_dim
    stride = getattr(param, "partition_stride", 1)
    nparts = getattr(param, "num_partitions", world)
    order = getattr(param, "rank_ordering", None)
    r = (order[rank] if order else rank) % nparts
    if getattr(param, "fused_qkv", False) and full.shape[dim] != param.shape[dim] * nparts:
        raise ValueError("fused qkv tensor has unexpected size; run preshard hooks first")
    if full.shape[dim] == param.shape[dim]:
        return full  # already local (e.g. SPMDRank handled by hook)
    per = full.shape[dim] // nparts
    return create_local_weight(full, dim, per, stride, rank=r, world_size=nparts).clone()


def shard_state_dict_for_rank(model: nn.Module, full_sd: Dict[str, Any], rank: int, world: int,
                              run_hooks: bool = True) -> Dict[str, Any]:
    sd = dict(full_sd)
    if run_hooks:
        _run_preshard_hooks(model, sd)
    out: Dict[str, Any] = {}
    params = dict(model.named_parameters(remove_duplicate=False))
    for k, v in sd.items():
        p = params.get(k)
        out[k] = shard_tensor(v, p, rank, world) if (p is not None and isinstance(v, torch.Tensor)) else v
    # parameters whose tensor lives under another key / format in the checkpoint (quantised layers reading torch
    # ``_packed_params`` entries): fetch through the parameter's ``get_tensor_from_state_dict`` hook (reference
    # quantization_layers.py:185-186, trace/trace.py shard_children)
    consumed = set()
    for k, p in params.items():
        getter = getattr(p, "get_tensor_from_state_dict", None)
        if getter is None:
            continue
        prefix = k[: k.rfind(".") + 1]
        try:
            t = getter(prefix=prefix, state_dict=sd)
        except (RuntimeError, KeyError):
            continue
        if t is not None and (k not in sd or t is not sd[k]):
            out[k] = shard_tensor(t, p, rank, world)
            consumed.update(x for x in sd if x.startswith(prefix + "_packed_params") or x == prefix + "zero_point")
    for x in consumed:
        out.pop(x, None)
    return out
