"""ZeRO-1 optimizer state through ``torch.distributed.checkpoint`` (reference ``optimizer/zero_dcp_utils.py:84-518``).
The flat fp32 shards (master weights + Adam moments) of every rank are described as slices of one global 1-D tensor
per param group and written/read with DCP's filesystem planner (one directory per ZeRO-1 rank); loading with a different
data-parallel degree re-slices the saved shards on the fly (``_reshard``), without the offline converter."""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

import torch
import torch.distributed as dist


def _to_dtensor_dict(inner) -> Dict[str, Any]:
    from torch.distributed._shard.sharded_tensor import Shard, ShardedTensor, ShardMetadata, ShardedTensorMetadata  # noqa: F401

    out: Dict[str, Any] = {}
    for g, fg in enumerate(inner.flat_groups):
        out[f"group{g}.master"] = fg.master_shard.detach().float().cpu()
        st = inner.base_optimizer.state.get(fg.base_param, {})
        for k, v in st.items():
            out[f"group{g}.{k}"] = v.detach().cpu() if isinstance(v, torch.Tensor) else torch.tensor(v)
    return out


MAX_RETRY = 100               # the reference polls for other ranks' files this often; loads here open files after a barrier


def _inner_of(optim_or_aux):
    """The ZeRO-1 optimizer from either the optimizer itself (possibly wrapped) or an ``aux_infos`` dict of
    :func:`get_dcp_aux_infos`."""
    if isinstance(optim_or_aux, dict):
        assert "optimizer" in optim_or_aux, "aux_infos must come from get_dcp_aux_infos(model, optimizer)"
        optim_or_aux = optim_or_aux["optimizer"]
    return optim_or_aux if hasattr(optim_or_aux, "pg") else getattr(optim_or_aux, "optimizer", optim_or_aux)


def save_optim_state_dict(path: str, state_dict: Dict[str, Any], aux_infos, dedup: bool = False) -> None:
    """Each rank writes its shard file set under ``path`` via DCP (one sub-directory per zero1 rank so that shards of
    differently-sized worlds never collide) plus a small layout file.  Third argument: the optimizer, or — the reference's
    signature (zero_dcp_utils.py:383-388) — the ``aux_infos`` of :func:`get_dcp_aux_infos`.  ``dedup`` (skip tensor-parallel
    duplicates) is moot: a rank's directory holds only the state it owns."""
    import torch.distributed.checkpoint as dcp

    inner = _inner_of(aux_infos)

    r = dist.get_rank(inner.pg)
    d = os.path.join(path, f"zero1_rank_{r:02d}_of_{dist.get_world_size(inner.pg):02d}")
    os.makedirs(d, exist_ok=True)
    dcp.save(_to_dtensor_dict(inner), checkpoint_id=d, no_dist=True)
    torch.save({"param_groups": state_dict["param_groups"], "shape_info": state_dict["shape_info"],
                "flat_layout": state_dict.get("flat_layout")}, os.path.join(d, "layout.pt"))


def _read_shard_dir(d: str) -> Dict[str, torch.Tensor]:
    """All tensors of one rank's DCP directory (targets allocated from the checkpoint's own metadata: a freshly built
    optimizer has no Adam moments yet, so its live state cannot serve as the template)."""
    import torch.distributed.checkpoint as dcp
    from torch.distributed.checkpoint.metadata import TensorStorageMetadata

    md = dcp.FileSystemReader(d).read_metadata()
    target = {k: torch.empty(tuple(v.size), dtype=v.properties.dtype) for k, v in md.state_dict_metadata.items()
              if isinstance(v, TensorStorageMetadata)}
    dcp.load(target, checkpoint_id=d, no_dist=True)
    return target


def _saved_world(path: str) -> int:
    import re

    worlds = {int(m.group(1)) for n in os.listdir(path) if (m := re.match(r"zero1_rank_\d+_of_(\d+)$", n))}
    if len(worlds) != 1:
        raise FileNotFoundError(f"expected the shard directories of exactly one ZeRO-1 world size under {path}, found {sorted(worlds)}")
    return worlds.pop()


def _reshard(path: str, saved_world: int, inner) -> Dict[str, torch.Tensor]:
    """This rank's slices cut out of a checkpoint written with another data-parallel degree.  The flat layout of a param
    group (parameter offsets) does not depend on the world size — only the tail padding does — so every 1-D state is a
    slice of ONE global vector ``cat(saved shards)``; rank r of the new world owns ``[r·S', (r+1)·S')`` of it (zeros beyond
    the saved length).  Only the saved shards that overlap that range are read."""
    r = dist.get_rank(inner.pg)
    cache: Dict[int, Dict[str, torch.Tensor]] = {}

    def shard(k: int) -> Dict[str, torch.Tensor]:
        if k not in cache:
            cache[k] = _read_shard_dir(os.path.join(path, f"zero1_rank_{k:02d}_of_{saved_world:02d}"))
        return cache[k]

    first = shard(0)
    out: Dict[str, torch.Tensor] = {}
    for g, fg in enumerate(inner.flat_groups):
        new_len = fg.master_shard.numel()
        lo, hi = r * new_len, (r + 1) * new_len
        for key, t0 in first.items():
            if not key.startswith(f"group{g}."):
                continue
            if t0.dim() == 0:                                    # replicated scalars (Adam ``step``)
                out[key] = t0
                continue
            old_len = t0.numel()
            piece = torch.zeros(new_len, dtype=t0.dtype)
            for k in range(lo // old_len, min(saved_world, (hi + old_len - 1) // old_len)):
                a, b = max(lo, k * old_len), min(hi, (k + 1) * old_len)
                if a < b:
                    piece[a - lo:b - lo] = shard(k)[key][a - k * old_len:b - k * old_len]
            out[key] = piece
    return out


def load_optim_state_dict(path: str, optimizer, aux_infos: Optional[Dict[str, Any]] = None, dedup: bool = False) -> Dict[str, Any]:
    """Load this rank's ZeRO-1 state.  Same data-parallel degree as at save time: read the rank's own directory; a different
    degree: re-slice the saved shards on the fly (reference: DCP load plans over ShardedTensors, ``zero_dcp_utils.py:329-370``)."""
    inner = _inner_of(optimizer)
    r, world = dist.get_rank(inner.pg), dist.get_world_size(inner.pg)
    d = os.path.join(path, f"zero1_rank_{r:02d}_of_{world:02d}")
    if os.path.isdir(d):
        target = _read_shard_dir(d)
        layout_dir = d
    else:
        saved = _saved_world(path)
        target = _reshard(path, saved, inner)
        layout_dir = os.path.join(path, f"zero1_rank_00_of_{saved:02d}")
    layout = torch.load(os.path.join(layout_dir, "layout.pt"), weights_only=False)
    base_state = {}
    smw = {}
    for g, fg in enumerate(inner.flat_groups):
        smw[g] = target[f"group{g}.master"]
        st = {}
        for k, v in target.items():
            if k.startswith(f"group{g}.") and not k.endswith(".master"):
                name = k.split(".", 1)[1]
                st[name] = v if v.dim() > 0 else v.item()
        base_state[g] = st
    return {"state": base_state, "base_state": base_state, "param_groups": layout["param_groups"],
            "shape_info": layout["shape_info"], "sharded_master_weights": smw}


def get_dcp_aux_infos(model: torch.nn.Module, optim) -> Dict[str, Any]:
    """Side information a DCP save/load of the optimizer needs (reference ``zero_dcp_utils.py:372-380``): which parameter
    object / parameter NAME every optimizer param id stands for, and the flat-shard ``shape_info`` of the ZeRO-1 state."""
    inner = getattr(optim, "optimizer", optim)
    names = {id(p): n for n, p in model.named_parameters()}
    pid_to_params, pid_to_names, pid = {}, {}, 0
    for group in inner.param_groups:
        for p in group["params"]:
            pid_to_params[pid] = p
            pid_to_names[pid] = names.get(id(p))
            pid += 1
    sd = inner.state_dict() if hasattr(inner, "state_dict") else {}
    return {"optim_pid_to_params": pid_to_params, "optim_pid_to_pnames": pid_to_names, "shape_info": sd.get("shape_info"),
            "optimizer": inner}
