"""ZeRO-1 optimizer state through ``torch.distributed.checkpoint`` (reference ``optimizer/zero_dcp_utils.py:84-518``).
The flat fp32 shards (master weights + Adam moments) of every rank are described as slices of one global 1-D tensor
per param group and written/read with DCP's filesystem planner, so a checkpoint can be loaded with a different
data-parallel degree without the offline converter."""
from __future__ import annotations

import os
from typing import Any, Dict

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


def save_optim_state_dict(path: str, state_dict: Dict[str, Any], inner) -> None:
    """Each rank writes its shard file set under ``path`` via DCP (one sub-directory per zero1 rank so that shards of
    differently-sized worlds never collide) plus a small layout file."""
    import torch.distributed.checkpoint as dcp

    r = dist.get_rank(inner.pg)
    d = os.path.join(path, f"zero1_rank_{r:02d}_of_{dist.get_world_size(inner.pg):02d}")
    os.makedirs(d, exist_ok=True)
    dcp.save(_to_dtensor_dict(inner), checkpoint_id=d, no_dist=True)
    torch.save({"param_groups": state_dict["param_groups"], "shape_info": state_dict["shape_info"],
                "flat_layout": state_dict.get("flat_layout")}, os.path.join(d, "layout.pt"))


def load_optim_state_dict(path: str, inner) -> Dict[str, Any]:
    import torch.distributed.checkpoint as dcp

    r = dist.get_rank(inner.pg)
    d = os.path.join(path, f"zero1_rank_{r:02d}_of_{dist.get_world_size(inner.pg):02d}")
    # allocate the load targets from the checkpoint's own metadata: a freshly built optimizer has no Adam moments yet, so its
    # live state cannot serve as the template
    from torch.distributed.checkpoint.metadata import TensorStorageMetadata

    md = dcp.FileSystemReader(d).read_metadata()
    target = {k: torch.empty(tuple(v.size), dtype=v.properties.dtype) for k, v in md.state_dict_metadata.items()
              if isinstance(v, TensorStorageMetadata)}
    dcp.load(target, checkpoint_id=d, no_dist=True)
    layout = torch.load(os.path.join(d, "layout.pt"), weights_only=False)
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
    return {"optim_pid_to_params": pid_to_params, "optim_pid_to_pnames": pid_to_names, "shape_info": sd.get("shape_info")}
