"""Legacy sharded checkpoint format (reference ``parallel_layers/checkpointing.py:70-271``):

    <output_dir>/tp_rank_XX_pp_rank_XX[_dp_rank_XX]/checkpoint.pt

Only dp-rank 0 writes unless ``master_dp_only=False``.  ``load(..., sharded=False)`` takes a *full*
(unsharded) checkpoint and shards it on the fly using the parameters' parallel attributes, running
every module's ``preshard_hook`` first (padding, KV replication, QKV fusion).
"""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, Optional

import torch
import torch.distributed as dist
from torch import nn

from . import parallel_state as ps
from .utils import cast_all


NXD_SKIP_RENDEZVOUS = "NXD_SKIP_RENDEZVOUS"          # =1: no barriers around the filesystem operations (single-process tools)
PreShardHookFn = Callable[[nn.Module, dict, str], bool]   # module.preshard_hook(model_state_dict, prefix)


def _chkpt_dir(base: str, with_dp: bool = False) -> str:
    name = f"tp_rank_{ps.get_tensor_model_parallel_rank():02d}_pp_rank_{ps.get_pipeline_model_parallel_rank():02d}"
    if with_dp:
        name += f"_dp_rank_{ps.get_data_parallel_rank():02d}"
    return os.path.join(base, name)


def _barrier():
    if dist.is_initialized() and os.environ.get(NXD_SKIP_RENDEZVOUS, "0") != "1":
        dist.barrier()


def _to_cpu(obj):
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu()
    if isinstance(obj, dict):
        return {k: _to_cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cpu(v) for v in obj)
    return obj


def save(checkpoint: Dict[str, Any], output_dir: str, save_serially: bool = True, save_xser: bool = False,
         down_cast_bf16: bool = False, master_dp_only: bool = True) -> None:
    """Every (tp, pp) coordinate writes its own file; writes are staggered tp-rank by tp-rank when
    ``save_serially`` to bound host memory / filesystem pressure (reference :70-142)."""
    if master_dp_only and ps.get_data_parallel_rank() != 0:
        _barrier()
        return
    d = _chkpt_dir(output_dir, with_dp=not master_dp_only)
    os.makedirs(d, exist_ok=True)
    state = _to_cpu(checkpoint)
    if down_cast_bf16:
        state = cast_all(state, torch.float32, torch.bfloat16)
    if save_xser:
        from ..trainer.checkpoint import _flatten_tensors, _torch_xla_pickle_names

        tensors = []
        ref = _flatten_tensors(state, tensors)
        os.makedirs(os.path.join(d, "checkpoint.pt.tensors"), exist_ok=True)
        for i, t in enumerate(tensors):
            torch.save(t, os.path.join(d, "checkpoint.pt.tensors", f"tensor_{i}.pt"))
        with _torch_xla_pickle_names():                # stubs pickled as torch_xla.utils.serialization.TensorReference
            torch.save(ref, os.path.join(d, "checkpoint.pt"))
    else:
        torch.save(state, os.path.join(d, "checkpoint.pt"))
    _barrier()


def get_sharded_model_dict(model: nn.Module, model_state_dict: Dict[str, Any]) -> Dict[str, Any]:
    """Shard a full state dict for this rank using the parallel attributes on ``model``'s parameters."""
    from ..inference.sharding import shard_state_dict_for_rank

    return shard_state_dict_for_rank(model, model_state_dict, ps.get_tensor_model_parallel_rank(),
                                     ps.get_tensor_model_parallel_size())


def load(chkpt_path: str, model: Optional[nn.Module] = None, model_or_optimizer: Any = None, model_key: Optional[str] = "model",
         load_xser: bool = False, sharded: bool = True, strict: bool = True, master_dp_only: bool = True,
         weights_only: bool = False) -> Dict[str, Any]:
    """Load a legacy checkpoint.  ``sharded=True``: read this rank's ``tp_rank_XX_pp_rank_XX`` file (``master_dp_only=False``:
    the per-DP-rank file ``…_dp_rank_XX`` written by ``save(master_dp_only=False)`` — ZeRO-sharded optimizer state);
    ``sharded=False``: ``chkpt_path`` is one full checkpoint file that is sharded on the fly."""
    target = model if model is not None else model_or_optimizer
    if sharded:
        d = _chkpt_dir(chkpt_path, with_dp=not master_dp_only)
        f = os.path.join(d, "checkpoint.pt")
        from ..trainer.checkpoint import _torch_xla_pickle_names

        with _torch_xla_pickle_names():                # xser reference files (ours or the reference's) name that class
            ckpt = torch.load(f, map_location="cpu", weights_only=False)
        if load_xser or os.path.isdir(f + ".tensors"):
            from ..trainer.checkpoint import _unflatten_tensors

            n = len(os.listdir(f + ".tensors"))
            tensors = {i: torch.load(os.path.join(f + ".tensors", f"tensor_{i}.pt"), map_location="cpu") for i in range(n)}
            ckpt = _unflatten_tensors(ckpt, tensors)
    else:
        ckpt = torch.load(chkpt_path, map_location="cpu", weights_only=False)
    if target is not None:
        sd = ckpt[model_key] if (model_key is not None and model_key in ckpt) else ckpt
        if not sharded and isinstance(target, nn.Module):
            sd = get_sharded_model_dict(target, dict(sd))
        if isinstance(target, nn.Module):
            target.load_state_dict(sd, strict=strict)
        else:
            target.load_state_dict(sd)
    _barrier()
    return ckpt


def ensure_directory_exists(filename: str) -> None:
    """Create the parent directory of ``filename`` (reference checkpointing.py:26-29)."""
    d = os.path.dirname(filename)
    if d:
        os.makedirs(d, exist_ok=True)
