"""Latency-oriented all-reduce over NVLink peer memory (``csrc/allreduce.cu``): one kernel in which every rank writes its
vector into every peer's symmetric buffer and then sums the ``world`` slots locally.  Used for small tensors (decode-time
row-parallel outputs, embedding partial sums); anything above ``max_bytes`` stays on NCCL's bandwidth-optimal rings.
CUDA-graph safe (the epoch lives in device memory)."""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.distributed as dist

from ..utils.plan_registry import plan_op
from . import _ext, symm

_MAX_BYTES = int(os.environ.get("NXD_ONESHOT_AR_MAX_KB", "2048")) << 10
# "infer" (default): only under torch.no_grad()/inference_mode — the latency-bound decode path it was built for, validated on
# 2 GPUs incl. CUDA-graph replay and in the 4-GPU training bench; "1": always; "0": never (NCCL only).
_MODE = os.environ.get("NXD_ONESHOT_AR", "infer")
_STATE: Dict[int, tuple] = {}


def _workspace(group):
    st = _STATE.get(id(group))
    if st is None:
        world = dist.get_world_size(group)
        ws = symm.get_workspace(group, "oneshot_ar", 2 * world * _MAX_BYTES, nflags=world * 64)
        state = torch.zeros(2, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
        st = _STATE[id(group)] = (ws, state)
    return st


def reset() -> None:
    _STATE.clear()


def eligible(x: torch.Tensor, group) -> bool:
    if _MODE == "0" or (_MODE == "infer" and torch.is_grad_enabled()):
        return False
    if not (x.is_cuda and x.dtype in (torch.bfloat16, torch.float32) and x.is_contiguous()):
        return False
    nbytes = x.numel() * x.element_size()
    if nbytes == 0 or nbytes % 16 or nbytes > _MAX_BYTES:
        return False
    e = _ext.ext()
    if e is None or not hasattr(e, "oneshot_allreduce") or not symm.available():
        return False
    world = dist.get_world_size(group)
    return 1 < world <= 8 and dist.get_backend(group) == "nccl"


@plan_op("allreduce.all_reduce_sum", pure=True)
def all_reduce_sum(x: torch.Tensor, group) -> Optional[torch.Tensor]:
    """Returns the reduced tensor (new storage) or ``None`` when the caller should use NCCL.  Every rank of ``group`` must
    make the same decision — it depends only on shape/dtype, which are identical across ranks for TP collectives."""
    if not eligible(x, group):
        return None
    if os.environ.get("NXD_NVLS_AR", "1") == "1":
        from . import nvls

        # in-switch reduction (multimem.ld_reduce) when the group's symmetric region has an NVLS multicast mapping:
        # one local write + one reduced read per element instead of `world` peer writes + `world` local reads
        if nvls.available() and nvls.has_multicast(group):
            return nvls.all_reduce_sum(x, group)
    ws, state = _workspace(group)
    _ext.count_launch()
    return _ext.ext().oneshot_allreduce(x, ws.ptrs, ws.flag_ptrs, _MAX_BYTES, state, ws.rank, ws.world)
