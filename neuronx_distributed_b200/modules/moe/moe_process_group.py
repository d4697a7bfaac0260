"""Hybrid prefill/decode sharding groups for MoE inference (reference ``modules/moe/moe_process_group.py``): prefill
may use (tp_cte, ep_cte) while decode uses (tp_tkg, ep_tkg) over the same ranks."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch.distributed as dist


@dataclass
class MoEProcessGroupInfo:
    tp_ranks: List[List[int]]
    ep_ranks: List[List[int]]
    tp_group: Optional[object] = None
    ep_group: Optional[object] = None


def build_moe_groups(world_size: int, tp: int, ep: int, create: bool = True) -> MoEProcessGroupInfo:
    assert world_size % (tp * ep) == 0
    grid = np.arange(world_size).reshape(-1, ep, tp)
    tp_ranks = grid.reshape(-1, tp).tolist()
    ep_ranks = np.moveaxis(grid, 1, -1).reshape(-1, ep).tolist()
    info = MoEProcessGroupInfo(tp_ranks, ep_ranks)
    if create and dist.is_initialized():
        me = dist.get_rank()
        for ranks in tp_ranks:
            g = dist.new_group(ranks)
            if me in ranks:
                info.tp_group = g
        for ranks in ep_ranks:
            g = dist.new_group(ranks)
            if me in ranks:
                info.ep_group = g
    return info


# ---------------------------------------------------------------------------------------------------------------------
# module-level registry with the reference's function names (moe_process_group.py:13-79)
# ---------------------------------------------------------------------------------------------------------------------
_MOE_TKG: Optional[MoEProcessGroupInfo] = None
_MOE_CTE: Optional[MoEProcessGroupInfo] = None


def init_tensor_expert_parallel_moe_process_groups(tkg_tp_degree: int, tkg_ep_degree: int, cte_tp_degree: int,
                                                   cte_ep_degree: int) -> None:
    """Create the decode (``tkg``) and prefill (``cte``) TP / EP groups for hybrid MoE sharding.  Collective: every rank of
    the world must call it.  Idempotent."""
    global _MOE_TKG, _MOE_CTE
    world = dist.get_world_size() if dist.is_initialized() else 1
    if _MOE_TKG is None:
        _MOE_TKG = build_moe_groups(world, tkg_tp_degree, tkg_ep_degree)
    if _MOE_CTE is None:
        _MOE_CTE = (_MOE_TKG if (cte_tp_degree, cte_ep_degree) == (tkg_tp_degree, tkg_ep_degree)
                    else build_moe_groups(world, cte_tp_degree, cte_ep_degree))


def _info(prefill: bool) -> MoEProcessGroupInfo:
    info = _MOE_CTE if prefill else _MOE_TKG
    assert info is not None, f"MoE {'CTE' if prefill else 'TKG'} process groups are not initialized"
    return info


def get_moe_tp_ep_group(prefill: bool = True):
    """TP group used inside the MoE block for prefill (``True``) or decode."""
    return _info(prefill).tp_group


def get_moe_ep_group(prefill: bool = True):
    return _info(prefill).ep_group


def get_moe_group_ranks(prefill: bool = True) -> MoEProcessGroupInfo:
    return _info(prefill)


def destroy_moe_model_parallel() -> None:
    global _MOE_TKG, _MOE_CTE
    _MOE_TKG = _MOE_CTE = None
