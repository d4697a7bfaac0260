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
