"""Build a model for an arbitrary (tp_size, rank) in ONE process without any process group — role of reference
``trace/parallel_context.py:8-93`` / ``mock_torchdist.py:8-84`` (``NxDParallelState``): used for offline checkpoint
sharding and shape-only construction.  Implemented with the parallel-state rank/size overrides."""
from __future__ import annotations

import contextlib

from typing import Optional

import torch.distributed as dist

from ..parallel_layers import parallel_state as ps


class NxDParallelState(contextlib.AbstractContextManager):
    def __init__(self, world_size: int = 1, rank: int = 0, tensor_model_parallel_size: int = 1,
                 pipeline_model_parallel_size: int = 1, context_parallel_size: int = 1, expert_model_parallel_size: int = 1,
                 lnc_size: Optional[int] = None):
        """``lnc_size`` (logical cores fused into one rank on the reference hardware) is recorded only: a rank is one GPU."""
        self.lnc_size = lnc_size
        self.world_size, self.rank = world_size, rank
        self.tp, self.pp, self.cp, self.ep = (tensor_model_parallel_size, pipeline_model_parallel_size,
                                              context_parallel_size, expert_model_parallel_size)
        self._created_pg = False

    def __enter__(self):
        if not dist.is_initialized():
            import os

            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29577")
            dist.init_process_group("gloo", rank=0, world_size=1)
            self._created_pg = True
        if not ps.model_parallel_is_initialized():
            ps.initialize_model_parallel(1, 1, 1)
            self._init_ps = True
        else:
            self._init_ps = False
        mesh = ps.RankMesh(self.world_size, self.tp, self.pp, self.cp, self.ep)
        tp_rank = mesh.coords(self.rank)[3]
        self._saved = dict(ps._STATE.overrides)
        ps.set_tensor_model_parallel_size(self.tp)
        ps.set_tensor_model_parallel_rank(tp_rank)
        ps.set_aot_mode(False)
        return self

    def __exit__(self, *exc):
        ps._STATE.overrides.clear()
        ps._STATE.overrides.update(self._saved)
        if self._init_ps:
            ps.destroy_model_parallel()
        if self._created_pg:
            dist.destroy_process_group()
        return False
