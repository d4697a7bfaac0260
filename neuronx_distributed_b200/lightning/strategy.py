"""Strategy (reference ``lightning/strategy.py:36-238``): brings up ``torch.distributed`` (NCCL / gloo), the model-parallel
groups from ``nxd_config``, reports the *data-parallel* rank/size to the sampler, makes Lightning's broadcast/reduce
no-ops where the TP/PP engine already keeps ranks consistent, and routes checkpoints through ``NeuronCheckpointIO``."""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

import torch
import torch.distributed as dist

from ..parallel_layers import parallel_state as ps
from ..utils import cpu_mode, default_backend
from ._compat import _BaseStrategy


class NxDStrategy(_BaseStrategy):
    def __init__(self, nxd_config: Optional[Dict[str, Any]] = None, tensor_parallel_size: int = 1,
                 pipeline_parallel_size: int = 1, save_load_xser: bool = True, **kwargs):
        super().__init__(**kwargs) if kwargs else super().__init__()
        self.nxd_config = nxd_config
        self.tensor_parallel_size = nxd_config["tensor_parallel_size"] if nxd_config else tensor_parallel_size
        self.pipeline_parallel_size = nxd_config["pipeline_parallel_size"] if nxd_config else pipeline_parallel_size
        self.save_load_xser = save_load_xser

    def setup_distributed(self) -> None:
        if not dist.is_initialized():
            rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if not cpu_mode():
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
            dist.init_process_group(default_backend(), rank=rank, world_size=world)
        if not ps.model_parallel_is_initialized():
            cfg = self.nxd_config or {}
            ps.initialize_model_parallel(self.tensor_parallel_size, self.pipeline_parallel_size,
                                         cfg.get("expert_parallel_size", 1), context_parallel_size=cfg.get("context_parallel_size", 1))

    @property
    def distributed_sampler_kwargs(self) -> Dict[str, int]:
        return {"num_replicas": ps.get_data_parallel_size(), "rank": ps.get_data_parallel_rank()}

    @property
    def is_global_zero(self) -> bool:
        return (dist.get_rank() if dist.is_initialized() else 0) == 0

    def broadcast(self, obj, src: int = 0):
        return obj          # all ranks construct identical objects from the same seed/config

    def reduce(self, tensor, *a, **k):
        return tensor       # losses are already averaged over DP by the model/optimizer wrappers

    def barrier(self, name: Optional[str] = None) -> None:
        if dist.is_initialized():
            dist.barrier()

    def teardown(self) -> None:
        if ps.model_parallel_is_initialized():
            ps.destroy_model_parallel()


NeuronXLAStrategy = NxDStrategy   # reference name
