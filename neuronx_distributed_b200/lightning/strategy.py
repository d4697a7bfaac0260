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

    def _configure_launcher(self) -> None:
        from .launcher import _NeuronXLALauncher

        self._launcher = _NeuronXLALauncher(self)

    @property
    def root_device(self) -> torch.device:
        from ..utils import get_device

        return get_device()

    def model_to_device(self) -> None:
        """The NxD model is built on its device by ``initialize_parallel_model`` (meta-init + per-stage materialisation):
        nothing to move, like the reference (:131-133)."""

    def batch_to_device(self, batch: Any, device: Optional[torch.device] = None, dataloader_idx: int = 0) -> Any:
        from ..utils.device_loader import _map

        dev = device or self.root_device
        return _map(batch, lambda t: t.to(dev, non_blocking=True))

    def process_dataloader(self, dataloader):
        """Wrap the host dataloader in the device prefetcher (reference :201-218 wraps it in ``MpDeviceLoader``)."""
        from ..utils.device_loader import DevicePrefetchLoader

        return dataloader if isinstance(dataloader, DevicePrefetchLoader) else DevicePrefetchLoader(dataloader, self.root_device)

    def broadcast(self, obj, src: int = 0):
        return obj          # all ranks construct identical objects from the same seed/config

    def reduce(self, output=None, group: Optional[Any] = None, reduce_op: Optional[Any] = None, tensor=None):
        """Metric reduction across DATA-parallel replicas (TP / PP ranks of one replica hold the same value): ``mean`` /
        ``avg`` / ``sum``; non-tensors and single-replica runs pass through (reference :159-199)."""
        tensor = output if tensor is None else tensor            # ``output`` is the reference's / Lightning's argument name
        if not isinstance(tensor, torch.Tensor) or not dist.is_initialized() or not ps.model_parallel_is_initialized():
            return tensor
        n = ps.get_data_parallel_size()
        if n == 1:
            return tensor
        op = str(reduce_op).lower() if reduce_op is not None else "sum"
        if op not in ("mean", "avg", "sum", "reduceop.sum", "reduceop.avg"):
            raise ValueError(f"unsupported reduce_op {reduce_op!r}: use 'mean' / 'avg' / 'sum'")
        out = tensor.detach().clone().float()
        dist.all_reduce(out, group=group if group is not None else ps.get_data_parallel_group())
        if op in ("mean", "avg", "reduceop.avg"):
            out = out / n
        return out.to(tensor.dtype)

    # ---- checkpoints go through NeuronCheckpointIO (reference :220-238) ---------------------------------------------------
    @property
    def checkpoint_io(self):
        if getattr(self, "_ckpt_io", None) is None:
            from .checkpoint_io import NeuronCheckpointIO

            self._ckpt_io = NeuronCheckpointIO(save_load_xser=self.save_load_xser)
        return self._ckpt_io

    @checkpoint_io.setter
    def checkpoint_io(self, io) -> None:
        self._ckpt_io = io

    def save_checkpoint(self, checkpoint: Dict[str, Any], filepath: str, storage_options: Optional[Any] = None) -> None:
        self.checkpoint_io.save_checkpoint(checkpoint, filepath, storage_options)

    def load_checkpoint(self, checkpoint_path: str, **kwargs) -> Any:
        return self.checkpoint_io.load_checkpoint(checkpoint_path, **kwargs)

    def barrier(self, name: Optional[str] = None) -> None:
        if dist.is_initialized():
            dist.barrier()

    def teardown(self) -> None:
        if ps.model_parallel_is_initialized():
            ps.destroy_model_parallel()


NeuronXLAStrategy = NxDStrategy   # reference name
