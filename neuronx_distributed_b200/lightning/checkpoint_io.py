"""Checkpoint IO plugin (reference ``lightning/checkpoint_io.py:15``): Lightning's save/load calls are mapped onto
``trainer.checkpoint.save_checkpoint / load_checkpoint`` so PTL runs produce the same sharded directory layout."""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

from ..trainer import checkpoint as ckpt
from ._compat import CheckpointIO


class NeuronCheckpointIO(CheckpointIO):
    """Two on-disk layouts:

    * ``layout="trainer"`` (default): ``nxd.save_checkpoint`` tags — ``<dir>/<tag>/{model,optim,…}`` with completion markers,
      async save and retention;
    * ``layout="legacy"``: what the reference's plugin writes (``lightning/checkpoint_io.py``: the whole Lightning checkpoint
      dict through ``parallel_layers.checkpointing.save`` → ``<path>/tp_rank_XX_pp_rank_XX[_dp_rank_XX]/checkpoint.pt`` or its
      xser form), so checkpoints of runs made with the reference load unchanged (``master_dp_only``, ``weights_only``)."""

    def __init__(self, save_load_xser: bool = True, weights_only: bool = False, async_save: bool = False,
                 num_kept_ckpts: Optional[int] = None, layout: str = "trainer"):
        assert layout in ("trainer", "legacy"), layout
        self.save_load_xser, self.weights_only, self.layout = save_load_xser, weights_only, layout
        self.async_save, self.num_kept_ckpts = async_save, num_kept_ckpts

    def save_checkpoint(self, checkpoint: Dict[str, Any], path: str, storage_options: Optional[Any] = None,
                        master_dp_only: bool = True) -> None:
        if storage_options is not None:
            raise TypeError(f"`Trainer.save_checkpoint(..., storage_options=...)` with `storage_options` arg is not supported "
                            f"for `{self.__class__.__name__}`.")
        if self.layout == "legacy":
            from ..parallel_layers import checkpointing

            os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
            checkpointing.save(checkpoint=checkpoint, output_dir=path, save_xser=self.save_load_xser, master_dp_only=master_dp_only)
            return
        d, tag = os.path.split(path.rstrip("/"))
        ckpt.save_checkpoint(d, tag, model=checkpoint.get("state_dict"), optimizer=(checkpoint.get("optimizer_states") or [None])[0],
                             scheduler=(checkpoint.get("lr_schedulers") or [None])[0],
                             user_content={k: v for k, v in checkpoint.items() if k not in ("state_dict", "optimizer_states", "lr_schedulers")},
                             use_xser=self.save_load_xser, async_save=self.async_save, num_kept_ckpts=self.num_kept_ckpts)

    def load_checkpoint(self, checkpoint_path: Optional[str] = None, map_location: Optional[Any] = None, model=None, optimizer=None,
                        scheduler=None, master_dp_only: bool = True, path: Optional[str] = None) -> Any:
        checkpoint_path = checkpoint_path if checkpoint_path is not None else path
        if self.layout == "legacy":
            from ..parallel_layers import checkpointing

            return checkpointing.load(chkpt_path=checkpoint_path, load_xser=self.save_load_xser, master_dp_only=master_dp_only,
                                      weights_only=self.weights_only)
        d, tag = os.path.split(checkpoint_path.rstrip("/"))
        return ckpt.load_checkpoint(d, tag, model=model, optimizer=optimizer, scheduler=scheduler)

    def remove_checkpoint(self, path: str) -> None:
        from ..trainer.checkpoint_storage import create_checkpoint_storage

        d, tag = os.path.split(path.rstrip("/"))
        create_checkpoint_storage(d).remove_dir(tag)
