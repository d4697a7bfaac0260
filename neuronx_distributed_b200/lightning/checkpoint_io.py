"""Checkpoint IO plugin (reference ``lightning/checkpoint_io.py:15``): Lightning's save/load calls are mapped onto
``trainer.checkpoint.save_checkpoint / load_checkpoint`` so PTL runs produce the same sharded directory layout."""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

from ..trainer import checkpoint as ckpt
from ._compat import CheckpointIO


class NeuronCheckpointIO(CheckpointIO):
    def __init__(self, save_load_xser: bool = True, async_save: bool = False, num_kept_ckpts: Optional[int] = None):
        self.save_load_xser, self.async_save, self.num_kept_ckpts = save_load_xser, async_save, num_kept_ckpts

    def save_checkpoint(self, checkpoint: Dict[str, Any], path: str, storage_options: Optional[Any] = None) -> None:
        d, tag = os.path.split(path.rstrip("/"))
        ckpt.save_checkpoint(d, tag, model=checkpoint.get("state_dict"), optimizer=(checkpoint.get("optimizer_states") or [None])[0],
                             scheduler=(checkpoint.get("lr_schedulers") or [None])[0],
                             user_content={k: v for k, v in checkpoint.items() if k not in ("state_dict", "optimizer_states", "lr_schedulers")},
                             use_xser=self.save_load_xser, async_save=self.async_save, num_kept_ckpts=self.num_kept_ckpts)

    def load_checkpoint(self, path: str, map_location: Optional[Any] = None, model=None, optimizer=None, scheduler=None) -> Any:
        d, tag = os.path.split(path.rstrip("/"))
        return ckpt.load_checkpoint(d, tag, model=model, optimizer=optimizer, scheduler=scheduler)

    def remove_checkpoint(self, path: str) -> None:
        from ..trainer.checkpoint_storage import create_checkpoint_storage

        d, tag = os.path.split(path.rstrip("/"))
        create_checkpoint_storage(d).remove_dir(tag)
