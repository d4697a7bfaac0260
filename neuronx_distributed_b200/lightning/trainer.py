"""A small ``Trainer`` with Lightning's fit-loop call order, so the PTL-style entry points (reference
``examples/training/llama/lightning/run_llama_nxd_ptl.py``: ``Trainer(strategy=…, callbacks=…, logger=…).fit(module,
datamodule=dm)``) run in images without the ``lightning`` package.  With Lightning installed use ``lightning.pytorch.Trainer``
— the strategy / module / checkpoint-IO / logger / callback classes of this package derive from its base classes then.

Order of calls: ``strategy.setup_distributed`` → ``datamodule.setup`` → ``module.setup`` → ``module.configure_optimizers`` →
(optional resume through the checkpoint IO) → callbacks ``setup`` / ``on_train_start`` → for every batch:
``module.training_step`` (manual optimisation: the module steps its optimizer) → callbacks ``on_train_batch_end`` → logger →
periodic checkpoint → callbacks ``on_train_end`` → ``strategy.teardown``."""
from __future__ import annotations

import os
from typing import Any, Dict, Iterable, List, Optional

import torch

from ..utils import get_device
from .checkpoint_io import NeuronCheckpointIO


def _to_device(batch, device):
    if isinstance(batch, torch.Tensor):
        return batch.to(device, non_blocking=True)
    if isinstance(batch, dict):
        return {k: _to_device(v, device) for k, v in batch.items()}
    if isinstance(batch, (list, tuple)):
        return type(batch)(_to_device(v, device) for v in batch)
    return batch


class Trainer:
    def __init__(self, strategy=None, callbacks: Optional[List[Any]] = None, logger: Any = None, max_steps: int = -1, max_epochs: int = 1,
                 log_every_n_steps: int = 1, default_root_dir: Optional[str] = None, every_n_train_steps: int = 0,
                 plugins: Optional[List[Any]] = None, enable_checkpointing: bool = True, **_ignored):
        self.strategy, self.callbacks, self.logger = strategy, list(callbacks or []), logger
        self.max_steps, self.max_epochs, self.log_every_n_steps = max_steps, max_epochs, max(1, log_every_n_steps)
        self.default_root_dir, self.every_n_train_steps = default_root_dir, every_n_train_steps
        self.checkpoint_io = next((p for p in (plugins or []) if isinstance(p, NeuronCheckpointIO)), None) or NeuronCheckpointIO(
            save_load_xser=getattr(strategy, "save_load_xser", True))
        self.enable_checkpointing = enable_checkpointing and default_root_dir is not None
        self.global_step, self.current_epoch = 0, 0
        self.lightning_module = None
        self.callback_metrics: Dict[str, Any] = {}

    # ---- helpers ------------------------------------------------------------------------------------------------------
    def _call(self, hook: str, *args, **kwargs) -> None:
        for cb in self.callbacks:
            fn = getattr(cb, hook, None)
            if fn is not None:
                fn(self, self.lightning_module, *args, **kwargs)

    def _ckpt_path(self, step: int) -> str:
        return os.path.join(self.default_root_dir, f"step_{step}")

    def save_checkpoint(self, path: Optional[str] = None) -> None:
        m = self.lightning_module
        self.checkpoint_io.save_checkpoint(
            {"state_dict": m.model, "optimizer_states": [m.optimizers()], "lr_schedulers": [m.lr_schedulers()],
             "global_step": self.global_step, "epoch": self.current_epoch}, path or self._ckpt_path(self.global_step))

    def _resume(self, ckpt_path: str) -> None:
        m = self.lightning_module
        user = self.checkpoint_io.load_checkpoint(ckpt_path, model=m.model, optimizer=m.optimizers(), scheduler=m.lr_schedulers())
        if isinstance(user, dict):
            self.global_step, self.current_epoch = int(user.get("global_step", 0)), int(user.get("epoch", 0))

    # ---- the loop -----------------------------------------------------------------------------------------------------
    def fit(self, model, train_dataloaders: Optional[Iterable] = None, datamodule=None, ckpt_path: Optional[str] = None) -> None:
        self.lightning_module = model
        model.trainer = self
        if self.strategy is not None:
            self.strategy.setup_distributed()
        if datamodule is not None:
            datamodule.trainer = self
            datamodule.setup("fit")
            train_dataloaders = datamodule.train_dataloader()
        model.setup("fit")
        model.configure_optimizers()
        if ckpt_path:
            self._resume(ckpt_path)
        self._call("setup", "fit")
        self._call("on_train_start")
        device = get_device()
        if self.strategy is not None and hasattr(self.strategy, "process_dataloader"):
            train_dataloaders = self.strategy.process_dataloader(train_dataloaders)      # pinned, prefetched H2D copies
        done = False
        while not done and self.current_epoch < self.max_epochs:
            for batch_idx, batch in enumerate(train_dataloaders):
                out = model.training_step(_to_device(batch, device), batch_idx)
                stepped = model._micro % model.grad_accum_steps == 0 if hasattr(model, "_micro") else True
                self._call("on_train_batch_end", out, batch, batch_idx)
                model.on_train_batch_end(out, batch, batch_idx)
                if stepped:
                    self.global_step += 1
                    self.callback_metrics = dict(getattr(model, "_logged", {}))
                    if self.logger is not None and self.global_step % self.log_every_n_steps == 0 and self.callback_metrics:
                        self.logger.log_metrics({k: float(v) for k, v in self.callback_metrics.items()}, step=self.global_step)
                    if self.enable_checkpointing and self.every_n_train_steps and self.global_step % self.every_n_train_steps == 0:
                        self.save_checkpoint()
                    if 0 < self.max_steps <= self.global_step:
                        done = True
                        break
            self.current_epoch += 1
        self._call("on_train_end")
        if self.logger is not None:
            self.logger.finalize("success")
        from ..trainer import checkpoint as _ckpt

        _ckpt.finalize_checkpoint()
