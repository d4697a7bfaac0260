"""``NeuronLTModule`` (reference ``lightning/module.py:24-322``): a LightningModule with *manual optimisation* that
builds the parallel model and optimizer from ``nxd_config``, runs ``model.run_train`` (works with and without pipeline
parallelism), steps the NxD optimizer wrapper / scheduler, and logs loss, lr, grad-norm and throughput."""
from __future__ import annotations

import time
from typing import Any, Callable, Dict, Optional, Tuple

import torch

from ..trainer import initialize_parallel_model, initialize_parallel_optimizer
from ._compat import LightningModule


class NeuronLTModule(LightningModule):
    def __init__(self, nxd_config: Dict[str, Any], model_fn: Callable, opt_cls: Callable, scheduler_cls: Optional[Callable] = None,
                 model_args: Tuple = (), model_kwargs: Optional[Dict] = None, opt_args: Tuple = (), opt_kwargs: Optional[Dict] = None,
                 scheduler_args: Tuple = (), scheduler_kwargs: Optional[Dict] = None, grad_accum_steps: int = 1,
                 log_rank0: bool = False, manual_opt: bool = True, train_batch_size: int = 1, logging_interval: int = 1):
        super().__init__()
        self.nxd_config, self.model_fn, self.opt_cls, self.scheduler_cls = nxd_config, model_fn, opt_cls, scheduler_cls
        self.model_args, self.model_kwargs = model_args, model_kwargs or {}
        self.opt_args, self.opt_kwargs = opt_args, opt_kwargs or {}
        self.scheduler_args, self.scheduler_kwargs = scheduler_args, scheduler_kwargs or {}
        self.grad_accum_steps, self.train_batch_size, self.logging_interval = grad_accum_steps, train_batch_size, logging_interval
        self.automatic_optimization = not manual_opt
        self.model = None
        self.averaged_loss = torch.zeros(())
        self._micro = 0
        self._t_last = None

    def setup(self, stage: Optional[str] = None) -> None:
        self.model = initialize_parallel_model(self.nxd_config, self.model_fn, False, *self.model_args, **self.model_kwargs)
        self.averaged_loss = torch.zeros((), device=next(self.model.parameters()).device)

    def configure_optimizers(self):
        opt = initialize_parallel_optimizer(self.nxd_config, self.opt_cls, self.model.parameters(), *self.opt_args, **self.opt_kwargs)
        self._optimizers = opt
        if self.scheduler_cls is None:
            return opt
        sch = self.scheduler_cls(opt.optimizer if hasattr(opt, "optimizer") else opt, *self.scheduler_args, **self.scheduler_kwargs)
        self._schedulers = sch
        return [opt], [{"scheduler": sch, "interval": "step"}]

    def training_step(self, batch: Dict[str, torch.Tensor], batch_idx: int = 0):
        opt = self.optimizers()
        sch = self.lr_schedulers()
        loss = self.model.run_train(**batch)
        self.averaged_loss = self.averaged_loss + loss.detach().to(self.averaged_loss.device) / self.grad_accum_steps
        self._micro += 1
        if self._micro % self.grad_accum_steps == 0:
            opt.step()
            opt.zero_grad()
            if sch is not None:
                sch.step()
            now = time.time()
            if self._t_last is not None:
                self.log("throughput_seq_per_s", self.train_batch_size * self.grad_accum_steps / max(now - self._t_last, 1e-9))
            self._t_last = now
            self.log("loss", self.averaged_loss.detach())
            if getattr(opt, "grad_norm", None) is not None:
                self.log("grad_norm", opt.grad_norm)
            if sch is not None:
                self.log("lr", sch.get_last_lr()[0])
            self.averaged_loss = torch.zeros_like(self.averaged_loss)
        return loss

    def forward(self, *a, **k):
        return self.model(*a, **k)
