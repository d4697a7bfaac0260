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
                 log_rank0: bool = False, manual_opt: bool = True, train_batch_size: int = 16, logging_interval: int = 1):
        super().__init__()
        self.nxd_config, self.model_fn, self.opt_cls, self.scheduler_cls = nxd_config, model_fn, opt_cls, scheduler_cls
        self.model_args, self.model_kwargs = model_args, model_kwargs or {}
        self.opt_args, self.opt_kwargs = opt_args, opt_kwargs or {}
        self.scheduler_args, self.scheduler_kwargs = scheduler_args, scheduler_kwargs or {}
        self.grad_accum_steps, self.train_batch_size, self.logging_interval = grad_accum_steps, train_batch_size, logging_interval
        self.automatic_optimization = not manual_opt
        self.log_rank0 = log_rank0
        self.model = None
        self.averaged_loss = torch.zeros(())
        self._micro = 0
        self._t_last = None

    def setup(self, stage: Optional[str] = None, include_buffers: bool = False) -> None:
        """``include_buffers``: with meta-device initialisation also create the buffers on the meta device (reference
        module.py ``setup(stage, include_buffers)``)."""
        self.model = initialize_parallel_model(self.nxd_config, self.model_fn, include_buffers, *self.model_args, **self.model_kwargs)
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

    # ---- logging (reference :141-313 overrides ``LightningModule.log`` for device tensors and rank filtering) -------------
    @staticmethod
    def _to_tensor(value, name: str) -> torch.Tensor:
        if isinstance(value, torch.Tensor):
            if value.numel() != 1:
                raise ValueError(f"`self.log({name}, {value})` was called, but the tensor must have a single element.")
            return value.detach().reshape(())
        if isinstance(value, bool) or not isinstance(value, (int, float)):
            raise ValueError(f"`self.log({name}, {value!r})` was called, but `{type(value).__name__}` values cannot be logged.")
        return torch.tensor(float(value))

    def log(self, name: str, value, prog_bar: bool = False, logger: Optional[bool] = None, on_step: Optional[bool] = None,
            on_epoch: Optional[bool] = None, reduce_fx="mean", sync_dist: bool = False, rank_zero_only: bool = False,
            batch_size: Optional[int] = None, **kwargs) -> None:
        """Record a scalar metric.  ``log_rank0`` (constructor) / ``rank_zero_only`` keep non-zero ranks silent; ``sync_dist``
        averages over the data-parallel replicas through the strategy; values stay on their device (no ``.item()`` in the
        step — the trainer / logger reads them at its logging interval)."""
        if not isinstance(name, str):
            raise TypeError(f"metric names must be strings, got {type(name).__name__}")
        if isinstance(value, dict):
            raise ValueError(f"`self.log({name}, {value})` was called, but nested dictionaries cannot be logged; use log_dict")
        v = self._to_tensor(value, name)
        strategy = getattr(getattr(self, "trainer", None), "strategy", None)
        if sync_dist and strategy is not None and hasattr(strategy, "reduce"):
            v = strategy.reduce(v, reduce_op=reduce_fx if isinstance(reduce_fx, str) else "mean")      # collective: every rank
        if (self.log_rank0 or rank_zero_only) and torch.distributed.is_initialized() and torch.distributed.get_rank() != 0:
            return
        from ._compat import HAVE_LIGHTNING

        if HAVE_LIGHTNING and getattr(self, "_trainer", None) is not None:      # pragma: no cover - needs the lightning wheel
            return super().log(name, v, prog_bar=prog_bar, logger=logger, on_step=on_step, on_epoch=on_epoch, reduce_fx=reduce_fx,
                               sync_dist=False, rank_zero_only=rank_zero_only, batch_size=batch_size, **kwargs)
        self._logged[name] = v
        if prog_bar:
            self.__dict__.setdefault("_progress_bar_metrics", {})[name] = v

    def log_dict(self, dictionary: Dict[str, Any], **kwargs) -> None:
        for k, v in dictionary.items():
            self.log(k, v, **kwargs)

    # ---- hooks Lightning calls that the NxD wrappers already cover (reference :89-139) ---------------------------------
    def configure_gradient_clipping(self, *args, **kwargs) -> None:
        """Clipping runs inside ``NxDOptimizer.step`` / the ZeRO-1 optimizer (global norm over TP / PP / EP groups)."""

    def clip_gradients(self, *args, **kwargs) -> None:
        """See :meth:`configure_gradient_clipping`."""

    def on_train_batch_end(self, *args, **kwargs) -> None:
        """User hook."""

    def named_parameters(self, *args, **kwargs):
        if self.model is None:
            return iter(())
        return self.model.named_parameters(*args, **kwargs)

    def state_dict(self, *args, **kwargs):
        return self.model.state_dict() if self.model is not None else {}

    def load_state_dict(self, state_dict, strict: bool = True):
        return self.model.load_state_dict(state_dict, strict=strict)

    def get_param_groups_by_weight_decay(self, weight_decay: float = 0.01, no_decay=("bias", "norm")):
        """Two optimizer groups: decayed weights and un-decayed biases / norm gains.  Uses the LOCAL parameters of a
        pipeline-partitioned model.  Override for other policies."""
        m = self.model
        named = list(m.local_named_parameters()) if getattr(m, "partitioned", False) and hasattr(m, "local_named_parameters") \
            else list(m.named_parameters())
        return [{"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": weight_decay},
                {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
