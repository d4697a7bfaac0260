"""Optimizer wrapper (reference ``trainer/optimizer.py:10-147``).

``step()`` = CP grad average → SP-param grad all-reduce → (non-ZeRO) DP bucket all-reduce
(+ EP second pass) → clip → inner step.  ``grad_norm`` exposes the last global norm."""
from __future__ import annotations

from typing import Any, List, Optional

import torch

from ..optimizer.zero_redundancy_optimizer import NeuronEPZero1Optimizer, Zero1Optimizer
from ..parallel_layers import grads
from ..parallel_layers import parallel_state as ps


class NxDOptimizer(torch.optim.Optimizer):
    def __init__(self, optimizer: torch.optim.Optimizer, nxd_config: dict):
        self.optimizer = optimizer
        self.nxd_config = nxd_config
        self._grad_norm: Optional[torch.Tensor] = None
        self._is_zero = isinstance(optimizer, (Zero1Optimizer, NeuronEPZero1Optimizer))

    # the wrapper owns no optimizer state of its own: these are live views of the wrapped optimizer (reference :30-55),
    # so LR schedulers that mutate ``param_groups`` and code that replaces ``state`` act on the real thing
    @property
    def state(self):
        return self.optimizer.state

    @state.setter
    def state(self, value) -> None:
        self.optimizer.state = value

    @property
    def param_groups(self):
        return self.optimizer.param_groups

    @param_groups.setter
    def param_groups(self, value) -> None:
        self.optimizer.param_groups = value

    @property
    def defaults(self):
        return getattr(self.optimizer, "defaults", {})

    @defaults.setter
    def defaults(self, value) -> None:
        self.optimizer.defaults = value

    def add_param_group(self, param_group) -> None:
        self.optimizer.add_param_group(param_group)

    def __getstate__(self):
        return {"optimizer": self.optimizer, "nxd_config": self.nxd_config, "_grad_norm": None, "_is_zero": self._is_zero}

    def __setstate__(self, state) -> None:
        self.__dict__.update(state)

    def __repr__(self) -> str:
        return f"NxDOptimizer({self.optimizer!r})"

    def save_state_dict(self, output_dir: str, num_workers_per_step: int = 8) -> None:
        """Deprecated in the reference (:150-155) in favour of ``save_checkpoint``; kept: per-(dp, tp) rank shard files."""
        assert self.nxd_config["optimizer_config"]["zero_one_enabled"], "save_state_dict needs the ZeRO-1 optimizer"
        self.optimizer.save_sharded_state_dict(output_dir, num_workers_per_step)

    def load_state_dict_from(self, output_dir: str, num_workers_per_step: int = 8) -> None:
        assert self.nxd_config["optimizer_config"]["zero_one_enabled"], "load_state_dict_from needs the ZeRO-1 optimizer"
        self.optimizer.load_sharded_state_dict(output_dir, num_workers_per_step)

    @property
    def grad_norm(self) -> Optional[torch.Tensor]:
        return self._grad_norm

    @property
    def params(self) -> List[torch.nn.Parameter]:
        return [p for g in self.param_groups for p in g["params"]]

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.optimizer.zero_grad(set_to_none=set_to_none)

    def no_sync(self):
        """Context manager for the non-final micro-batches of a gradient-accumulation step (DDP semantics): the ZeRO-1
        optimizer does not release gradient buckets to its overlapped reduce-scatter inside it.  A no-op otherwise."""
        inner = getattr(self.optimizer, "no_sync", None)
        if inner is not None:
            return inner()
        import contextlib

        return contextlib.nullcontext()

    def state_dict(self) -> Any:
        return self.optimizer.state_dict()

    def load_state_dict(self, state_dict: Any) -> None:
        sd = state_dict      # reference parameter names in the signature
        self.optimizer.load_state_dict(sd)

    def step(self, closure=None):
        cfg = self.nxd_config["optimizer_config"]
        grads.allreduce_context_parallel_gradients(self.optimizer)
        if self.nxd_config.get("sequence_parallel", False):
            grads.allreduce_sequence_parallel_gradients(self.optimizer)
        if self._is_zero:
            out = self.optimizer.step(closure=closure)
            self._grad_norm = self.optimizer.grad_norm
            return out
        pp_handles_dp = self.nxd_config.get("pipeline_parallel_size", 1) > 1 and \
            not (self.nxd_config.get("pipeline_config") or {}).get("use_optimizer_wrapper", True)
        if not pp_handles_dp and ps.get_data_parallel_size() > 1:
            allg = [p.grad for p in self.params if p.grad is not None]
            grads.bucket_allreduce_gradients(allg)
            if ps.get_expert_model_parallel_size() > 1:
                non_ep = [p.grad for p in self.params if p.grad is not None and not getattr(p, "expert_model_parallel", False)]
                grads.bucket_allreduce_gradients(non_ep, reduce_over_ep_group=True)
        if cfg["grad_clipping"]:
            self._grad_norm = grads.clip_grad_norm(self.params, cfg["max_grad_norm"])
        return self.optimizer.step(closure=closure)
