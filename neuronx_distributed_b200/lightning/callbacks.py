"""Callbacks (reference ``lightning/neuron_hooks_callback.py:9-218``, ``tqdm_progressbar.py``): dump per-layer forward
activations / backward gradients for selected steps; progress bar that prints on the loss-owning rank only."""
from __future__ import annotations

import os
from typing import Any, List, Optional

import torch

from ._compat import Callback


class NeuronHooksCallback(Callback):
    def __init__(self, dump_dir: str = "hooks_dump", steps: Optional[List[int]] = None, module_filter: Optional[str] = None,
                 dump_grads: bool = True):
        self.dump_dir, self.steps, self.module_filter, self.dump_grads = dump_dir, set(steps or [0]), module_filter, dump_grads
        self._handles: List[Any] = []
        self._step = 0

    def _want(self, name: str) -> bool:
        return self.module_filter is None or self.module_filter in name

    def attach(self, model: torch.nn.Module) -> None:
        os.makedirs(self.dump_dir, exist_ok=True)
        for name, mod in model.named_modules():
            if not name or not self._want(name) or list(mod.children()):
                continue

            def fwd(m, inp, out, _n=name):
                if self._step in self.steps:
                    t = out[0] if isinstance(out, (tuple, list)) else out
                    if isinstance(t, torch.Tensor):
                        torch.save(t.detach().float().cpu(), os.path.join(self.dump_dir, f"step{self._step}_{_n}_fwd.pt"))

            self._handles.append(mod.register_forward_hook(fwd))
            if self.dump_grads:
                def bwd(m, gin, gout, _n=name):
                    if self._step in self.steps and gout and isinstance(gout[0], torch.Tensor):
                        torch.save(gout[0].detach().float().cpu(), os.path.join(self.dump_dir, f"step{self._step}_{_n}_bwd.pt"))

                self._handles.append(mod.register_full_backward_hook(bwd))

    def on_train_start(self, trainer=None, pl_module=None) -> None:
        if pl_module is not None and getattr(pl_module, "model", None) is not None:
            self.attach(pl_module.model)

    def on_train_batch_end(self, *a, **k) -> None:
        self._step += 1

    def detach(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles.clear()


class NeuronTQDMProgressBar(Callback):
    def __init__(self, refresh_rate: int = 1):
        self.refresh_rate, self._bar = refresh_rate, None

    def on_train_start(self, trainer=None, pl_module=None) -> None:
        from tqdm import tqdm

        from .logger import NeuronTensorBoardLogger

        if NeuronTensorBoardLogger("", "").should_print():
            self._bar = tqdm(desc="train", unit="step")

    def on_train_batch_end(self, trainer=None, pl_module=None, outputs=None, *a, **k) -> None:
        if self._bar is not None:
            self._bar.update(1)
