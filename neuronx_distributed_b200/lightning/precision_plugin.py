"""Precision plugin (reference ``lightning/precision_plugin.py:11-36`` — ``NeuronXLAPrecisionPlugin``).

Precision is owned by the NxD model / optimizer wrappers (bf16 parameters, fp32 master weights and gradient accumulation
inside the ZeRO-1 optimizer), not by Lightning: the plugin therefore only forwards the optimizer step (manual
optimisation, as in the reference) and, with ``mixed_precision_enabled``, wraps forward in bf16 autocast."""
from __future__ import annotations

import contextlib
from typing import Any, Callable

import torch

try:  # pragma: no cover - lightning is not installed in the offline image
    from lightning.pytorch.plugins.precision import Precision as _Base
except Exception:  # noqa: BLE001
    class _Base:  # type: ignore[no-redef]
        pass


class NeuronXLAPrecisionPlugin(_Base):
    precision = "bf16-mixed"

    def __init__(self, mixed_precision_enabled: bool = False) -> None:
        self.mixed_precision_enabled = mixed_precision_enabled

    def forward_context(self):
        if not self.mixed_precision_enabled:
            return contextlib.nullcontext()
        return torch.autocast("cuda" if torch.cuda.is_available() else "cpu", dtype=torch.bfloat16)

    def optimizer_step(self, optimizer, model=None, closure: Callable[[], Any] = None, **kwargs: Any) -> Any:
        if closure is not None:
            closure()
        return optimizer.step()

    def clip_gradients(self, *args, **kwargs) -> None:
        """Clipping happens inside ``NxDOptimizer.step`` (global norm over TP/PP/EP groups)."""


B200PrecisionPlugin = NeuronXLAPrecisionPlugin
