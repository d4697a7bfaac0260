"""Progress bar that only prints on the ranks that log (reference ``lightning/progress_bar.py:1-22``)."""
from __future__ import annotations

from typing import Optional

from ._compat import Callback


class NeuronTQDMProgressBar(Callback):
    def __init__(self, refresh_rate: int = 1):
        self.refresh_rate, self._bar = refresh_rate, None

    def setup(self, trainer=None, pl_module=None, stage: Optional[str] = None) -> None:
        """Progress output is decided when training starts (rank topology must be initialised first)."""

    def on_train_start(self, trainer=None, pl_module=None) -> None:
        from tqdm import tqdm

        from .logger import NeuronTensorBoardLogger

        if NeuronTensorBoardLogger("", "").should_print():
            self._bar = tqdm(desc="train", unit="step")

    def on_train_batch_end(self, trainer=None, pl_module=None, outputs=None, *a, **k) -> None:
        if self._bar is not None:
            self._bar.update(1)
