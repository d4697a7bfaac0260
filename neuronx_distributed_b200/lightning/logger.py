"""Rank-aware scalar logger (reference ``lightning/logger.py:24-139``): only the rank that owns the loss (last PP stage,
tp-rank 0, dp-rank 0) writes.  Uses TensorBoard when available, else a JSON-lines file with the same scalars."""
from __future__ import annotations

import json
import os
import time
from typing import Dict, Optional

from ..parallel_layers import parallel_state as ps


class NeuronTensorBoardLogger:
    def __init__(self, save_dir: str, name: str = "default", version: Optional[str] = None, log_rank0: bool = False):
        self.dir = os.path.join(save_dir, name, version or "version_0")
        self.log_rank0 = log_rank0
        self._writer = None
        self._file = None

    def should_print(self) -> bool:
        if not ps.model_parallel_is_initialized():
            return True
        pp_ok = ps.get_pipeline_model_parallel_rank() == (0 if self.log_rank0 else ps.get_pipeline_model_parallel_size() - 1)
        return pp_ok and ps.get_tensor_model_parallel_rank() == 0 and ps.get_data_parallel_rank() == 0

    def _open(self):
        os.makedirs(self.dir, exist_ok=True)
        try:
            from torch.utils.tensorboard import SummaryWriter

            self._writer = SummaryWriter(self.dir)
        except Exception:
            self._file = open(os.path.join(self.dir, "scalars.jsonl"), "a")

    def log_metrics(self, metrics: Dict[str, float], step: Optional[int] = None) -> None:
        if not self.should_print():
            return
        if self._writer is None and self._file is None:
            self._open()
        if self._writer is not None:
            for k, v in metrics.items():
                self._writer.add_scalar(k, float(v), step)
        else:
            self._file.write(json.dumps({"step": step, "time": time.time(), **{k: float(v) for k, v in metrics.items()}}) + "\n")
            self._file.flush()

    # ---- TensorBoard-logger surface (reference logger.py:24-139) ----------------------------------------------------------
    @property
    def experiment(self):
        """The underlying writer (a ``SummaryWriter`` when TensorBoard is importable, else the JSON-lines file handle);
        created lazily on the printing rank only, ``None`` elsewhere."""
        if not self.should_print():
            return None
        if self._writer is None and self._file is None:
            self._open()
        return self._writer if self._writer is not None else self._file

    def print_step(self) -> bool:
        """Alias of :meth:`should_print` (reference spelling)."""
        return self.should_print()

    def log_graph(self, model, input_array=None) -> None:
        """Graph export is skipped: a TP/PP-sharded model has no single-process graph to draw."""

    def save(self) -> None:
        if self._writer is not None:
            self._writer.flush()
        if self._file is not None:
            self._file.flush()

    def finalize(self, status: str = "success") -> None:
        self.save()
        if self._writer is not None:
            self._writer.close()
            self._writer = None
        if self._file is not None:
            self._file.close()
            self._file = None
