"""Host → device input pipeline (role of torch_xla's ``MpDeviceLoader`` that the reference wraps every dataloader in:
``lightning/strategy.py:201-218``, ``examples/training/*``): batches are pinned and copied to the GPU on a side stream
``prefetch`` steps ahead of the consumer, so the H2D copy of step ``n+1`` overlaps the compute of step ``n``; the consumer's
stream waits on the copy event only (no host synchronisation).  On CPU the loader is a pass-through."""
from __future__ import annotations

from collections import deque
from typing import Any, Iterable, Iterator, Optional

import torch


def _map(batch: Any, fn):
    if isinstance(batch, torch.Tensor):
        return fn(batch)
    if isinstance(batch, dict):
        return {k: _map(v, fn) for k, v in batch.items()}
    if isinstance(batch, (list, tuple)):
        out = [_map(v, fn) for v in batch]
        return type(batch)(out) if not hasattr(batch, "_fields") else type(batch)(*out)
    return batch


class DevicePrefetchLoader:
    def __init__(self, loader: Iterable, device: Optional[torch.device] = None, prefetch: int = 2, pin: bool = True):
        self.loader, self.prefetch, self.pin = loader, max(1, int(prefetch)), pin
        self.device = device if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
        self._stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self.bytes_copied = 0

    def __len__(self) -> int:
        return len(self.loader)  # type: ignore[arg-type]

    def _stage(self, batch: Any):
        if self._stream is None:
            return _map(batch, lambda t: t.to(self.device)), None

        def h2d(t: torch.Tensor) -> torch.Tensor:
            if t.device.type != "cpu":
                return t
            src = t.pin_memory() if (self.pin and not t.is_pinned()) else t
            self.bytes_copied += src.numel() * src.element_size()
            return src.to(self.device, non_blocking=True)

        with torch.cuda.stream(self._stream):
            dev = _map(batch, h2d)
            ev = torch.cuda.Event()
            ev.record(self._stream)
        return dev, ev

    def __iter__(self) -> Iterator[Any]:
        it = iter(self.loader)
        queue: deque = deque()
        try:
            while len(queue) < self.prefetch:
                queue.append(self._stage(next(it)))
        except StopIteration:
            pass
        while queue:
            dev, ev = queue.popleft()
            try:
                queue.append(self._stage(next(it)))
            except StopIteration:
                pass
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
                # the tensors were allocated on the copy stream: tell the allocator the consumer stream uses them
                _map(dev, lambda t: t.record_stream(torch.cuda.current_stream(self.device)) or t)
            yield dev


MpDeviceLoader = DevicePrefetchLoader      # torch_xla name used by ported training scripts
