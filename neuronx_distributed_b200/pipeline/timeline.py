"""Chrome-trace timeline of pipeline tasks (reference ``utils/timeline.py:15-140``,
``pipeline/timeline.py:10-21`` — force-disabled there; enabled here when ``trace_file_path`` is set).
Events carry host timestamps plus, on CUDA, device timestamps from CUDA events resolved at dump time."""
from __future__ import annotations

import json
import time
from typing import Any, Dict, List, Optional

import torch


class PPTimeline:
    def __init__(self, trace_file_path: Optional[str], pp_rank: int):
        self.path = trace_file_path
        self.enabled = trace_file_path is not None
        self.pp_rank = pp_rank
        self.events: List[Dict[str, Any]] = []
        self._open: Dict[str, Any] = {}
        self.step = 0

    @property
    def should_record(self) -> bool:
        """Recording is on whenever a trace file was given (the reference force-disables its pipeline timeline)."""
        return self.enabled

    def mark_event_start(self, label: str) -> None:
        if not self.enabled:
            return
        ev = torch.cuda.Event(enable_timing=True) if torch.cuda.is_available() else None
        if ev is not None:
            ev.record()
        self._open[label] = (time.time(), ev)

    def mark_event_end(self, label: str) -> None:
        if not self.enabled or label not in self._open:
            return
        t0, ev0 = self._open.pop(label)
        ev1 = torch.cuda.Event(enable_timing=True) if ev0 is not None else None
        if ev1 is not None:
            ev1.record()
        self.events.append({"name": label, "t0": t0, "t1": time.time(), "ev": (ev0, ev1), "step": self.step})

    def mark_step_end(self) -> None:
        if not self.enabled:
            return
        self.step += 1
        self.dump()

    def dump(self) -> None:
        if not self.enabled:
            return
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        out = []
        for e in self.events:
            dur_dev = e["ev"][0].elapsed_time(e["ev"][1]) * 1e3 if e["ev"][0] is not None else None
            out.append({"name": e["name"], "ph": "X", "pid": self.pp_rank, "tid": 0, "ts": e["t0"] * 1e6,
                        "dur": (e["t1"] - e["t0"]) * 1e6, "args": {"device_us": dur_dev, "step": e["step"]}})
        with open(f"{self.path}.pp{self.pp_rank}.json", "w") as f:
            json.dump({"traceEvents": out}, f)
