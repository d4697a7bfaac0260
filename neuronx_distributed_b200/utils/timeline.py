"""Multi-rank Chrome-trace timeline (reference ``utils/timeline.py:7-140``).

``mark_event_start(label)`` / ``mark_event_end(label)`` bracket an event on this rank; ``mark_step_end()`` gathers every
rank's events and rank 0 appends them to the trace file (``chrome://tracing`` / Perfetto "B"/"E" pairs, ``pid`` = rank).
Subclasses say when to record (``should_record``) and how to gather (``_collect_events_for_all_ranks``);
``pipeline.timeline.PPTimeline`` is the pipeline engine's variant with CUDA-event device times."""
from __future__ import annotations

import json
import time
from abc import ABC, abstractmethod
from collections import OrderedDict
from typing import Any, List, Optional


class Event:
    def __init__(self, label: str, rank: int, start: float = -1, end: float = -1):
        self.label, self.rank, self.start, self.end = label, rank, start, end


class Timeline(ABC):
    def __init__(self, trace_file_path: Optional[str], rank: int):
        self.enabled = trace_file_path is not None
        self.trace_file_path, self.rank, self.step = trace_file_path, rank, 0
        self.current_rank_events: "OrderedDict[str, Event]" = OrderedDict()
        self.all_rank_events: Optional[List[Any]] = None
        if self.enabled and self.should_record and rank == 0:
            with open(trace_file_path, "a") as f:
                f.write("[")                              # the trace viewers accept an unterminated JSON array

    @property
    @abstractmethod
    def should_record(self) -> bool:
        ...

    @abstractmethod
    def _collect_events_for_all_ranks(self) -> None:
        """Fill ``self.all_rank_events`` with one ``{label: Event}`` mapping per rank."""

    def _get_timestamp(self) -> float:
        return time.time() * 1e6

    def mark_event_start(self, label: str) -> None:
        if not (self.enabled and self.should_record):
            return
        assert label not in self.current_rank_events, f"event {label!r} is already open"
        self.current_rank_events[label] = Event(label, self.rank, start=self._get_timestamp())

    def mark_event_end(self, label: str) -> None:
        if not (self.enabled and self.should_record):
            return
        assert label in self.current_rank_events, f"event {label!r} was never started"
        self.current_rank_events[label].end = self._get_timestamp()

    def mark_step_end(self) -> None:
        if not (self.enabled and self.should_record):
            return
        self._collect_events_for_all_ranks()
        if self.rank == 0:
            self._dump_events()
        self._clean_states()
        self.step += 1

    def _clean_states(self) -> None:
        self.current_rank_events = OrderedDict()
        self.all_rank_events = None

    @staticmethod
    def _trace_line(ph: str, label: str, ts: float, pid: int, tid: int = 0) -> str:
        return json.dumps({"cat": "comp", "ph": ph, "name": label, "ts": ts, "tid": tid, "pid": pid}) + ",\n"

    def _creat_sync_event(self, ev: Event) -> List[str]:          # (sic) reference spelling
        assert ev.start != -1 and ev.end != -1, f"event {ev.label!r} is incomplete"
        return [self._trace_line("B", ev.label, ev.start, ev.rank), self._trace_line("E", ev.label, ev.end, ev.rank)]

    def _create_instant_event(self, label: str, timestamp: float) -> str:
        return self._trace_line("i", label, timestamp, self.rank, self.step)

    def _dump_events(self) -> None:
        assert self.all_rank_events is not None
        with open(self.trace_file_path, "a") as f:
            for events in self.all_rank_events:
                for ev in events.values():
                    f.writelines(self._creat_sync_event(ev))


class DistributedTimeline(Timeline):
    """Concrete timeline over a ``torch.distributed`` group (``all_gather_object`` of the step's events)."""

    def __init__(self, trace_file_path: Optional[str], rank: Optional[int] = None, group=None, record: bool = True):
        import torch.distributed as dist

        self._group, self._record = group, record
        super().__init__(trace_file_path, rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0))

    @property
    def should_record(self) -> bool:
        return self._record

    def _collect_events_for_all_ranks(self) -> None:
        import torch.distributed as dist

        if not dist.is_initialized() or dist.get_world_size(self._group) == 1:
            self.all_rank_events = [self.current_rank_events]
            return
        out: List[Any] = [None] * dist.get_world_size(self._group)
        dist.all_gather_object(out, self.current_rank_events, group=self._group)
        self.all_rank_events = out
