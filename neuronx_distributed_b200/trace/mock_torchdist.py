"""Pretend ``torch.distributed`` is initialised with ``world_size`` ranks — for building parallel modules (shapes,
parameter attributes, checkpoint sharding) in ONE process without any process group (reference ``trace/mock_torchdist.py``).
Collectives are NOT emulated: code that communicates must run under a real group."""
from __future__ import annotations

import contextlib
from typing import Any, Iterator
from unittest.mock import MagicMock

import torch


class MockDistributed(MagicMock):
    """Stand-in for the ``torch.distributed`` module: answers the topology queries, fakes group creation."""

    def __init__(self, *args, world_size: int = 1, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__["_world_size"] = world_size
        self.__dict__["_rank"] = 0
        self.__dict__["_alive"] = True

    def is_initialized(self) -> bool:
        return self.__dict__["_alive"]

    def is_available(self) -> bool:
        return True

    def init_process_group(self, backend=None, rank: int = 0, world_size: int = 1, **_):
        """Which rank this process pretends to be (reference :22-25: the rank whose shard is being built)."""
        self.__dict__["_world_size"], self.__dict__["_rank"], self.__dict__["_alive"] = world_size, rank, True

    def get_rank(self, group=None) -> int:
        ranks = getattr(group, "ranks", None)
        me = self.__dict__["_rank"]
        if ranks:
            return ranks.index(me) if me in ranks else -1
        return me

    def get_backend(self, group=None) -> str:
        return "gloo"

    def get_world_size(self, group=None) -> int:
        ranks = getattr(group, "ranks", None)
        return len(ranks) if ranks else self.__dict__["_world_size"]

    def get_process_group_ranks(self, group=None):
        return list(getattr(group, "ranks", range(self.__dict__["_world_size"])))

    def destroy_process_group(self, group=None) -> None:
        if group is None:
            self.__dict__["_alive"] = False

    def new_group(self, ranks=None, *args, **kwargs):
        g = MagicMock(spec=torch.distributed.ProcessGroup)
        g.ranks = list(ranks) if ranks is not None else list(range(self.__dict__["_world_size"]))
        g.size.return_value = len(g.ranks)
        me = self.__dict__["_rank"]
        g.rank.return_value = g.ranks.index(me) if me in g.ranks else -1
        return g

    def barrier(self, *a, **k) -> None:
        return None


@contextlib.contextmanager
def mock_distributed(world_size: int) -> Iterator[Any]:
    """``with mock_distributed(8): …`` — inside, ``torch.distributed`` reports an initialised 8-rank world (this process is
    rank 0).  The real module is restored on exit.  For building a specific rank's shard prefer
    ``trace.parallel_context.NxDParallelState`` which also sets the parallel-state overrides."""
    import sys

    real = torch.distributed
    mock = MockDistributed(world_size=world_size)
    for name in ("ProcessGroup", "ReduceOp", "P2POp", "Work", "GroupMember"):
        if hasattr(real, name):
            setattr(mock, name, getattr(real, name))
    torch.distributed = mock
    saved = sys.modules.get("torch.distributed")
    sys.modules["torch.distributed"] = mock
    try:
        yield mock
    finally:
        torch.distributed = real
        if saved is not None:
            sys.modules["torch.distributed"] = saved
