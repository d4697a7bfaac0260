"""Rank-aware logging (behaviour of reference ``utils/logger.py:16-112``).

* level from ``NXD_LOG_LEVEL`` (off, fatal, error, warning, info, debug, trace)
* ``NXD_LOG_HIDE_TIME=1`` drops the timestamp
* ``rank0_only=True`` loggers are silent on every rank but 0
"""
from __future__ import annotations

import logging
import os
import sys
from typing import Any

_LEVELS = {
    "off": logging.CRITICAL + 10,
    "fatal": logging.CRITICAL,
    "error": logging.ERROR,
    "warning": logging.WARNING,
    "info": logging.INFO,
    "debug": logging.DEBUG,
    "trace": 5,
}
logging.addLevelName(5, "TRACE")


class PackagePathFilter(logging.Filter):
    """Adds ``record.relativepath`` = the source path relative to the longest ``sys.path`` entry containing it, so log lines
    show ``neuronx_distributed_b200/pipeline/model.py:123`` instead of an absolute path (reference logger.py:38-49)."""

    def filter(self, record: Any) -> bool:
        import sys

        record.relativepath = record.pathname
        for root in sorted((os.path.abspath(p) for p in sys.path if p), key=len, reverse=True):
            root = root if root.endswith(os.sep) else root + os.sep
            if record.pathname.startswith(root):
                record.relativepath = os.path.relpath(record.pathname, root)
                break
        return True


class _RankZeroFilter(logging.Filter):
    def filter(self, record: logging.LogRecord) -> bool:  # noqa: A003
        try:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                return dist.get_rank() == 0
        except Exception:
            pass
        return int(os.environ.get("RANK", "0")) == 0


_CACHE: dict[tuple[str, bool], logging.Logger] = {}


def get_log_level() -> int:
    name = os.environ.get("NXD_LOG_LEVEL", "info").lower()
    if name not in _LEVELS:
        raise ValueError(f"NXD_LOG_LEVEL={name!r} is not supported; use one of {sorted(_LEVELS)}")
    return _LEVELS[name]


def get_logger(name: str = "nxd_b200", rank0_only: bool = True) -> logging.Logger:
    key = (name, rank0_only)
    if key in _CACHE:
        return _CACHE[key]
    logger = logging.getLogger(f"{name}{'.r0' if rank0_only else '.all'}")
    level = get_log_level()
    if level > logging.CRITICAL:                         # "off": nothing is emitted, whatever the record's level
        logger.disabled = True
    else:
        logger.setLevel(level)
    logger.propagate = False
    if not logger.handlers:
        handler = logging.StreamHandler(sys.stdout)
        if os.environ.get("NXD_LOG_HIDE_TIME", "0") == "1":
            fmt = "[%(levelname).1s %(filename)s:%(lineno)d] %(message)s"
        else:
            fmt = "[%(asctime)s %(levelname).1s %(filename)s:%(lineno)d] %(message)s"
        handler.setFormatter(logging.Formatter(fmt))
        logger.addHandler(handler)
    if rank0_only:
        logger.addFilter(_RankZeroFilter())
    _CACHE[key] = logger
    return logger
