"""Model-parallel RNG streams (reference ``parallel_layers/random.py:20-127``).

Two streams per rank: the *default* stream is seeded identically inside a TP group (dropout
on replicated activations must match across TP ranks); the *model-parallel* stream is seeded
``seed + 2718 + tp_rank`` (sharded-weight init and dropout on sharded activations must
differ).  ``fork()`` temporarily swaps the device RNG state to a named stream.
"""
from __future__ import annotations

import contextlib
from typing import Dict

import torch

from ..utils import cpu_mode
from . import parallel_state as ps

_MODEL_PARALLEL_RNG_TRACKER_NAME = "model-parallel-rng"
_TP_SEED_OFFSET = 2718


def _get_state() -> torch.Tensor:
    return torch.get_rng_state() if cpu_mode() else torch.cuda.get_rng_state()


def _set_state(state: torch.Tensor) -> None:
    if cpu_mode():
        torch.set_rng_state(state)
    else:
        torch.cuda.set_rng_state(state)


def _seed(seed: int) -> None:
    if cpu_mode():
        torch.manual_seed(seed)
    else:
        torch.cuda.manual_seed(seed)


class RNGStatesTracker:
    def __init__(self) -> None:
        self.states_: Dict[str, torch.Tensor] = {}
        self.seeds_ = set()

    def reset(self) -> None:
        self.states_, self.seeds_ = {}, set()

    def get_states(self) -> Dict[str, torch.Tensor]:
        return dict(self.states_)

    def set_states(self, states: Dict[str, torch.Tensor]) -> None:
        self.states_ = states

    def add(self, name: str, seed: int) -> None:
        if seed in self.seeds_:
            raise RuntimeError(f"seed {seed} already exists")
        if name in self.states_:
            raise RuntimeError(f"rng state {name} already exists")
        self.seeds_.add(seed)
        orig = _get_state()
        _seed(seed)
        self.states_[name] = _get_state()
        _set_state(orig)

    @contextlib.contextmanager
    def fork(self, name: str = _MODEL_PARALLEL_RNG_TRACKER_NAME):
        if name not in self.states_:
            # un-seeded use (e.g. unit tests constructing layers directly): behave as identity
            yield
            return
        orig = _get_state()
        _set_state(self.states_[name])
        try:
            yield
        finally:
            self.states_[name] = _get_state()
            _set_state(orig)


_TRACKER = RNGStatesTracker()


def get_rng_tracker() -> RNGStatesTracker:
    return _TRACKER


# reference name kept as an alias so user code ports unchanged
get_xla_rng_tracker = get_rng_tracker
XLARNGStatesTracker = RNGStatesTracker       # reference class name (random.py:20); the states here are CUDA / CPU generator states


def model_parallel_manual_seed(seed: int) -> None:
    tp_seed = seed + _TP_SEED_OFFSET + ps.get_tensor_model_parallel_rank()
    _TRACKER.reset()
    torch.manual_seed(seed)
    if not cpu_mode():
        torch.cuda.manual_seed(seed)
    _TRACKER.add(_MODEL_PARALLEL_RNG_TRACKER_NAME, tp_seed)


model_parallel_xla_manual_seed = model_parallel_manual_seed
