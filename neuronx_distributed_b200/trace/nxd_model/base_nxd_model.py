"""State initialiser and the interface of the runtime model (reference ``trace/nxd_model/base_nxd_model.py:11-182``)."""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
from torch import nn


class StateInitializer(nn.Module):
    """Allocates the state (KV-cache) buffers of a model: one zero tensor per ``(name → shape, dtype)`` entry on this rank's
    device (reference ``base_nxd_model.py:11-33`` creates them for every local rank of the process; here a process drives
    exactly one GPU)."""

    def __init__(self, shapes: Dict[str, Sequence[int]], dtypes: Dict[str, torch.dtype], local_ranks_size: int = 1):
        super().__init__()
        self.shapes, self.dtypes, self.local_ranks_size = dict(shapes), dict(dtypes), local_ranks_size

    def forward(self) -> List[Dict[str, torch.Tensor]]:
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        return [{k: torch.zeros(tuple(self.shapes[k]), dtype=self.dtypes[k], device=dev) for k in self.shapes}
                for _ in range(self.local_ranks_size)]


class BaseNxDModel(nn.Module):
    """Interface of the runtime model (reference ``base_nxd_model.py:36-182``)."""

    def add(self, key: str, trace_artifacts, compilation_artifacts):
        """Register a compiled bucket under ``key`` (reference: ``add(key, hlo, neff)``)."""
        raise NotImplementedError

    def get_available_keys(self):
        raise NotImplementedError

    def get_hlo(self, key: str):
        """The program of a bucket — here its launch plan (reference: the HLO module)."""
        raise NotImplementedError

    def get_neff(self, key: str) -> bytes:
        """The serialised executable of a bucket (reference: NEFF bytes)."""
        raise NotImplementedError

    def get_metaneff(self, key: str):
        """Input / output / state description of a bucket (reference: MetaNeff)."""
        raise NotImplementedError

    def read_from_neuron_buffer(self, state_buffer_key: str) -> torch.Tensor:
        """Copy of a state buffer (KV cache …) on the host."""
        raise NotImplementedError

    def write_to_neuron_buffer(self, tensor: torch.Tensor, state_buffer_key: str):
        """Overwrite a state buffer in place from a host tensor."""
        raise NotImplementedError

    def save(self, path_to_save: str, save_weights: bool = False):
        raise NotImplementedError

    @classmethod
    def load(cls, path_to_model: str, start_rank: int = 0):
        raise NotImplementedError

    def set_weights(self, sharded_checkpoint):
        raise NotImplementedError

    def to_neuron(self):
        raise NotImplementedError

    def replace_weights(self, sharded_checkpoint):
        raise NotImplementedError

    def router(self, inputs, arg_names=None):
        raise NotImplementedError
