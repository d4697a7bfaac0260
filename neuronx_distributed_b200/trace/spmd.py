"""v1 runtime pieces (reference ``trace/spmd.py:11-291``): bucket container, state initialiser, runtime model."""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
from torch import nn

from .nxd_model import BucketProgram, NxDModel  # noqa: F401
from .nxd_model import StateInitializer as _StateInitializer


def default_bucket_kernel(inputs: List[torch.Tensor]):
    """Bucket selector used when a key has a single bucket: inputs unchanged, bucket index 0."""
    return inputs, torch.tensor(0, dtype=torch.int)


class SPMDBucketModelScript(nn.Module):
    """The buckets of one key; ``forward(inputs, bucket_idx_tensor)`` runs bucket ``bucket_idx`` (reference :14-76)."""

    def __init__(self, compiled_models: Sequence[BucketProgram]):
        super().__init__()
        self.models = list(compiled_models)

    def _pick(self, bucket_idx_tensor) -> BucketProgram:
        i = int(bucket_idx_tensor)
        if not 0 <= i < len(self.models):
            raise ValueError(f"bucket index {i} out of range (0..{len(self.models) - 1})")
        return self.models[i]

    def forward(self, inputs: List[torch.Tensor], bucket_idx_tensor: torch.Tensor):
        return self._pick(bucket_idx_tensor)(*inputs)

    def forward_ranked(self, input_collection: List[List[torch.Tensor]], bucket_idx_tensor: torch.Tensor):
        return [self._pick(bucket_idx_tensor)(*inputs) for inputs in input_collection]

    forward_async = forward_ranked                         # CUDA launches are asynchronous already


class StateInitializer(_StateInitializer):
    """Adds the reference's ``combine_kv_on_device`` option (:79-137): consecutive ``…past_key_values…`` K and V entries are
    allocated as ONE ``[2, …]`` tensor and exposed as two views — one allocation and, for the decode kernel, K and V of a
    layer adjacent in memory."""

    def __init__(self, shapes, dtypes, local_ranks_size: int = 1, combine_kv_on_device: bool = False):
        super().__init__(shapes, dtypes, local_ranks_size)
        self.kv_cache_keys_map: Dict[str, Tuple[str, str]] = {}
        kv_keys = [k for k in self.shapes if combine_kv_on_device and ".past_key_values." in k]
        self.state_keys: List[str] = [k for k in self.shapes if k not in kv_keys]
        it = iter(kv_keys)
        for idx, (k_key, v_key) in enumerate(zip(it, it)):
            if self.dtypes[k_key] != self.dtypes[v_key] or math.prod(self.shapes[k_key]) != math.prod(self.shapes[v_key]):
                raise ValueError("Could not combine KV allocations due to incompatible dtype or shape")
            kv_key = k_key.rsplit(".", 1)[0] + f".combined.{idx}"
            self.state_keys.append(kv_key)
            self.dtypes[kv_key], self.shapes[kv_key] = self.dtypes[k_key], [2, *self.shapes[k_key]]
            self.kv_cache_keys_map[kv_key] = (k_key, v_key)

    def forward(self) -> List[Dict[str, torch.Tensor]]:
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        out = []
        for _ in range(self.local_ranks_size):
            st: Dict[str, torch.Tensor] = {}
            for key in self.state_keys:
                val = torch.zeros(tuple(self.shapes[key]), dtype=self.dtypes[key], device=dev)
                st[key] = val
                if key in self.kv_cache_keys_map:
                    k_key, v_key = self.kv_cache_keys_map[key]
                    st[k_key], st[v_key] = val[0].view(tuple(self.shapes[k_key])), val[1].view(tuple(self.shapes[v_key]))
            out.append(st)
        return out


class NxDModelExecutor(nn.Module):
    """Thin callable around a runtime model (reference :276-291): ``executor(*tensors)`` → ``nxd_model(*tensors)``."""

    def __init__(self, nxd_model: nn.Module):
        super().__init__()
        self.nxd_model = nxd_model

    def forward(self, *inputs):
        return self.nxd_model(*inputs)
