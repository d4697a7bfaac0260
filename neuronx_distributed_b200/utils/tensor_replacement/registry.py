"""Run-time arguments of tensor replacement (reference ``utils/tensor_replacement/registry.py:4-35``)."""
from __future__ import annotations

from typing import Dict, List

import torch


class RuntimeRegister:
    """Static registry: ``module_superset`` is the ordered list of replaceable modules fixed when the model is prepared;
    every step supplies one (tensor, mask) pair per entry, in that order."""

    _tr_runtime_list: Dict[str, torch.Tensor] = {}
    _tr_mask_list: Dict[str, torch.Tensor] = {}
    module_superset: List[str] = []

    @classmethod
    def register_runtime_args(cls, tr_args: List[torch.Tensor], mask_args: List[torch.Tensor]) -> None:
        n = len(cls.module_superset)
        if len(tr_args) != n or len(mask_args) != n:
            raise ValueError(f"[tensor replacement] expected {n} tensors and {n} masks, got {len(tr_args)} and {len(mask_args)}")
        cls.clear_runtime_args()
        for name, t, m in zip(cls.module_superset, tr_args, mask_args):
            if not isinstance(t, torch.Tensor) or not isinstance(m, torch.Tensor):
                raise TypeError(f"[tensor replacement] tensor and mask for '{name}' must be torch.Tensor, got {type(t)} / {type(m)}")
            cls._tr_runtime_list[name], cls._tr_mask_list[name] = t, m

    @classmethod
    def clear_runtime_args(cls) -> None:
        cls._tr_runtime_list, cls._tr_mask_list = {}, {}
