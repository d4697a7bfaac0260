"""Storage for captured tensors (reference ``utils/tensor_capture/registry.py:6-262``).

Two kinds of entries: *module* tensors (inputs / outputs of the monitored modules, keyed ``<module>.outputs[.i]`` /
``<module>.inputs[.i]``) and *manual* tensors registered from model code with ``register_tensor`` (keyed
``manual_<name>``, de-duplicated with a running suffix, capped by ``max_tensors``).  Entries are detached clones, so they
stay valid after the next step — under CUDA-graph capture the clone is part of the graph and is refreshed by every replay.
"""
from __future__ import annotations

from collections import OrderedDict, defaultdict
from typing import Any, Dict, List, Optional

import torch


class CapturedModelInfo:
    def __init__(self, modules_to_capture: List[str], max_tensors: Optional[int] = None, capture_inputs: bool = False):
        self.modules_to_capture = list(modules_to_capture)
        self.max_tensors = max_tensors
        self.capture_inputs = capture_inputs
        self.hooks: List[Any] = []
        self.module_tensors: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        self.manual_tensors: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        self.manual_tensors_keys: Dict[str, int] = defaultdict(int)


class TensorRegistry:
    """Process-wide singleton (``TensorRegistry.get_instance()``)."""

    _instance: Optional["TensorRegistry"] = None

    @classmethod
    def get_instance(cls) -> "TensorRegistry":
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def __init__(self):
        self.enabled = False
        self.model_info = CapturedModelInfo([], 10, False)

    def clear(self) -> None:
        self.remove_hooks()
        self.model_info = CapturedModelInfo([], 10, False)

    def reset_tensors(self) -> None:
        """Drop the captured values but keep the configuration and hooks (between steps)."""
        self.model_info.module_tensors.clear()
        self.model_info.manual_tensors.clear()
        self.model_info.manual_tensors_keys.clear()

    def configure(self, enabled: bool = False, modules=None, max_tensors: Optional[int] = None, capture_inputs: bool = False) -> None:
        self.enabled = enabled
        self.model_info = CapturedModelInfo(list(modules or []), max_tensors, capture_inputs)

    def _is_monitored(self, name) -> bool:
        if not isinstance(name, str):
            return False
        return any(name == m or name.startswith(m + ".") for m in self.model_info.modules_to_capture)

    def register_tensor(self, name, tensor: torch.Tensor, neu_module=None) -> None:
        if not self.enabled:
            return
        info = self.model_info
        if self._is_monitored(name):
            t = tensor.detach().clone()
            if neu_module is not None and "lm_head" in name:
                # logits of a padded, vocabulary-parallel head: strip the padding so captures compare against HF logits
                assert not getattr(neu_module, "sequence_parallel_enabled", False), \
                    "Sequence parallel must be disabled for lm_head tensor capture to gather logits from all ranks"
                if getattr(neu_module, "pad", False) and getattr(neu_module, "pad_size", 0) > 0 \
                        and not getattr(neu_module, "keep_padded_output", False) and t.shape[-1] == neu_module.output_size:
                    t = t.narrow(-1, 0, neu_module.output_size - neu_module.pad_size)
            info.module_tensors[name] = t
            return
        if info.max_tensors is None or len(info.manual_tensors) >= info.max_tensors:
            return                                        # no budget for manual tensors (None = module tensors only)
        if isinstance(name, str):
            base = f"manual_{name}"
            key = base if base not in info.manual_tensors else f"{base}_{info.manual_tensors_keys[base]}"
            info.manual_tensors_keys[base] += 1
        else:
            key = f"manual_tensor_{len(info.manual_tensors)}"
        info.manual_tensors[key] = tensor.detach().clone()

    def get_captured_tensors_dict(self) -> "OrderedDict[str, torch.Tensor]":
        """Module tensors in ``modules_to_capture`` order (inputs before outputs), then the manual tensors."""
        info, out = self.model_info, OrderedDict()
        for m in info.modules_to_capture:
            if info.capture_inputs:
                for k in sorted(k for k in info.module_tensors if k.startswith(f"{m}.inputs")):
                    out[k] = info.module_tensors[k]
            for k in sorted(k for k in info.module_tensors if k.startswith(f"{m}.outputs") or k == m):
                out[k] = info.module_tensors[k]
        out.update(self.get_manual_tensors())
        return out

    def _add_moe_tensors_to_manual_registry(self, manual: Dict[str, torch.Tensor]) -> None:
        """Stack the per-layer ``…moe_auto…`` entries (expert indices the MoE layers register automatically) into one
        ``auto_moe_stats.expert_index`` tensor."""
        keys = [k for k in manual if "moe_auto" in k]
        if keys:
            manual["auto_moe_stats.expert_index"] = torch.stack([manual.pop(k) for k in keys])

    def get_module_tensors(self):
        return self.model_info.module_tensors

    def get_manual_tensors(self):
        self._add_moe_tensors_to_manual_registry(self.model_info.manual_tensors)
        return self.model_info.manual_tensors

    def get_manual_tensor_count(self) -> int:
        return len(self.model_info.manual_tensors)

    def get_monitored_tensor_count(self) -> int:
        return len(self.model_info.module_tensors)

    def get_total_tensor_count(self) -> int:
        return self.get_monitored_tensor_count() + self.get_manual_tensor_count()

    def remove_hooks(self) -> None:
        for h in self.model_info.hooks:
            h.remove()
        self.model_info.hooks = []
