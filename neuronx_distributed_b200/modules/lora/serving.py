"""Multi-adapter LoRA serving (role of reference ``examples/inference/modules/lora_serving/*``: ``LoraServingConfig``,
``MultiLoraLinear`` / ``MultiLoraModule*``, the serving ``LoraModel``).

``max_loras`` adapters live on the device at once as stacked, rank-padded weights ``A [L, r_max, in]`` and
``B [L, out, r_max]``.  Every request of a batch selects its adapter with ``adapter_ids [B]`` (``-1`` = base model only), so
one decode step serves a mixed batch: ``y_b = W x_b + s_{a_b} · B_{a_b} (A_{a_b} x_b)`` — two batched GEMMs over gathered
adapter slices (the reference applies ONE adapter per forward call).  Adapter ids are data, not control flow: a captured
CUDA graph stays valid when the mix of adapters changes, and loading a new adapter into a slot is an in-place copy.

Tensor parallelism follows ``tp_layer.py``: column-parallel base → ``A`` replicated, ``B`` sharded on ``out``;
row-parallel base → ``A`` sharded on ``in`` (the ``[…, r]`` partial product is all-reduced — a few KB), ``B`` replicated.
Full (un-sharded) HF-PEFT adapters are sharded on load.
"""
from __future__ import annotations

import contextlib
import json
import os
import re
from typing import Any, Dict, List, Optional, Sequence, Union

import torch
from torch import nn

from ...parallel_layers import mappings
from ...parallel_layers.layers import ColumnParallelLinear, RowParallelLinear


class LoraServingConfig:
    def __init__(self, max_loras: int = 1, max_lora_rank: int = 16, max_loras_on_cpu: int = 2, lora_dtype=torch.float32,
                 target_modules: Optional[List[str]] = None, lora_bias: str = "none", lora_ckpt_paths: Optional[List[str]] = None):
        self.max_loras, self.max_lora_rank, self.max_loras_on_cpu = max_loras, max_lora_rank, max_loras_on_cpu
        self.lora_dtype, self.target_modules, self.lora_bias = lora_dtype, target_modules, lora_bias
        self.lora_ckpt_paths = lora_ckpt_paths

    def to_json_string(self) -> str:
        d = dict(self.__dict__)
        d["lora_dtype"] = str(d["lora_dtype"]).replace("torch.", "")
        return json.dumps(d, indent=2, sort_keys=True)

    def to_json_file(self, json_file: Union[str, os.PathLike]) -> None:
        with open(json_file, "w", encoding="utf-8") as f:
            f.write(self.to_json_string() + "\n")

    @classmethod
    def from_json_string(cls, json_string: str, **kwargs) -> "LoraServingConfig":
        d = {**json.loads(json_string), **kwargs}
        if isinstance(d.get("lora_dtype"), str):
            d["lora_dtype"] = getattr(torch, d["lora_dtype"])
        return cls(**d)

    @classmethod
    def from_json_file(cls, json_file: Union[str, os.PathLike], **kwargs) -> Optional["LoraServingConfig"]:
        if not os.path.exists(json_file):
            return None
        with open(json_file, "r", encoding="utf-8") as f:
            return cls.from_json_string(f.read(), **kwargs)


class _AdapterSelection:
    """Shared, mutable holder of the current batch's adapter ids (one per serving model)."""

    def __init__(self):
        self.ids: Optional[torch.Tensor] = None


class MultiLoraLinear(nn.Module):
    """A (parallel) linear with ``max_loras`` stacked adapters."""

    def __init__(self, base_layer: nn.Module, config: LoraServingConfig, selection: _AdapterSelection):
        super().__init__()
        self.base_layer, self.config = base_layer, config
        object.__setattr__(self, "_sel", selection)
        w = base_layer.weight
        L, r = config.max_loras, config.max_lora_rank
        self.is_row = isinstance(base_layer, RowParallelLinear)
        self.is_col = isinstance(base_layer, ColumnParallelLinear)
        in_local = w.shape[1]
        out_local = w.shape[0]
        dt = config.lora_dtype if w.dtype == torch.float32 else w.dtype
        self.lora_A = nn.Parameter(torch.zeros(L, r, in_local, dtype=dt, device=w.device), requires_grad=False)
        self.lora_B = nn.Parameter(torch.zeros(L, out_local, r, dtype=dt, device=w.device), requires_grad=False)
        self.register_buffer("scaling", torch.zeros(L, dtype=torch.float32, device=w.device))
        for p in base_layer.parameters():
            p.requires_grad_(False)

    # ---- loading -----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def load_adapter(self, slot: int, a: torch.Tensor, b: torch.Tensor, alpha: float, rank: Optional[int] = None,
                     use_rslora: bool = False) -> None:
        """``a [r, in]`` / ``b [out, r]`` full (un-sharded) or already local; rank ``r ≤ max_lora_rank`` is zero-padded."""
        assert 0 <= slot < self.config.max_loras, f"slot {slot} out of range (max_loras={self.config.max_loras})"
        r = a.shape[0] if rank is None else rank
        assert r <= self.config.max_lora_rank, f"adapter rank {r} exceeds max_lora_rank {self.config.max_lora_rank}"
        base = self.base_layer
        if self.is_row and a.shape[1] != self.lora_A.shape[2]:          # shard A along in
            tp, rk = base.tensor_model_parallel_size, base._tp_rank
            a = a.chunk(tp, dim=1)[rk]
        if self.is_col and b.shape[0] != self.lora_B.shape[1]:          # shard B along out like the base weight
            from ...parallel_layers.utils import create_local_weight

            tp, rk = base.tensor_model_parallel_size, base._tp_rank
            b = create_local_weight(b, 0, b.shape[0] // tp, base.stride, rank=rk, world_size=tp)
        self.lora_A[slot].zero_()
        self.lora_B[slot].zero_()
        self.lora_A[slot, :r].copy_(a.to(self.lora_A.dtype))
        self.lora_B[slot, :, :r].copy_(b.to(self.lora_B.dtype))
        self.scaling[slot] = alpha / (r ** 0.5 if use_rslora else r)

    @torch.no_grad()
    def unload_adapter(self, slot: int) -> None:
        self.lora_A[slot].zero_()
        self.lora_B[slot].zero_()
        self.scaling[slot] = 0.0

    # ---- forward -----------------------------------------------------------------------------------------------------
    def _delta(self, x: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
        """``x [B, …, in]``, ``ids [B]`` → adapter output ``[B, …, out_local]`` (zeros where ``ids < 0``)."""
        valid = (ids >= 0)
        idx = ids.clamp(min=0)
        A, Bm = self.lora_A[idx], self.lora_B[idx]                       # [B, r, in], [B, out, r]
        s = (self.scaling[idx] * valid.to(self.scaling.dtype))
        shape = x.shape
        x3 = x.reshape(shape[0], -1, shape[-1]).to(A.dtype)
        h = torch.bmm(x3, A.transpose(1, 2))                             # [B, T, r]
        if self.is_row and self.base_layer.tensor_model_parallel_size > 1:
            h = mappings.reduce_from_tensor_model_parallel_region(h, self.base_layer.tensor_parallel_group)
        y = torch.bmm(h, Bm.transpose(1, 2)) * s.view(-1, 1, 1).to(h.dtype)
        return y.reshape(*shape[:-1], Bm.shape[1])

    def forward(self, x: torch.Tensor, *args, adapter_ids: Optional[torch.Tensor] = None, **kwargs):
        y = self.base_layer(x, *args, **kwargs)
        ids = adapter_ids if adapter_ids is not None else self._sel.ids
        if ids is None:
            return y
        bias = None
        if isinstance(y, tuple):
            y, bias = y
        if not isinstance(ids, torch.Tensor):
            ids = torch.full((x.shape[0],), int(ids), device=x.device, dtype=torch.long)
        base = self.base_layer
        seq_first = getattr(base, "sequence_parallel_enabled", False) or (x.dim() == 3 and x.shape[0] != ids.shape[0]
                                                                           and x.shape[1] == ids.shape[0])
        xin = x
        if getattr(base, "sequence_parallel_enabled", False) and self.is_col:
            xin = mappings.gather_from_sequence_parallel_region(x, base.sequence_dimension, True, base.tensor_parallel_group)
        if seq_first:                                                    # [S, B, H] activations: batch is dim 1
            d = self._delta(xin.transpose(0, 1), ids.to(x.device)).transpose(0, 1)
        else:
            d = self._delta(xin, ids.to(x.device))
        if self.is_row and getattr(base, "sequence_parallel_enabled", False):
            d = mappings.scatter_to_sequence_parallel_region(d, base.sequence_dimension, base.tensor_parallel_group)
        if self.is_col and getattr(base, "gather_output", False):
            d = mappings.gather_from_tensor_model_parallel_region(d, base.tensor_parallel_group)
        y = y + d.to(y.dtype)
        return (y, bias) if bias is not None else y


class LoraServingModel(nn.Module):
    """Wrap a model for multi-adapter serving: targets become :class:`MultiLoraLinear`; ``load_adapter`` fills a slot from
    an HF-PEFT style state dict; ``adapter_ids`` are set per batch with :meth:`set_adapter_ids` / the ``adapters`` context."""

    _DEFAULT_TARGETS = ["q_proj", "k_proj", "v_proj", "o_proj", "gate_up_proj", "gate_proj", "up_proj", "down_proj"]

    def __init__(self, module: nn.Module, config: LoraServingConfig):
        super().__init__()
        self.module, self.lora_config = module, config
        object.__setattr__(self, "_sel", _AdapterSelection())
        self.slots: Dict[int, Dict[str, Any]] = {}
        self.inject_adapter()
        for i, path in enumerate(config.lora_ckpt_paths or []):
            self.load_adapter(i, path)

    def _is_target(self, name: str) -> bool:
        t = self.lora_config.target_modules or self._DEFAULT_TARGETS
        if isinstance(t, str):
            return re.fullmatch(t, name) is not None
        return any(name == x or name.endswith("." + x) for x in t)

    def get_leave_module_names(self) -> List[str]:
        return [n for n, m in self.module.named_modules() if n and not list(m.children())]

    def inject_adapter(self) -> None:
        n = 0
        for name, child in list(self.module.named_modules()):
            if not name or isinstance(child, MultiLoraLinear) or not self._is_target(name):
                continue
            if not isinstance(child, (nn.Linear, ColumnParallelLinear, RowParallelLinear)):
                continue
            parent_name, _, leaf = name.rpartition(".")
            parent = self.module.get_submodule(parent_name) if parent_name else self.module
            if isinstance(parent, MultiLoraLinear):
                continue
            setattr(parent, leaf, MultiLoraLinear(child, self.lora_config, self._sel))
            n += 1
        if n == 0:
            raise ValueError(f"no module matched target_modules={self.lora_config.target_modules}")

    def lora_layers(self) -> Dict[str, MultiLoraLinear]:
        return {n: m for n, m in self.module.named_modules() if isinstance(m, MultiLoraLinear)}

    def load_adapter(self, slot: int, state: Union[str, Dict[str, torch.Tensor]], alpha: Optional[float] = None,
                     rank: Optional[int] = None, use_rslora: bool = False) -> None:
        """``state``: path of an adapter file (``torch.save`` of a state dict, optionally with ``lora_config`` inside, or a
        ``.safetensors``) or the state dict itself; keys ``[base_model.model.]<module>.lora_A.weight`` / ``lora_B.weight``."""
        if isinstance(state, str):
            if state.endswith(".safetensors"):
                from ...utils.safetensors_utils import load_state_dict_safetensors

                state = load_state_dict_safetensors(state)
            else:
                state = torch.load(state, map_location="cpu", weights_only=False)
        cfg = state.get("lora_config") if isinstance(state, dict) else None
        tensors = state.get("state_dict", state)
        if cfg:
            alpha = cfg.get("lora_alpha", alpha) if alpha is None else alpha
            rank = cfg.get("lora_rank", cfg.get("r", rank)) if rank is None else rank
            use_rslora = cfg.get("use_rslora", use_rslora)
        sd = {(k[len("base_model.model."):] if k.startswith("base_model.model.") else k): v
              for k, v in tensors.items() if isinstance(v, torch.Tensor)}
        loaded = 0
        for name, layer in self.lora_layers().items():
            ka, kb = f"{name}.lora_A.weight", f"{name}.lora_B.weight"
            if ka in sd and kb in sd:
                r = sd[ka].shape[0] if rank is None else rank
                layer.load_adapter(slot, sd[ka], sd[kb], alpha if alpha is not None else float(r), r, use_rslora)
                loaded += 1
            else:
                layer.unload_adapter(slot)
        if loaded == 0:
            raise ValueError("the adapter has no tensors for any LoRA target of this model")
        self.slots[slot] = {"alpha": alpha, "rank": rank, "modules": loaded}

    def unload_adapter(self, slot: int) -> None:
        for layer in self.lora_layers().values():
            layer.unload_adapter(slot)
        self.slots.pop(slot, None)

    def set_adapter_ids(self, adapter_ids: Optional[Union[torch.Tensor, Sequence[int], int]]) -> None:
        if adapter_ids is None or isinstance(adapter_ids, torch.Tensor):
            self._sel.ids = adapter_ids
        elif isinstance(adapter_ids, int):
            self._sel.ids = adapter_ids                                  # broadcast to the batch in the layer
        else:
            self._sel.ids = torch.as_tensor(list(adapter_ids), dtype=torch.long)

    @contextlib.contextmanager
    def adapters(self, adapter_ids):
        prev = self._sel.ids
        self.set_adapter_ids(adapter_ids)
        try:
            yield self
        finally:
            self._sel.ids = prev

    def forward(self, *args, adapter_ids=None, **kwargs):
        if adapter_ids is None:
            return self.module(*args, **kwargs)
        with self.adapters(adapter_ids):
            return self.module(*args, **kwargs)

    def generate(self, *args, adapter_ids=None, **kwargs):
        with self.adapters(adapter_ids):
            return self.module.generate(*args, **kwargs)

    def get_base_model(self) -> nn.Module:
        return self.module

    def __getattr__(self, name: str) -> Any:
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("module"), name)


def wrap_model_with_lora(model: nn.Module, config: LoraServingConfig) -> LoraServingModel:
    return LoraServingModel(model, config)
