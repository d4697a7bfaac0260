"""``LoraModel`` — inject adapters into a model, save / load them, merge / unmerge
(reference ``modules/lora/model.py:74-746``).  Adapter checkpoints use HF-PEFT-compatible key names
(``base_model.model.<module>.lora_A.weight``) plus ``adapter_config.json`` content under ``"lora_config"``."""
from __future__ import annotations

import os
import re
from typing import Any, Dict, List, Optional

import torch
from torch import nn

from ...parallel_layers.layers import ColumnParallelLinear, RowParallelLinear
from ..qkv_linear import GQAQKVColumnParallelLinear
from .config import LoraConfig
from .layer import LoraConv2d, LoraEmbedding, LoraLayer, LoraLinear
from .tp_layer import LoraGQAQKVParallelLinear, LoraParallelLinear

_DEFAULT_TARGETS = ["q_proj", "k_proj", "v_proj", "o_proj", "qkv_proj", "gate_up_proj", "down_proj", "up_proj", "gate_proj"]


def _wrap(module: nn.Module, cfg: LoraConfig) -> Optional[nn.Module]:
    if isinstance(module, GQAQKVColumnParallelLinear):
        return LoraGQAQKVParallelLinear(module, cfg)
    if isinstance(module, (ColumnParallelLinear, RowParallelLinear)):
        return LoraParallelLinear(module, cfg)
    if isinstance(module, nn.Linear):
        return LoraLinear(module, cfg)
    if isinstance(module, nn.Embedding):
        return LoraEmbedding(module, cfg)
    if isinstance(module, nn.Conv2d):
        return LoraConv2d(module, cfg)
    return None


class LoraModel(nn.Module):
    def __init__(self, module: nn.Module, config: LoraConfig):
        super().__init__()
        self.module = module
        self.lora_config = config
        self.modules_to_save: List[str] = []
        if config.enable_lora:
            self.inject_adapter()
        if config.load_lora_from_ckpt and config.lora_save_path:
            self.load_lora(config.lora_save_path, config.lora_load_tag)

    # ------------------------------------------------------------------ injection
    def _is_target(self, name: str) -> bool:
        t = self.lora_config.target_modules
        if t is None:
            t = _DEFAULT_TARGETS
        if isinstance(t, str):
            return re.fullmatch(t, name) is not None
        return any(name == x or name.endswith("." + x) for x in t)

    def inject_adapter(self) -> None:
        replaced = 0
        for name, child in list(self.module.named_modules()):
            if not name or isinstance(child, LoraLayer) or not self._is_target(name):
                continue
            wrapped = _wrap(child, self.lora_config)
            if wrapped is None:
                continue
            parent_name, _, leaf = name.rpartition(".")
            parent = self.module.get_submodule(parent_name) if parent_name else self.module
            setattr(parent, leaf, wrapped)
            replaced += 1
        if replaced == 0:
            raise ValueError(f"no module matched LoRA target_modules={self.lora_config.target_modules}")
        self.mark_only_lora_as_trainable()

    def mark_only_lora_as_trainable(self) -> None:
        bias = self.lora_config.bias
        for n, p in self.module.named_parameters():
            is_lora = "lora_" in n
            p.requires_grad_(is_lora or (bias == "all" and n.endswith("bias")))
        if bias == "lora_only":
            for m in self.module.modules():
                if isinstance(m, LoraLayer) and getattr(m.base_layer, "bias", None) is not None:
                    m.base_layer.bias.requires_grad_(True)

    # ------------------------------------------------------------------ forward / passthrough
    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def __getattr__(self, name: str) -> Any:
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("module"), name)

    # ------------------------------------------------------------------ merge
    def merge_lora(self) -> None:
        for m in self.module.modules():
            if isinstance(m, LoraLayer) and not isinstance(m, LoraGQAQKVParallelLinear):
                m.merge()

    def unmerge_lora(self) -> None:
        for m in self.module.modules():
            if isinstance(m, LoraLayer) and not isinstance(m, LoraGQAQKVParallelLinear):
                m.unmerge()

    # ------------------------------------------------------------------ state dicts
    def lora_state_dict(self) -> Dict[str, torch.Tensor]:
        sd = {}
        for k, v in self.module.state_dict().items():
            if "lora_" in k:
                sd["base_model.model." + k] = v
        return sd

    def base_state_dict(self) -> Dict[str, torch.Tensor]:
        return {k.replace(".base_layer", ""): v for k, v in self.module.state_dict().items() if "lora_" not in k}

    def save_lora(self, path: Optional[str] = None, tag: Optional[str] = None) -> str:
        from ...parallel_layers import parallel_state as ps

        path = path or self.lora_config.lora_save_path
        assert path is not None
        d = os.path.join(path, tag) if tag else path
        os.makedirs(d, exist_ok=True)
        tp = ps.get_tensor_model_parallel_rank() if ps.model_parallel_is_initialized() else 0
        pp = ps.get_pipeline_model_parallel_rank() if ps.model_parallel_is_initialized() else 0
        f = os.path.join(d, f"adapter_tp_rank_{tp:02d}_pp_rank_{pp:02d}.pt")
        torch.save({"lora_config": self.lora_config.to_dict(), "state_dict": {k: v.cpu() for k, v in self.lora_state_dict().items()}}, f)
        return f

    def load_lora(self, path: str, tag: Optional[str] = None) -> None:
        from ...parallel_layers import parallel_state as ps

        d = os.path.join(path, tag) if tag else path
        tp = ps.get_tensor_model_parallel_rank() if ps.model_parallel_is_initialized() else 0
        pp = ps.get_pipeline_model_parallel_rank() if ps.model_parallel_is_initialized() else 0
        ck = torch.load(os.path.join(d, f"adapter_tp_rank_{tp:02d}_pp_rank_{pp:02d}.pt"), map_location="cpu", weights_only=False)
        sd = {k[len("base_model.model."):]: v for k, v in ck["state_dict"].items()}
        missing, unexpected = self.module.load_state_dict(sd, strict=False)
        bad = [k for k in unexpected]
        if bad:
            raise RuntimeError(f"unexpected adapter keys: {bad}")

    def state_dict(self, *a, **k):
        if self.lora_config.save_lora_base:
            return self.module.state_dict(*a, **k)
        return self.lora_state_dict()

    def load_state_dict(self, sd, strict: bool = True):
        sd = {(k[len("base_model.model."):] if k.startswith("base_model.model.") else k): v for k, v in sd.items()}
        return self.module.load_state_dict(sd, strict=False)


def get_lora_model(model: nn.Module, lora_config: LoraConfig) -> LoraModel:
    return LoraModel(model, lora_config)
