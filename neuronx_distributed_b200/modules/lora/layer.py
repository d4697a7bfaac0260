"""LoRA adapter layers (reference ``modules/lora/layer.py``): ``y = base(x) + scaling · B(A(dropout(x)))``."""
from __future__ import annotations

import math

from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .config import LoraConfig


class LoraLayer(nn.Module):
    """Common state: base layer, low-rank pair, scaling, merge bookkeeping."""

    adapter_layer_names = ("lora_A", "lora_B", "lora_embedding_A", "lora_embedding_B")
    other_param_names = ("lora_rank", "lora_alpha", "scaling", "lora_dropout")

    def __init__(self, base_layer: nn.Module, config: Optional[LoraConfig] = None, lora_config: Optional[LoraConfig] = None, **_compat):
        config = config if config is not None else lora_config          # ``lora_config=`` is the reference's keyword
        assert config is not None, "a LoraConfig is required"
        if config.lora_rank <= 0:
            raise ValueError(f"`lora_rank` should be a positive integer value but the value passed is {config.lora_rank}")
        super().__init__()
        self.base_layer = base_layer
        self.lora_config = config
        self.r, self.scaling = config.lora_rank, config.scaling
        self.lora_rank, self.lora_alpha, self.lora_dropout = config.lora_rank, config.lora_alpha, config.lora_dropout
        self.dropout = nn.Dropout(config.lora_dropout) if config.lora_dropout > 0 else nn.Identity()
        self.merged = False
        for p in base_layer.parameters():
            p.requires_grad_(False)

    def init_lora_parameters(self, a=None, b: Optional[torch.Tensor] = None) -> None:
        """``init_lora_parameters(A, B)`` initialises the two given tensors; the reference's form
        ``init_lora_parameters("default" | "gaussian")`` (layer.py:133-151) re-initialises this layer's own adapter."""
        if a is None or isinstance(a, str):
            mode = (a or self.lora_config.init_lora_weights or "default").lower()
            assert mode in ("default", "gaussian"), f"Unknown LoRA parameters initialization with {mode}"
            saved, self.lora_config.init_lora_weights = self.lora_config.init_lora_weights, mode
            try:
                la, lb = getattr(self, "lora_A", None), getattr(self, "lora_B", None)
                if la is not None and lb is not None:
                    self.init_lora_parameters(getattr(la, "weight", la), getattr(lb, "weight", lb))
            finally:
                self.lora_config.init_lora_weights = saved
            return
        if self.lora_config.init_lora_weights == "gaussian":
            nn.init.normal_(a, std=1.0 / self.r)
        else:
            nn.init.kaiming_uniform_(a, a=math.sqrt(5))
        nn.init.zeros_(b)

    def delta_weight(self) -> torch.Tensor:
        raise NotImplementedError

    def get_delta_weight(self) -> torch.Tensor:
        """``scaling · B·A`` in the base weight's layout (reference name)."""
        return self.delta_weight()

    def get_base_layer(self) -> nn.Module:
        """Innermost wrapped layer (adapters may wrap adapters)."""
        base = self
        while hasattr(base, "base_layer"):
            base = base.base_layer
        return base

    def update_layer(self, *args, **kwargs) -> None:
        """(Re)create the low-rank pair for the current config — subclasses build it in ``__init__``; calling this
        re-initialises the adapter weights in place."""
        a = getattr(self, "lora_A", None)
        b = getattr(self, "lora_B", None)
        if a is not None and b is not None and hasattr(a, "weight"):
            self.init_lora_parameters(a.weight.data, b.weight.data)

    @staticmethod
    def transpose(weight):
        return nn.Parameter(weight.T) if isinstance(weight, nn.Parameter) else weight.T

    def __repr__(self) -> str:
        return "lora." + super().__repr__()

    def merge(self, safe_merge: bool = False) -> None:
        """Fold the adapter into the base weight.  ``safe_merge`` merges into a copy first and refuses non-finite results
        (a broken adapter must not poison the served base weights)."""
        if self.merged:
            return
        w = self.base_layer.weight
        delta = self.delta_weight().to(w.dtype)
        if safe_merge:
            merged = w.data + delta
            if not torch.isfinite(merged).all():
                raise ValueError("NaNs detected in the merged weights. The adapter seems to be broken")
            w.data.copy_(merged)
        else:
            w.data += delta
        self.merged = True

    def unmerge(self) -> None:
        if self.merged:
            self.base_layer.weight.data -= self.delta_weight().to(self.base_layer.weight.dtype)
            self.merged = False


class LoraLinear(LoraLayer):
    def __init__(self, base_layer: nn.Linear, config: Optional[LoraConfig] = None, lora_config: Optional[LoraConfig] = None, **_compat):
        config = config if config is not None else lora_config          # ``lora_config=`` is the reference's keyword
        assert config is not None, "a LoraConfig is required"
        super().__init__(base_layer, config)
        dt, dev = base_layer.weight.dtype, base_layer.weight.device
        self.lora_A = nn.Linear(base_layer.in_features, self.r, bias=False, dtype=dt, device=dev)
        self.lora_B = nn.Linear(self.r, base_layer.out_features, bias=False, dtype=dt, device=dev)
        self.init_lora_parameters(self.lora_A.weight, self.lora_B.weight)

    def delta_weight(self) -> torch.Tensor:
        return (self.lora_B.weight @ self.lora_A.weight) * self.scaling

    def forward(self, x: torch.Tensor, *a, **k):
        y = self.base_layer(x, *a, **k)
        if self.merged:
            return y
        return y + self.lora_B(self.lora_A(self.dropout(x))) * self.scaling


class LoraEmbedding(LoraLayer):
    def __init__(self, base_layer: nn.Embedding, config: Optional[LoraConfig] = None, lora_config: Optional[LoraConfig] = None, **_compat):
        config = config if config is not None else lora_config          # ``lora_config=`` is the reference's keyword
        assert config is not None, "a LoraConfig is required"
        super().__init__(base_layer, config)
        dt, dev = base_layer.weight.dtype, base_layer.weight.device
        self.lora_embedding_A = nn.Parameter(torch.empty(self.r, base_layer.num_embeddings, dtype=dt, device=dev))
        self.lora_embedding_B = nn.Parameter(torch.empty(base_layer.embedding_dim, self.r, dtype=dt, device=dev))
        nn.init.zeros_(self.lora_embedding_A)
        nn.init.normal_(self.lora_embedding_B)

    def delta_weight(self) -> torch.Tensor:
        return (self.lora_embedding_B @ self.lora_embedding_A).t() * self.scaling

    def forward(self, x: torch.Tensor):
        ids = x      # reference parameter names in the signature
        y = self.base_layer(ids)
        if self.merged:
            return y
        after_a = F.embedding(ids, self.lora_embedding_A.t())
        return y + (after_a @ self.lora_embedding_B.t()) * self.scaling


class LoraConv2d(LoraLayer):
    def __init__(self, base_layer: nn.Conv2d, config: Optional[LoraConfig] = None, lora_config: Optional[LoraConfig] = None, **_compat):
        config = config if config is not None else lora_config          # ``lora_config=`` is the reference's keyword
        assert config is not None, "a LoraConfig is required"
        super().__init__(base_layer, config)
        dt, dev = base_layer.weight.dtype, base_layer.weight.device
        self.lora_A = nn.Conv2d(base_layer.in_channels, self.r, base_layer.kernel_size, base_layer.stride,
                                base_layer.padding, bias=False, dtype=dt, device=dev)
        self.lora_B = nn.Conv2d(self.r, base_layer.out_channels, 1, 1, bias=False, dtype=dt, device=dev)
        self.init_lora_parameters(self.lora_A.weight, self.lora_B.weight)

    def delta_weight(self) -> torch.Tensor:
        a, b = self.lora_A.weight, self.lora_B.weight
        return (b.flatten(1) @ a.flatten(1)).view(self.base_layer.weight.shape) * self.scaling

    def forward(self, x: torch.Tensor):
        y = self.base_layer(x)
        if self.merged:
            return y
        return y + self.lora_B(self.lora_A(self.dropout(x))) * self.scaling
