"""LayerNorm whose affine params are tagged for sequence-parallel gradient all-reduce
(reference ``parallel_layers/layer_norm.py:17-48``)."""
from __future__ import annotations

import torch
from torch import nn


def _set_sequence_parallel_enabled(param: torch.Tensor, enabled: bool) -> None:
    setattr(param, "sequence_parallel_enabled", enabled)


class LayerNorm(nn.LayerNorm):
    def __init__(self, normalized_shape, eps: float = 1e-5, elementwise_affine: bool = True,
                 sequence_parallel_enabled: bool = False, dtype: torch.dtype = torch.float32, device=None,
                 bias: bool = True):
        super().__init__(normalized_shape, eps=eps, elementwise_affine=elementwise_affine, bias=bias,
                         device=device, dtype=dtype)
        self.sequence_parallel_enabled = sequence_parallel_enabled
        if elementwise_affine:
            _set_sequence_parallel_enabled(self.weight, sequence_parallel_enabled)
            if self.bias is not None:
                _set_sequence_parallel_enabled(self.bias, sequence_parallel_enabled)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        # compute in fp32, return in the caller's dtype (the reference forgets to assign the
        # cast result, layer_norm.py:47; the intent is implemented here)
        x = input      # reference parameter names in the signature
        dt = x.dtype
        w = self.weight.float() if self.weight is not None else None
        b = self.bias.float() if self.bias is not None else None
        return torch.nn.functional.layer_norm(x.float(), self.normalized_shape, w, b, self.eps).to(dt)
