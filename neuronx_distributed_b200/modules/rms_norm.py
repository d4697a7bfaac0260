"""RMSNorm with fp32 statistics and a sequence-parallel-tagged weight
(reference ``modules/rms_norm.py:10-72``).  Forward/backward run the sm_100a kernels in
``csrc/elementwise.cu`` via :func:`ops.norm.rms_norm`."""
from __future__ import annotations

import torch
from torch import nn

from ..ops.norm import rms_norm


def manual_rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """Plain-PyTorch form (kept for numerics tests)."""
    dt = x.dtype
    xf = x.float()
    return (weight.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps))).to(dt)


class RMSNorm(nn.Module):
    def __init__(self, hidden_size: int, eps: float = 1e-6, sequence_parallel_enabled: bool = False,
                 dtype: torch.dtype = torch.float32, device=None):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size, dtype=dtype, device=device))
        self.variance_epsilon = eps
        self.sequence_parallel_enabled = sequence_parallel_enabled
        setattr(self.weight, "sequence_parallel_enabled", sequence_parallel_enabled)

    def reset_parameters(self) -> None:
        """Re-initialise after meta-device materialisation (``reinit_model`` calls it): unit gain."""
        with torch.no_grad():
            self.weight.fill_(1.0)
        setattr(self.weight, "sequence_parallel_enabled", self.sequence_parallel_enabled)

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        return rms_norm(hidden_states, self.weight, self.variance_epsilon)

    def extra_repr(self) -> str:
        return f"{tuple(self.weight.shape)}, eps={self.variance_epsilon}"
