"""Dispatch of TP linear layers onto the fused GEMM+collective kernels.

``dispatch(x, weight, in_mode, out_mode, seq_dim, group)`` returns an object with
``forward``/``backward`` when the sm_100a fused path applies, else ``None`` (the caller then
runs collectives + matmul through the library path).  Filled in by ``_fused_impl`` once the
extension exposes the kernels; selection can be forced off with ``NXD_TP_BACKEND=nccl``.
"""
from __future__ import annotations

import os

import torch


_BACKEND = os.environ.get("NXD_TP_BACKEND", "fused")


def set_backend(name: str) -> None:
    """``"fused"`` (hand-written kernels; default on CUDA) or ``"nccl"`` (library baseline)."""
    global _BACKEND
    assert name in ("fused", "nccl")
    _BACKEND = name


def get_backend() -> str:
    return _BACKEND


def dispatch(x: torch.Tensor, weight: torch.Tensor, in_mode: str, out_mode: str, seq_dim: int, group):
    if _BACKEND != "fused" or not x.is_cuda:
        return None
    try:
        from . import _fused_impl
    except ImportError:
        return None
    return _fused_impl.select(x, weight, in_mode, out_mode, seq_dim, group)
