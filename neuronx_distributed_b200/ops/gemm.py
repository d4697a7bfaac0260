"""bf16 GEMM on tcgen05 tensor cores (kernel: ``csrc/gemm_sm100.cu``).

``matmul(a, b, trans_a, trans_b)`` computes ``op(a) @ op(b)`` for row-major 2-D bf16 tensors with
fp32 accumulation in TMEM.  The three layouts training needs map to UMMA operand majors without
any transposition copies:

    fwd   : x[M,K] @ W[N,K]^T     → A K-major,  B K-major     ("nt")
    dgrad : g[M,K] @ W[K,N]       → A K-major,  B MN-major    ("nn")
    wgrad : g[K,M]^T @ x[K,N]     → A MN-major, B MN-major    ("tn")
"""
from __future__ import annotations

import os

import torch

from . import _ext

_MIN_DIM = 64
_USE_2CTA = os.environ.get("NXD_GEMM_2CTA", "1") == "1"   # CTA-pair kernel (cta_group::2)


def fused_wgrad_enabled() -> bool:
    return os.environ.get("NXD_DISABLE_FUSED_WGRAD", "0") != "1" and os.environ.get("NXD_DISABLE_TCGEN05_GEMM", "0") != "1"


def _eligible(a: torch.Tensor, b: torch.Tensor, M: int, N: int, K: int) -> bool:
    if os.environ.get("NXD_DISABLE_TCGEN05_GEMM", "0") == "1":
        return False
    if not (a.is_cuda and b.is_cuda and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16):
        return False
    if not (a.is_contiguous() and b.is_contiguous()) or a.data_ptr() % 16 or b.data_ptr() % 16:
        return False
    if not _ext.use_cuda(a, b) or not hasattr(_ext.ext(), "gemm_bf16"):
        return False
    # TMA needs 16-byte aligned rows; kernel handles ragged M/N/K tails via TMA zero fill
    return M >= 1 and N % 8 == 0 and K % 8 == 0 and M % 8 == 0


def matmul(a: torch.Tensor, b: torch.Tensor, trans_a: bool = False, trans_b: bool = False,
           out: torch.Tensor | None = None, accumulate: bool = False, out_dtype=None) -> torch.Tensor:
    """``op(a) @ op(b)`` for 2-D tensors. ``accumulate`` adds into ``out`` (fp32 or bf16)."""
    M = a.shape[1] if trans_a else a.shape[0]
    K = a.shape[0] if trans_a else a.shape[1]
    N = b.shape[0] if trans_b else b.shape[1]
    out_dtype = out_dtype or (out.dtype if out is not None else a.dtype)
    if _eligible(a, b, M, N, K) and out_dtype in (torch.bfloat16, torch.float32):
        if out is None:
            out = torch.empty(M, N, dtype=out_dtype, device=a.device)
            accumulate = False
        _ext.count_launch()
        if _USE_2CTA and M >= 256 and hasattr(_ext.ext(), "gemm_bf16_2cta"):
            _ext.ext().gemm_bf16_2cta(a, b, out, bool(trans_a), bool(trans_b), bool(accumulate))
        else:
            _ext.ext().gemm_bf16(a, b, out, bool(trans_a), bool(trans_b), bool(accumulate))
        return out
    if a.is_cuda and a.dtype == torch.bfloat16 and not os.environ.get("NXD_DISABLE_TCGEN05_GEMM") \
            and (not a.is_contiguous() or not b.is_contiguous()):
        # make operands dense once and retry the tensor-core kernel
        return matmul(a.contiguous(), b.contiguous(), trans_a, trans_b, out, accumulate, out_dtype)
    aa = a.t() if trans_a else a
    bb = b.t() if trans_b else b
    res = torch.matmul(aa, bb)
    if out is not None:
        if accumulate:
            out.add_(res.to(out.dtype))
        else:
            out.copy_(res)
        return out
    return res.to(out_dtype)


def linear_nt(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """``x[..., K] @ w[N, K]^T``."""
    x2 = x.reshape(-1, x.shape[-1])
    return matmul(x2, w, False, True).view(*x.shape[:-1], w.shape[0])
