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
    if (M <= 8 and not trans_a and trans_b and out is None and out_dtype == torch.bfloat16 and a.is_cuda
            and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and K % 8 == 0 and a.is_contiguous() and b.is_contiguous()
            and _ext.ext() is not None and hasattr(_ext.ext(), "gemv") and os.environ.get("NXD_DISABLE_TCGEN05_GEMM", "0") != "1"):
        # decode-time GEMV: a bandwidth problem (every weight byte read once), not a tensor-core one — csrc/decode.cu
        _ext.count_launch()
        return _ext.ext().gemv(a, b, None)
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


# ---------------------------------------------------------------------------------------------------------------------
# Grouped (MoE blockwise) GEMM: rows of ``x`` come in ``block_size``-row blocks, block i is multiplied by the weights of
# expert ``block_to_expert[i]``.  CUDA: MODE 3/4 of ``csrc/gemm_sm100.cu`` (the per-block expert only shifts the TMA
# coordinate of the weight tile — no per-block weight copies, no host sync); CPU / odd shapes: plain loop.
# Role of the reference's blockwise NKI kernels (modules/moe/blockwise.py:180-468, 1037-1127).
def _grouped_ok(x: torch.Tensor, w: torch.Tensor, block_size: int) -> bool:
    e = _ext.ext()
    return (x.is_cuda and e is not None and hasattr(e, "grouped_gemm") and x.dtype == torch.bfloat16
            and w.dtype == torch.bfloat16 and block_size % 128 == 0 and w.shape[1] % 64 == 0 and w.shape[2] % 64 == 0
            and os.environ.get("NXD_DISABLE_GROUPED_GEMM", "0") != "1")


def _grouped_ref(x, w, b2e, block_size):
    nb = b2e.numel()
    xb = x.view(nb, block_size, x.shape[-1])
    return torch.einsum("bth,bhi->bti", xb, w[b2e.long()]).reshape(nb * block_size, w.shape[-1])


class _GroupedMatmul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b2e, seg_first_block, block_size):
        ctx.save_for_backward(x, w, b2e, seg_first_block)
        ctx.block_size = block_size
        _ext.count_launch()
        return _ext.ext().grouped_gemm(x.contiguous(), w.contiguous(), b2e, block_size, False)

    @staticmethod
    def backward(ctx, gy):
        x, w, b2e, seg = ctx.saved_tensors
        gy = gy.contiguous()
        gx = gw = None
        if ctx.needs_input_grad[0]:
            _ext.count_launch()
            gx = _ext.ext().grouped_gemm(gy, w.contiguous(), b2e, ctx.block_size, True)       # dy · W[e]ᵀ
        if ctx.needs_input_grad[1]:
            _ext.count_launch()
            mg = getattr(w, "main_grad", None)
            if mg is not None and fused_wgrad_enabled() and mg.dtype == torch.float32 and mg.is_contiguous():
                fresh = getattr(w, "main_grad_fresh", False)
                _ext.ext().grouped_wgrad(x.contiguous(), gy, mg.view_as(w), seg, ctx.block_size, not fresh)
                w.main_grad_fresh = False
                cb = getattr(w, "_nxd_grad_ready", None)
                if cb is not None:
                    cb(w)
            else:
                gw = torch.empty_like(w)
                _ext.ext().grouped_wgrad(x.contiguous(), gy, gw, seg, ctx.block_size, False)
        return gx, gw, None, None, None


def grouped_matmul(x: torch.Tensor, w: torch.Tensor, block_to_expert: torch.Tensor, seg_first_block: torch.Tensor,
                   block_size: int) -> torch.Tensor:
    """``y[i·B:(i+1)·B] = x[i·B:(i+1)·B] @ w[block_to_expert[i]]`` with ``x [nb·B, K]``, ``w [E, K, N]``."""
    if _grouped_ok(x, w, block_size):
        return _GroupedMatmul.apply(x, w, block_to_expert.int(), seg_first_block.int(), block_size)
    return _grouped_ref(x, w, block_to_expert, block_size)
