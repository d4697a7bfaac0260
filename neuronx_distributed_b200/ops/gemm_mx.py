"""Linear layers on microscaling (OCP MX) weights.

Weights are the K-contiguous byte stream the checkpoints ship (``quantization/microscaling``): MXFP4 = e2m1 codes two per
byte (low nibble first; ``uint16`` x4 words are a view of the same bytes), MXFP8 = one e4m3 / e5m2 byte per element
(``uint32`` x4 words likewise); one E8M0 scale per 32 elements of a row.

* ``rows <= 8`` on CUDA (token generation): ``csrc/gemv_mx.cu`` — the codes are decoded in registers, nothing is expanded in
  memory, so a decode step reads 4.25 / 8.25 bits per weight instead of 16;
* MXFP8 or MXFP4 weights, more rows, opt-in ``NXD_GEMM_MX=1``: activations are quantised to MXFP8 online and both operands go through
  the block-scaled tensor-core GEMM (``csrc/gemm_mx_sm100.cu``, ``tcgen05.mma.kind::mxf8f6f4.block_scale``: the E8M0 scales
  are applied inside the tensor core, nothing is de-quantised) — W8A8-MX numerics, oracle :func:`matmul_mxfp8_reference`;
* MXFP4 weights, opt-in ``NXD_GEMM_F4=1``: activations quantised to MXFP4 as well and both operands read packed at the FP4 rate
  (``csrc/gemm_mxf4_sm100.cu``, ``kind::mxf4`` with E8M0 scales; the same kernel runs NVFP4 — UE4M3 scales per 16 elements —
  through :func:`matmul_f4`) — W4A4 numerics, oracle :func:`matmul_f4_reference`;
* otherwise: de-quantise to the activation dtype and run the dense GEMM (tcgen05 bf16 kernel on CUDA) — numerically the
  oracle ``experimental…mx_torch.mx_matmul``."""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _ext

_FMT = {"mxfp4": 0, "mxfp8": 1, "mxfp8_e4m3": 1, "mxfp8_e5m2": 2}                    # gemv_mx format codes
_MMA_FMT = {"mxfp4": 5, "mxfp8": 0, "mxfp8_e4m3": 0, "mxfp8_e5m2": 1}               # kind::mxf8f6f4 operand formats


def kind_of(weight: torch.Tensor, fp8_dtype: torch.dtype = torch.float8_e4m3fn) -> str:
    if weight.dtype in (torch.uint16, torch.float16, torch.int16):
        return "mxfp4"
    if weight.dtype in (torch.uint32, torch.int32):
        return "mxfp8_e5m2" if fp8_dtype == torch.float8_e5m2 else "mxfp8"
    raise ValueError(f"not an x4-packed MX weight: {weight.dtype}")


def dequantize(weight: torch.Tensor, scale: torch.Tensor, kind: str, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """``[N, K/4]`` x4 words (or the raw byte stream) + ``[N, K/32]`` E8M0 → ``[N, K]``."""
    b = weight.contiguous().view(torch.uint8)
    N = scale.shape[0]
    if _FMT[kind] == 0:
        lut = torch.tensor([0, .5, 1, 1.5, 2, 3, 4, 6, -0., -.5, -1, -1.5, -2, -3, -4, -6], dtype=torch.float32, device=b.device)
        vals = torch.stack((lut[(b & 0xF).long()], lut[(b >> 4).long()]), dim=-1).reshape(N, -1)
    else:
        vals = b.view(torch.float8_e4m3fn if _FMT[kind] == 1 else torch.float8_e5m2).float().reshape(N, -1)
    return torch.ldexp(vals.reshape(N, -1, 32), (scale.to(torch.int32) - 127).unsqueeze(-1)).reshape(N, -1).to(dtype)


def tile_scales(scale: torch.Tensor, pad: int = 127) -> torch.Tensor:
    """E8M0 scales ``[R, K/32]`` → the chunked layout the block-scaled GEMM streams: ``[ceil(R/128), K/128, 512]`` where the
    512-byte chunk of (row tile, 128-wide K block) holds scale (r, j) at byte ``(r % 32) * 16 + (r // 32) * 4 + j`` — what
    ``tcgen05.cp.32x128b.warpx4`` spreads over 4 TMEM columns.  Rows are padded with ``pad`` (E8M0 127 = 2^0).  The same chunking
    serves the 4-bit kinds (:func:`matmul_f4`): a chunk is always 128 rows × 4 consecutive scales, whatever they span."""
    R, KB = scale.shape
    assert KB % 4 == 0, "the number of scale columns must be a multiple of 4"
    T = (R + 127) // 128
    if T * 128 != R:
        scale = torch.cat([scale, torch.full((T * 128 - R, KB), pad, dtype=scale.dtype, device=scale.device)])
    s = scale.reshape(T, 4, 32, KB // 4, 4)                       # [tile, r // 32, r % 32, k block of 128, j]
    return s.permute(0, 3, 2, 1, 4).contiguous().reshape(T, KB // 4, 512)


def matmul_mxfp8_reference(a_q: torch.Tensor, a_scale: torch.Tensor, b_q: torch.Tensor, b_scale: torch.Tensor,
                           a_kind: str = "mxfp8", b_kind: str = "mxfp8") -> torch.Tensor:
    """fp32 oracle of the block-scaled GEMM: ``dequant(a) @ dequant(b)ᵀ``."""
    return dequantize(a_q, a_scale, a_kind) @ dequantize(b_q, b_scale, b_kind).t()


def gemm_mx_eligible(x2d: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor, kind: str) -> bool:
    if os.environ.get("NXD_GEMM_MX", "0") != "1":                 # opt-in until the kernel has run on hardware
        return False
    return (x2d.is_cuda and kind in _MMA_FMT and x2d.shape[1] % 128 == 0 and scale.shape[0] % 8 == 0
            and weight.is_contiguous() and _ext.use_cuda(x2d, weight, scale) and hasattr(_ext.ext(), "gemm_mxfp8"))


def matmul_mxfp8(a_q: torch.Tensor, a_scale: torch.Tensor, b_q: torch.Tensor, b_scale: torch.Tensor, a_kind: str = "mxfp8",
                 b_kind: str = "mxfp8", b_scale_tiled: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``[M, K] · [N, K]ᵀ`` on MX operands (byte streams or x4 words; MXFP8 or MXFP4 — the TMA engine unpacks e2m1 pairs into the
    8-bit containers the tensor core reads) with E8M0 block scales → bf16 ``[M, N]``."""
    a8 = a_q.contiguous().view(torch.uint8).reshape(a_scale.shape[0], -1)
    b8 = b_q.contiguous().view(torch.uint8).reshape(b_scale.shape[0], -1)
    _ext.count_launch()
    return _ext.ext().gemm_mxfp8(a8, b8, tile_scales(a_scale), b_scale_tiled if b_scale_tiled is not None else tile_scales(b_scale),
                                 _MMA_FMT[a_kind], _MMA_FMT[b_kind])


# ---- 4-bit × 4-bit at the FP4 rate (csrc/gemm_mxf4_sm100.cu) ------------------------------------------------------------
_E2M1 = (0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0)


def _e2m1_codes(v: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even onto the e2m1 grid (|v| ≤ 6 after scaling) → 4-bit codes (bit 3 = sign)."""
    mids = torch.tensor([0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0], dtype=torch.float32, device=v.device)
    mag = v.abs().float().clamp(max=6.0)
    lo, hi = torch.bucketize(mag, mids, right=False), torch.bucketize(mag, mids, right=True)      # differ only on exact ties
    idx = torch.where(lo % 2 == 0, lo, hi)
    return (idx | ((v < 0).long() << 3)).to(torch.uint8)


def _pack_nibbles(codes: torch.Tensor) -> torch.Tensor:
    c = codes.reshape(*codes.shape[:-1], -1, 2)
    return (c[..., 0] | (c[..., 1] << 4)).contiguous()


def quantize_mxfp4(x: torch.Tensor) -> "tuple[torch.Tensor, torch.Tensor]":
    """``[R, K]`` → (packed e2m1 bytes ``[R, K/2]``, E8M0 scales ``[R, K/32]``); OCP MXFP4: scale = 2^(⌊log2 amax⌋ − 2)."""
    R, K = x.shape
    xb = x.float().reshape(R, K // 32, 32)
    e = torch.floor(torch.log2(xb.abs().amax(-1).clamp(min=2.0 ** -120))) - 2
    codes = _e2m1_codes(torch.ldexp(xb, -e.to(torch.int32).unsqueeze(-1)).reshape(R, K))
    return _pack_nibbles(codes), (e + 127).clamp(0, 254).to(torch.uint8)


def quantize_nvfp4(x: torch.Tensor) -> "tuple[torch.Tensor, torch.Tensor, torch.Tensor]":
    """``[R, K]`` → (packed e2m1 bytes ``[R, K/2]``, UE4M3 block scales ``[R, K/16]`` as bytes, fp32 per-tensor factor): a block's
    scale is ``amax_block / 6`` expressed in units of the per-tensor factor ``amax_tensor / (6 · 448)`` and rounded to e4m3."""
    R, K = x.shape
    xb = x.float().reshape(R, K // 16, 16)
    g = (xb.abs().amax() / (6.0 * 448.0)).clamp(min=2.0 ** -120)
    sf = (xb.abs().amax(-1) / 6.0 / g).clamp(max=448.0).to(torch.float8_e4m3fn)
    eff = (sf.float() * g).clamp(min=2.0 ** -120).unsqueeze(-1)
    codes = _e2m1_codes((xb / eff).reshape(R, K))
    return _pack_nibbles(codes), sf.view(torch.uint8), g.reshape(())


def dequantize_f4(codes: torch.Tensor, scale: torch.Tensor, vec_size: int, global_scale=1.0) -> torch.Tensor:
    """Packed e2m1 ``[R, K/2]`` + block scales ``[R, K/vec_size]`` (E8M0 for 32, UE4M3 for 16) → fp32 ``[R, K]``."""
    b = codes.contiguous().view(torch.uint8)
    R = scale.shape[0]
    lut = torch.tensor(_E2M1 + tuple(-v for v in _E2M1), dtype=torch.float32, device=b.device)
    vals = torch.stack((lut[(b & 0xF).long()], lut[(b >> 4).long()]), dim=-1).reshape(R, -1, vec_size)
    if vec_size == 32:
        return torch.ldexp(vals, (scale.to(torch.int32) - 127).unsqueeze(-1)).reshape(R, -1)
    return (vals * (scale.view(torch.float8_e4m3fn).float() * global_scale).unsqueeze(-1)).reshape(R, -1)


def matmul_f4_reference(a_q, a_scale, b_q, b_scale, vec_size: int = 32, a_global=1.0, b_global=1.0) -> torch.Tensor:
    """fp32 oracle of the 4-bit block-scaled GEMM: ``dequant(a) @ dequant(b)ᵀ``."""
    return dequantize_f4(a_q, a_scale, vec_size, a_global) @ dequantize_f4(b_q, b_scale, vec_size, b_global).t()


def gemm_f4_eligible(M: int, N: int, K: int, *tensors: torch.Tensor) -> bool:
    if os.environ.get("NXD_GEMM_F4", "0") != "1":                 # opt-in: W4A4 numerics, and the kernel has not run on hardware yet
        return False
    return (K % 256 == 0 and N % 8 == 0 and all(t.is_cuda for t in tensors) and _ext.use_cuda(*tensors)
            and hasattr(_ext.ext(), "gemm_f4"))


def matmul_f4(a_q: torch.Tensor, a_scale: torch.Tensor, b_q: torch.Tensor, b_scale: torch.Tensor, vec_size: int = 32,
              a_global=1.0, b_global=1.0, b_scale_tiled: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``[M, K] · [N, K]ᵀ`` with BOTH operands e2m1 (packed bytes or x4 words) → bf16 ``[M, N]`` on
    ``tcgen05.mma.kind::mxf4`` (``vec_size`` 32, E8M0 scales) or ``kind::mxf4nvf4`` (``vec_size`` 16, UE4M3 scales, the product of
    the per-tensor factors applied in the epilogue): twice the tensor-core rate and half the operand bytes of the 8-bit kinds."""
    a8 = a_q.contiguous().view(torch.uint8).reshape(a_scale.shape[0], -1)
    b8 = b_q.contiguous().view(torch.uint8).reshape(b_scale.shape[0], -1)
    pad = 127 if vec_size == 32 else 0x38                          # 2^0 / 1.0: padded rows never reach the output anyway
    _ext.count_launch()
    return _ext.ext().gemm_f4(a8, b8, tile_scales(a_scale, pad), b_scale_tiled if b_scale_tiled is not None else tile_scales(b_scale, pad),
                              int(vec_size), float(a_global) * float(b_global))


def gemv_eligible(x2d: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor) -> bool:
    if os.environ.get("NXD_GEMV_MX", "0") != "1":                 # opt-in until the kernel has run on hardware
        return False
    return (x2d.is_cuda and x2d.dtype == torch.bfloat16 and x2d.dim() == 2 and 1 <= x2d.shape[0] <= 8 and x2d.shape[1] % 32 == 0
            and x2d.is_contiguous() and weight.is_contiguous() and scale.is_contiguous() and scale.dtype == torch.uint8
            and scale.dim() == 2 and _ext.use_cuda(x2d, weight, scale) and hasattr(_ext.ext(), "gemv_mx"))


def grouped_linear_mx(x: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor, expert: torch.Tensor, kind: Optional[str] = None
                      ) -> torch.Tensor:
    """Row ``s`` of ``x [S, K]`` against expert ``expert[s]`` of the stacked MX weights ``[E, N, K]`` (x4 words or bytes) with scales
    ``[E, N, K/32]`` → ``[S, N]``: the selective-loading step of a decode-time MoE block.  CUDA (opt-in ``NXD_GEMV_MX=1``): one
    launch that reads only the chosen experts' codes, ids stay on the device; otherwise gather + de-quantise + batched matmul."""
    kind = kind or kind_of(weight)
    E, N = scale.shape[0], scale.shape[1]
    if (os.environ.get("NXD_GEMV_MX", "0") == "1" and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 2 and x.shape[1] % 32 == 0
            and _ext.use_cuda(x, weight, scale) and hasattr(_ext.ext(), "gemv_mx_grouped")):
        _ext.count_launch()
        return _ext.ext().gemv_mx_grouped(x.contiguous(), weight.contiguous(), scale.contiguous(), expert.long().contiguous(), _FMT[kind])
    wb = weight.contiguous().view(torch.uint8).reshape(E, N, -1)[expert.long()]                      # [S, N, bytes]
    w = dequantize(wb.reshape(-1, wb.shape[-1]), scale[expert.long()].reshape(-1, scale.shape[-1]), kind).view(x.shape[0], N, -1)
    return torch.einsum("sk,snk->sn", x.float(), w).to(x.dtype)


def linear_mx(x: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor, kind: Optional[str] = None,
              residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``x [..., K] @ dequant(weight)[N, K]ᵀ (+ residual)`` → ``[..., N]`` in ``x.dtype``."""
    kind = kind or kind_of(weight)
    x2 = x.reshape(-1, x.shape[-1])
    if gemv_eligible(x2.contiguous(), weight, scale):
        _ext.count_launch()
        r2 = None if residual is None else residual.reshape(-1, residual.shape[-1]).contiguous()
        y = _ext.ext().gemv_mx(x2.contiguous(), weight, scale, _FMT[kind], r2)
        return y.view(*x.shape[:-1], y.shape[-1])
    if kind == "mxfp4" and gemm_f4_eligible(x2.shape[0], scale.shape[0], x2.shape[1], x2, weight, scale):
        xq, xs = quantize_mxfp4(x2)                                  # W4A4: activations to MXFP4 online, both operands at the FP4 rate
        y = matmul_f4(xq, xs, weight, scale, 32)
        if residual is not None:
            y = y + residual.reshape(-1, residual.shape[-1]).to(y.dtype)
        return y.view(*x.shape[:-1], y.shape[-1]).to(x.dtype)
    if gemm_mx_eligible(x2, weight, scale, kind):
        from ..quantization.microscaling.mx_torch import quantize_mxfp8

        xq, xs = quantize_mxfp8(x2.contiguous())
        y = matmul_mxfp8(xq, xs, weight, scale, "mxfp8", kind)
        if residual is not None:
            y = y + residual.reshape(-1, residual.shape[-1]).to(y.dtype)
        return y.view(*x.shape[:-1], y.shape[-1]).to(x.dtype)
    w = dequantize(weight, scale, kind, x.dtype if x.dtype in (torch.bfloat16, torch.float16) else torch.float32)
    y = torch.nn.functional.linear(x.to(w.dtype), w)
    if residual is not None:
        y = y + residual.to(y.dtype)
    return y.to(x.dtype)
