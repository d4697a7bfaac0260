"""Operand / scale-factor interleaves for block-scaled UMMA (role of reference
``experimental/quantization/microscaling/swizzle.py:6-56``).

The reference "swizzles" data ``[M, N] → [M/4, 4N]`` so that four K-neighbours share an SBUF partition.  Blackwell has
no such requirement for the *data* (x4 packing along a contiguous K already is that layout, see
``transform_weights``); what tcgen05 block-scaled MMAs do need is the **scale-factor** tile interleave: scales for a
128-row × 4-block (= 128 K-elements at block 32) tile are stored as 32 rows of 16 bytes,
``byte(r, c) = (r % 32)·16 + (r // 32)·4 + c``, tiles ordered K-block-major within a 128-row band — the layout
``tcgen05.cp`` expects in shared memory and ``cublasLt`` calls ``VEC32_UE8M0``.  ``swizzle_scale_factors`` /
``unswizzle_scale_factors`` implement it; the reference-named data helpers are kept for checkpoint tools that were
written against them (they are pure index permutations and device-agnostic).
"""
from __future__ import annotations

import torch


def swizzle_tensor(input_tensor: torch.Tensor) -> torch.Tensor:
    """``[M, N] → [M/4, 4N]``: rows 4i…4i+3 become the four fastest-varying entries of row i."""
    m, n = input_tensor.shape
    assert m % 4 == 0, f"M dimension ({m}) must be divisible by 4"
    return input_tensor.reshape(m // 4, 4, n).permute(0, 2, 1).reshape(m // 4, 4 * n).contiguous()


def unswizzle_tensor(swizzled_tensor: torch.Tensor, original_M: int, original_N: int) -> torch.Tensor:
    m4, n4 = swizzled_tensor.shape
    assert original_M == m4 * 4 and original_N * 4 == n4, "Dimension mismatch"
    return swizzled_tensor.reshape(m4, original_N, 4).permute(0, 2, 1).reshape(original_M, original_N).contiguous()


def swizzle_tiled_tensor(input_tensor: torch.Tensor) -> torch.Tensor:
    """``[TILE_M, NUM_TILES, N] → [TILE_M, NUM_TILES/4, 4N]`` grouping four consecutive tiles and four consecutive rows
    (same permutation as the reference, expressed as one permute)."""
    tm, nt, n = input_tensor.shape
    g = input_tensor.reshape(tm // 4, 4, nt // 4, 4, n)
    return g.permute(3, 0, 2, 4, 1).reshape(tm, nt // 4, 4 * n).contiguous()


def swizzle_scale_factors(scale: torch.Tensor) -> torch.Tensor:
    """E8M0 scales ``[M, KB]`` (KB = K/32) → the 1-D byte stream of 128×4 tiles tcgen05 consumes.  M and KB are padded
    with 2⁰ (=127) up to multiples of 128 / 4."""
    m, kb = scale.shape
    mp, kp = -(-m // 128) * 128, -(-kb // 4) * 4
    s = torch.full((mp, kp), 127, dtype=scale.dtype, device=scale.device)
    s[:m, :kb] = scale
    # [band, r//32 (4), r%32 (32), ktile, c (4)] → [band, ktile, r%32, r//32, c]
    t = s.reshape(mp // 128, 4, 32, kp // 4, 4).permute(0, 3, 2, 1, 4)
    return t.reshape(-1).contiguous()


def unswizzle_scale_factors(stream: torch.Tensor, m: int, kb: int) -> torch.Tensor:
    mp, kp = -(-m // 128) * 128, -(-kb // 4) * 4
    t = stream.reshape(mp // 128, kp // 4, 32, 4, 4).permute(0, 3, 2, 1, 4)
    return t.reshape(mp, kp)[:m, :kb].contiguous()
