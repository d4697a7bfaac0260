"""Drop-in replacements for HF ``modeling_llama`` helpers used by the reference's example models
(``overrides/transformer_overrides.py:4-35``).

With ``flash_attn and transpose`` the tensors are in the attention kernel's ``[B, H, D, S]`` operand layout (head dim on
axis -2), so "rotate half" acts on axis -2 and the cos/sin tables are transposed to match; otherwise this is the usual
rotary embedding on ``[B, H, S, D]``.  For the models in this package RoPE runs as a fused CUDA kernel (``ops.rope``);
these functions exist for HF-derived model code that is ported as is."""
from __future__ import annotations

import torch


def rotate_half(x: torch.Tensor, flash_attn: bool = False, transpose: bool = False) -> torch.Tensor:
    dim = -2 if (flash_attn and transpose) else -1
    x1, x2 = x.chunk(2, dim=dim)
    return torch.cat((-x2, x1), dim=dim)


def apply_rotary_pos_emb(q, k, cos, sin, position_ids=None, flash_attn: bool = False, transpose_nki_inputs: bool = True):
    if position_ids is not None:                    # [max_pos, D] tables indexed per token (older HF signature)
        cos, sin = cos[position_ids], sin[position_ids]
    cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)   # [B, 1, S, D]
    if flash_attn and transpose_nki_inputs:
        cos, sin = cos.transpose(-1, -2), sin.transpose(-1, -2)                   # [B, 1, D, S]
    q_embed = q * cos + rotate_half(q, flash_attn, transpose_nki_inputs) * sin
    k_embed = k * cos + rotate_half(k, flash_attn, transpose_nki_inputs) * sin
    return q_embed, k_embed
