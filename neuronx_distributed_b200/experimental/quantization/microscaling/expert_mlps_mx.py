"""MoE expert-MLP numerics oracles: bf16 all-experts / selected-experts, and MXFP4-weight × MXFP8-activation
(role of reference ``experimental/quantization/microscaling/expert_mlps_mx.py:15-299``, gpt-oss style clamped SwiGLU).

Layouts are the Blackwell ones (K contiguous, blocks of 32 along K):
``W_gate / W_up [E, I, H/4]`` x4-packed with ``scale [E, I, H/32]``; ``W_down [E, H, I/4]`` with ``scale [E, H, I/32]``.
"""
from __future__ import annotations

from typing import Optional

import torch

from ....quantization.microscaling.mx_torch import matmul_mx, quantize_mxfp8


def topk(router_logits: torch.Tensor, k: int = 4):
    top = torch.topk(router_logits, k)
    return torch.softmax(top.values, dim=1), top.indices


def expert_affinity_mask(router_logits: torch.Tensor, expert_index=None, k: int = 4) -> torch.Tensor:
    """Dense ``[T, E]`` matrix of softmax-over-top-k weights (zero for unselected experts)."""
    w, idx = topk(router_logits, k)
    mask = torch.zeros_like(router_logits).scatter_(1, idx, w)
    return mask if expert_index is None else mask[:, expert_index]


def swiglu(x_glu: torch.Tensor, x_linear: torch.Tensor, alpha: float = 1.702, limit: float = 7.0) -> torch.Tensor:
    x_glu = x_glu.clamp(max=limit)
    x_linear = x_linear.clamp(min=-limit, max=limit)
    return x_glu * torch.sigmoid(alpha * x_glu) * (x_linear + 1)


def expert_affinity_scale(down: torch.Tensor, expert_affinities_masked: torch.Tensor) -> torch.Tensor:
    assert down.shape[:2] == expert_affinities_masked.shape, (down.shape, expert_affinities_masked.shape)
    return torch.einsum("teh,te->th", down, expert_affinities_masked.to(down.dtype))


def all_expert_mlps_bf16(norm_out, router_logits, W_gate, W_up, W_down, bias_gate, bias_up, bias_down, k: int = 4):
    """Every expert on every token, weighted by the affinity mask.  ``W_gate/W_up [E, I, H]``, ``W_down [E, H, I]``."""
    mask = expert_affinity_mask(router_logits, k=k)
    gate = torch.einsum("eih,th->tei", W_gate, norm_out) + bias_gate
    up = torch.einsum("eih,th->tei", W_up, norm_out) + bias_up
    down = torch.einsum("ehi,tei->teh", W_down, swiglu(gate, up)) + bias_down
    return expert_affinity_scale(down, mask)


def select_expert_mlps_bf16(norm_out, router_logits, W_gate, W_up, W_down, bias_gate, bias_up, bias_down, k: int = 4):
    """Only the top-k experts of each token (gathers the expert weights per token — an oracle, not a fast path)."""
    w, idx = topk(router_logits, k)
    gate = torch.einsum("tkih,th->tki", W_gate[idx], norm_out) + bias_gate[idx]
    up = torch.einsum("tkih,th->tki", W_up[idx], norm_out) + bias_up[idx]
    down = torch.einsum("tkhi,tki->tkh", W_down[idx], swiglu(gate, up)) + bias_down[idx]
    return torch.einsum("tkh,tk->th", down, w.to(down.dtype))


def gate_up_projection_mx(input, input_scale, weight, scale, bias, matmul_accumulation_dtype=torch.float32,
                          matmul_output_dtype=torch.bfloat16):
    """``input [T, H/4]`` (fp8_x4) × ``weight [E, I, H/4]`` → ``[T, E, I]`` (+ bias ``[E, I]``)."""
    out = torch.stack([matmul_mx(input, weight[e], input_scale, scale[e], matmul_accumulation_dtype, matmul_output_dtype)
                       for e in range(weight.shape[0])], dim=1)
    return out + bias.to(out.dtype)


def down_projection_mx(act, act_scale, weight, scale, bias, matmul_accumulation_dtype=torch.float32,
                       matmul_output_dtype=torch.bfloat16):
    """``act [E, T, I/4]`` × ``weight [E, H, I/4]`` → ``[T, E, H]`` (+ bias ``[E, H]``)."""
    out = torch.stack([matmul_mx(act[e], weight[e], act_scale[e], scale[e], matmul_accumulation_dtype, matmul_output_dtype)
                       for e in range(weight.shape[0])], dim=1)
    return out + bias.to(out.dtype)


def all_expert_mlps_act_mxfp8_w_mxfp4(norm_out, W_gate, W_up, W_down, scale_gate, scale_up, scale_down, bias_gate, bias_up,
                                      bias_down, router_logits: Optional[torch.Tensor] = None, expert_index=None,
                                      expert_affinities_masked: Optional[torch.Tensor] = None,
                                      matmul_accumulation_dtype=torch.float32, matmul_output_dtype=torch.bfloat16,
                                      use_unbiased_scale_qmx_norm: bool = False, use_unbiased_scale_qmx_swiglu: bool = False,
                                      DBG: bool = False):
    """MXFP4 weights, activations quantised online to MXFP8 before each projection; result excludes the residual."""
    if expert_affinities_masked is None:
        expert_affinities_masked = expert_affinity_mask(router_logits, expert_index=expert_index)
    xq, xs = quantize_mxfp8(norm_out, use_unbiased_scale=use_unbiased_scale_qmx_norm)
    gate = gate_up_projection_mx(xq, xs, W_gate, scale_gate, bias_gate, matmul_accumulation_dtype, matmul_output_dtype)
    up = gate_up_projection_mx(xq, xs, W_up, scale_up, bias_up, matmul_accumulation_dtype, matmul_output_dtype)
    act = swiglu(gate, up)                                                    # [T, E, I]
    aq, asc = quantize_mxfp8(act.transpose(0, 1).contiguous(), use_unbiased_scale=use_unbiased_scale_qmx_swiglu)
    down = down_projection_mx(aq, asc, W_down, scale_down, bias_down, matmul_accumulation_dtype, matmul_output_dtype)
    result = expert_affinity_scale(down, expert_affinities_masked)
    if DBG:
        return xq, xs, gate, up, act, aq, asc, down, result
    return result
