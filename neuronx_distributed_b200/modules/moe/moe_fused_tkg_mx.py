"""Decode-time fused MoE block with MXFP4 expert weights (role of reference ``modules/moe/moe_fused_tkg_mx.py:81-260``).

At decode the expert GEMVs are pure weight streaming, so 4-bit weights cut the bytes per token ~4× against bf16.  The
expert weights are kept K-major, x4-packed along the contraction dim with one E8M0 scale per 32 elements —
``gate_up [E, 2I/tp, H/4]`` + ``[E, 2I/tp, H/32]`` and ``down [E, H, (I/tp)/4]`` + ``[E, H, (I/tp)/32]`` — which is the
operand layout of ``tcgen05.mma.kind::mxf4.block_scale`` (and of an fp4 GEMV: a warp reads 128 contiguous bytes = 256
weights and one 8-byte scale group).  The math below is the numerics contract of that kernel: de-quantise (16-entry
table + exponent add), contract in fp32, bf16 result.

``MoEFusedTKGMX`` packs the float experts once (``prepare_mx_weights``) and then runs
RMSNorm → router → top-k → experts (selected experts only when ``T·k < E``) → shared experts → one reduction.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from ...parallel_layers import mappings
from ...quantization.microscaling.transform_weights import get_mxfp4_tensor, pack_fp4_x4_uint16, quantize_to_mxfp4
from .model_utils import ACT2FN
from .moe_configs import (  # noqa: F401  (enums re-exported as in the reference module)
    ROUTER_ACT_FN_MAPPING,
    ActFnType,
    ExpertAffinityScaleMode,
    MoEFusedTKGConfig,
    RouterActFnType,
)
from .moe_fused_tkg import MoEFusedTKG


def pack_expert_weight_mxfp4(w_ekn: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Float expert weights ``[E, K, N]`` → K-major MXFP4 ``([E, N, K/4] uint16, [E, N, K/32] uint8)``."""
    blocks, scales = quantize_to_mxfp4(w_ekn.transpose(1, 2).contiguous())          # [E, N, K/32, 16], [E, N, K/32]
    return pack_fp4_x4_uint16(blocks).reshape(*blocks.shape[:2], -1), scales


def _dequant(w_x4: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """``[..., N, K/4] uint16`` + ``[..., N, K/32]`` → fp32 ``[..., N, K]``."""
    blocks = w_x4.contiguous().view(torch.uint8).reshape(*scale.shape, 16)
    return get_mxfp4_tensor(blocks, scale, dtype=torch.float32)


def mxfp4_moe_block_tkg_wrapper(inp: torch.Tensor, gamma: Optional[torch.Tensor], router_weights: torch.Tensor,
                                expert_gate_up_weights: torch.Tensor, expert_down_weights: torch.Tensor,
                                shared_expert_gate_w: Optional[torch.Tensor] = None,
                                shared_expert_up_w: Optional[torch.Tensor] = None,
                                shared_expert_down_w: Optional[torch.Tensor] = None,
                                expert_gate_up_weights_scale: Optional[torch.Tensor] = None,
                                expert_down_weights_scale: Optional[torch.Tensor] = None,
                                router_bias: Optional[torch.Tensor] = None, expert_gate_up_bias: Optional[torch.Tensor] = None,
                                expert_down_bias: Optional[torch.Tensor] = None, eps: float = 1e-6, top_k: int = 1,
                                router_act_fn: str = "sigmoid", router_pre_norm: bool = True, norm_topk_prob: bool = False,
                                hidden_act_fn: str = "silu", hidden_act_scale_factor: Optional[float] = None,
                                hidden_act_bias: Optional[float] = None, gate_clamp_upper_limit: Optional[float] = None,
                                gate_clamp_lower_limit: Optional[float] = None, up_clamp_upper_limit: Optional[float] = None,
                                up_clamp_lower_limit: Optional[float] = None, is_all_expert: bool = False,
                                residual: Optional[torch.Tensor] = None, skip_router_logits: bool = False, **_unused):
    """Functional decode MoE block on MXFP4 experts.  ``inp [T, H]``; ``router_weights [E, H]``; packed expert weights /
    scales as in the module docstring (gate|up halves concatenated along N); shared-expert weights in ``[out, in]``.
    Returns ``(out [T, H] partial over TP, router_logits [T, E] or None)`` (+ the updated residual when given)."""
    x = inp if residual is None else inp + residual
    xf = x.float()
    h = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps) * gamma.float() if gamma is not None else xf
    logits = torch.nn.functional.linear(h if router_pre_norm else xf, router_weights.float(),
                                        None if router_bias is None else router_bias.float())
    aff = torch.sigmoid(logits) if router_act_fn == "sigmoid" else torch.softmax(logits, -1)
    top_w, top_i = torch.topk(aff, top_k, dim=-1)
    if norm_topk_prob:
        top_w = top_w / top_w.sum(-1, keepdim=True).clamp(min=1e-12)
    T, E = aff.shape
    act = ACT2FN[hidden_act_fn]

    def glu(gu: torch.Tensor) -> torch.Tensor:
        g, u = gu.chunk(2, -1)
        if gate_clamp_upper_limit is not None or gate_clamp_lower_limit is not None:
            g = g.clamp(min=gate_clamp_lower_limit, max=gate_clamp_upper_limit)
        if up_clamp_upper_limit is not None or up_clamp_lower_limit is not None:
            u = u.clamp(min=up_clamp_lower_limit, max=up_clamp_upper_limit)
        if hidden_act_scale_factor is not None:                      # gpt-oss: g·σ(α g)·(u + β)
            return g * torch.sigmoid(hidden_act_scale_factor * g) * (u + (hidden_act_bias or 0.0))
        return act(g) * u

    if is_all_expert or T * top_k >= E:
        w1, w2 = _dequant(expert_gate_up_weights, expert_gate_up_weights_scale), _dequant(expert_down_weights, expert_down_weights_scale)
        gu = torch.einsum("th,enh->etn", h, w1)
        if expert_gate_up_bias is not None:
            gu = gu + expert_gate_up_bias.float().unsqueeze(1)
        y = torch.einsum("eti,ehi->eth", glu(gu), w2)
        if expert_down_bias is not None:
            y = y + expert_down_bias.float().unsqueeze(1)
        dense = torch.zeros(T, E, dtype=torch.float32, device=x.device).scatter_(1, top_i, top_w)
        out = torch.einsum("eth,te->th", y, dense)
    else:                                                            # touch only the T·k selected experts' bytes
        from ...ops import gemm_mx

        fe = top_i.reshape(-1)
        hk = h.repeat_interleave(top_k, 0)
        # one (token, slot) row per chosen expert: the MX codes of only those experts are read (decoded in registers by
        # ``gemv_mx_grouped`` on CUDA, gathered + de-quantised otherwise); expert ids never leave the device
        gu = gemm_mx.grouped_linear_mx(hk.to(inp.dtype), expert_gate_up_weights, expert_gate_up_weights_scale, fe, "mxfp4").float()
        if expert_gate_up_bias is not None:
            gu = gu + expert_gate_up_bias.float()[fe]
        y = gemm_mx.grouped_linear_mx(glu(gu).to(inp.dtype), expert_down_weights, expert_down_weights_scale, fe, "mxfp4").float()
        if expert_down_bias is not None:
            y = y + expert_down_bias.float()[fe]
        out = (y * top_w.reshape(-1, 1)).reshape(T, top_k, -1).sum(1)
    if shared_expert_gate_w is not None:
        sg = torch.nn.functional.linear(h, shared_expert_gate_w.float())
        su = torch.nn.functional.linear(h, shared_expert_up_w.float())
        out = out + torch.nn.functional.linear(act(sg) * su, shared_expert_down_w.float())
    out = out.to(inp.dtype)
    res = (out, None if skip_router_logits else logits)
    return res if residual is None else res + (x,)


class MoEFusedTKGMX(MoEFusedTKG):
    """``MoEFusedTKG`` whose routed experts run from MXFP4 copies of the weights (``config.is_mxfp4_compute``)."""

    def __init__(self, router: nn.Module, expert_mlps: nn.Module, shared_experts=None, rmsnorm=None,
                 config: Optional[MoEFusedTKGConfig] = None, sequence_dimension: int = 0,
                 post_attention_layernorm: Optional[nn.Module] = None, tensor_model_parallel_group=None,
                 logical_nc_config: int = 1, return_router_logits: bool = False, return_expert_index: bool = False):
        """Same two constructor orders as :class:`MoEFusedTKG` (this package's and the reference's
        ``(router, expert_mlps, config, sequence_dimension, shared_experts, post_attention_layernorm, …)``)."""
        super().__init__(router, expert_mlps, shared_experts, rmsnorm, config, sequence_dimension, post_attention_layernorm,
                         tensor_model_parallel_group, logical_nc_config, return_router_logits, return_expert_index)
        self.config.is_mxfp4_compute = True
        self.prepare_mx_weights()

    @torch.no_grad()
    def prepare_mx_weights(self) -> None:
        ops = self._experts.mlp_op
        assert ops.glu_mlp, "the MX decode block implements GLU experts"
        gu, gus = pack_expert_weight_mxfp4(ops.gate_up_proj.weight.data.float())
        dn, dns = pack_expert_weight_mxfp4(ops.down_proj.weight.data.float())
        for name, t in (("gate_up_x4", gu), ("gate_up_scale", gus), ("down_x4", dn), ("down_scale", dns)):
            self.register_buffer(name, t, persistent=True)

    def forward(self, hidden_states: torch.Tensor, residual: Optional[torch.Tensor] = None):
        x = hidden_states if residual is None else hidden_states + residual
        h = self._norm(x) if self._norm is not None else x
        logits, aff, idx = self._router(h)
        em, ops = self._experts, self._experts.mlp_op
        flat = h.reshape(-1, h.shape[-1])
        a = em._topk_affinities(aff, idx)                                           # [T, E]
        local = torch.as_tensor(ops.local_expert_ids, device=flat.device)
        w1 = _dequant(self.gate_up_x4, self.gate_up_scale)                          # [E_l, 2I/tp, H]
        w2 = _dequant(self.down_x4, self.down_scale)                                # [E_l, H, I/tp]
        gu = torch.einsum("th,enh->etn", flat.float(), w1)
        if ops.gate_up_proj.bias is not None:
            gu = gu + ops.gate_up_proj.bias.float().unsqueeze(1)
        y = torch.einsum("eti,ehi->eth", ops.activation(gu), w2)
        if ops.down_proj.bias is not None:
            y = y + ops.down_proj.bias.float().unsqueeze(1) / ops.down_proj.tp
        y = torch.einsum("eth,te->th", y, a[:, local].float()).to(h.dtype).view(h.shape)
        if self._shared is not None:
            y = y + self._shared(h)
        y = self._reduce(y)
        out = (y,)
        if self.return_router_logits:
            out += (logits,)
        if self.return_expert_index:
            out += (idx,)
        return out if residual is None else out + (x,)
