"""Mixtral / DBRX style sparse-MoE decoders (role of the reference's ``examples/training/mixtral`` and
``examples/training/dbrx``): Llama attention + an :class:`modules.moe.MoE` feed-forward per layer, with the router
auxiliary loss added to the LM loss.  DBRX differs by bias-free LayerNorm instead of RMSNorm, clamped QKV projections
(``clip_qkv``), 16 experts / top-4 with normalised top-k affinities (``DbrxConfig``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn

from ..modules.moe import MoE, ExpertMLPsV2, RoutedExpertsMLPOpsConfig, RouterTopK, load_balancing_loss_func
from ..modules.rms_norm import RMSNorm
from ..parallel_layers.layers import ColumnParallelLinear, ParallelEmbedding
from ..parallel_layers.loss_functions import parallel_cross_entropy
from .llama import LlamaAttention, LlamaConfig


@dataclass
class MixtralConfig(LlamaConfig):
    num_local_experts: int = 8
    num_experts_per_tok: int = 2
    router_aux_loss_coef: float = 0.02
    capacity_factor: Optional[float] = None
    intermediate_size: int = 14336
    norm_type: str = "rmsnorm"            # "layernorm" for DBRX (no bias)
    clip_qkv: Optional[float] = None
    normalize_top_k_affinities: bool = True


@dataclass
class DbrxConfig(MixtralConfig):
    """databricks/dbrx-base shapes: d_model 6144, 40 layers, 48 heads / 8 kv heads, ffn 10752, 16 experts top-4, vocab 100352."""
    vocab_size: int = 100352
    hidden_size: int = 6144
    num_hidden_layers: int = 40
    num_attention_heads: int = 48
    num_key_value_heads: int = 8
    intermediate_size: int = 10752
    num_local_experts: int = 16
    num_experts_per_tok: int = 4
    rope_theta: float = 500000.0
    norm_type: str = "layernorm"
    clip_qkv: Optional[float] = 8.0
    router_aux_loss_coef: float = 0.05


def _norm(cfg: MixtralConfig):
    if cfg.norm_type == "layernorm":
        from ..parallel_layers.layer_norm import LayerNorm

        return LayerNorm(cfg.hidden_size, 1e-5, sequence_parallel_enabled=cfg.sequence_parallel_enabled, dtype=cfg.dtype,
                         device=cfg.device, bias=False)
    return RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, cfg.sequence_parallel_enabled, cfg.dtype, cfg.device)


class MixtralDecoderLayer(nn.Module):
    def __init__(self, cfg: MixtralConfig):
        super().__init__()
        sp = cfg.sequence_parallel_enabled
        self.input_layernorm = _norm(cfg)
        self.post_attention_layernorm = _norm(cfg)
        self.self_attn = LlamaAttention(cfg)
        ecfg = RoutedExpertsMLPOpsConfig(num_experts=cfg.num_local_experts, top_k=cfg.num_experts_per_tok,
                                         hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                                         hidden_act="silu", glu_mlp=True, capacity_factor=cfg.capacity_factor,
                                         normalize_top_k_affinities=cfg.normalize_top_k_affinities)
        self.mlp = MoE(RouterTopK(cfg.num_local_experts, cfg.num_experts_per_tok, cfg.hidden_size,
                                  sequence_parallel_enabled=sp, sequence_dimension=0, device=cfg.device),
                       ExpertMLPsV2(ecfg, sequence_parallel_enabled=sp, dtype=cfg.dtype, device=cfg.device),
                       sequence_parallel_enabled=sp, sequence_dimension=0, return_router_logits=True)

    def forward(self, x, cos, sin):
        x = x + self.self_attn(self.input_layernorm(x), cos, sin)
        y, router_logits = self.mlp(self.post_attention_layernorm(x))
        return x + y, router_logits


class MixtralForCausalLM(nn.Module):
    _no_split_modules = ["MixtralDecoderLayer"]

    def __init__(self, cfg: MixtralConfig):
        super().__init__()
        from .. import ops

        self.config = cfg
        init = lambda w: nn.init.normal_(w, std=cfg.initializer_range)
        self.embed_tokens = ParallelEmbedding(cfg.vocab_size, cfg.hidden_size, init_method=init, dtype=cfg.dtype,
                                              sequence_parallel_enabled=cfg.sequence_parallel_enabled, device=cfg.device)
        self.layers = nn.ModuleList([MixtralDecoderLayer(cfg) for _ in range(cfg.num_hidden_layers)])
        self.norm = _norm(cfg)
        self.lm_head = ColumnParallelLinear(cfg.hidden_size, cfg.vocab_size, bias=False, gather_output=False, init_method=init,
                                            sequence_parallel_enabled=cfg.sequence_parallel_enabled, sequence_dimension=0,
                                            dtype=cfg.dtype, device=cfg.device)
        self._ops = ops

    def forward(self, input_ids, labels=None):
        B, S = input_ids.shape
        x = self.embed_tokens(input_ids)
        if not self.config.sequence_parallel_enabled:
            x = x.transpose(0, 1).contiguous()
        cos, sin = self._ops.rope.rope_tables(S, self.config.head_dim, self.config.rope_theta, input_ids.device)
        all_logits = []
        for layer in self.layers:
            x, rl = layer(x, cos, sin)
            all_logits.append(rl)
        logits = self.lm_head(self.norm(x))
        if labels is None:
            return None, logits
        tgt = labels.transpose(0, 1)
        tgt = torch.cat([tgt[1:], torch.full_like(tgt[:1], -100)], dim=0)
        mask = tgt != -100
        per_tok = parallel_cross_entropy(logits, torch.where(mask, tgt, torch.zeros_like(tgt)))
        loss = (per_tok * mask).sum() / mask.sum().clamp(min=1)
        aux = load_balancing_loss_func(all_logits, self.config.num_local_experts, self.config.num_experts_per_tok)
        return loss + self.config.router_aux_loss_coef * aux.to(loss.dtype), None
