"""Mixtral-style sparse-MoE decoder (role of the reference's Mixtral/DBRX training examples,
``examples/training/mixtral``): Llama attention + an :class:`modules.moe.MoE` feed-forward per layer, with the
router auxiliary loss added to the LM loss."""
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


class MixtralDecoderLayer(nn.Module):
    def __init__(self, cfg: MixtralConfig):
        super().__init__()
        sp = cfg.sequence_parallel_enabled
        self.input_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, sp, cfg.dtype, cfg.device)
        self.post_attention_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, sp, cfg.dtype, cfg.device)
        self.self_attn = LlamaAttention(cfg)
        ecfg = RoutedExpertsMLPOpsConfig(num_experts=cfg.num_local_experts, top_k=cfg.num_experts_per_tok,
                                         hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                                         hidden_act="silu", glu_mlp=True, capacity_factor=cfg.capacity_factor)
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
        self.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps, cfg.sequence_parallel_enabled, cfg.dtype, cfg.device)
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
