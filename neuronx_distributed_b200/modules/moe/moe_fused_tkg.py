"""Decode-time fused MoE block (reference ``modules/moe/moe_fused_tkg.py:85-450`` around the ``moe_block_tkg`` NKI mega-kernel, K8).

For a handful of tokens the block is weight-bandwidth bound and, composed from separate ops, launch bound.  When the
conditions of :meth:`MoEFusedTKG._can_use_kernel` hold, ``(residual add →) RMSNorm → router → activation / top-k /
normalisation → chosen local experts (gate|up, GLU, down) → affinity-weighted sum`` runs as ONE cooperative kernel
(``csrc/moe_tkg.cu`` through ``ops.moe_tkg``); shared experts, when present, run next to it on the normalised tokens.  Otherwise
the same steps run as the individual modules (the reference's "flat" path).  Either way there is no collective inside the
block: ONE delayed all-reduce at the end sums the intermediate shards (TP) and the expert shards (EP).

The block owns no parameters: it holds references to the prefill modules (whose preshard hooks load the weights)."""
from __future__ import annotations

from typing import Any, Optional

import torch
from torch import nn

from ...parallel_layers import mappings
from ...parallel_layers import parallel_state as ps
from ...utils.logger import get_logger
from .moe_configs import (  # noqa: F401  (enums re-exported as in the reference module)
    ROUTER_ACT_FN_MAPPING,
    ActFnType,
    ExpertAffinityScaleMode,
    MoEFusedTKGConfig,
    RouterActFnType,
)

logger = get_logger()


def _on_cuda(t: torch.Tensor) -> bool:                 # separate so host tests can walk the eligibility rules
    return t.device.type == "cuda"


class MoEFusedTKG(nn.Module):
    def __init__(self, router: nn.Module, expert_mlps: nn.Module, shared_experts: Any = None, rmsnorm: Any = None,
                 config: Optional[MoEFusedTKGConfig] = None, sequence_dimension: int = 0,
                 post_attention_layernorm: Optional[nn.Module] = None, tensor_model_parallel_group=None,
                 logical_nc_config: int = 1, return_router_logits: bool = False, return_expert_index: bool = False):
        """Accepts this package's order ``(router, expert_mlps, shared_experts, rmsnorm, config)`` and the reference's
        ``(router, expert_mlps, config, sequence_dimension, shared_experts, post_attention_layernorm, …)``."""
        super().__init__()
        if isinstance(shared_experts, MoEFusedTKGConfig):                      # reference positional order
            ref_cfg, ref_seq = shared_experts, rmsnorm
            shared_experts, rmsnorm = (config if isinstance(config, nn.Module) else None), None
            config = ref_cfg
            if isinstance(ref_seq, int):
                sequence_dimension = ref_seq
        if post_attention_layernorm is not None:
            rmsnorm = post_attention_layernorm
        # references only (no new parameters, no duplicate state-dict entries): reuse the prefill modules' weights
        object.__setattr__(self, "_router", router)
        object.__setattr__(self, "_experts", expert_mlps)
        object.__setattr__(self, "_shared", shared_experts)
        object.__setattr__(self, "_norm", rmsnorm)
        self.config = config or MoEFusedTKGConfig()
        self.sequence_dimension = sequence_dimension
        self.tensor_parallel_group = tensor_model_parallel_group
        self.logical_nc_config = logical_nc_config
        self.return_router_logits, self.return_expert_index = return_router_logits, return_expert_index
        self._why_not: Optional[str] = None

    # reference attribute names
    @property
    def router(self):
        return self._router

    @property
    def expert_mlps(self):
        return self._experts

    @property
    def shared_experts(self):
        return self._shared

    @property
    def post_attention_layernorm(self):
        return self._norm

    def preshard_hook(self, model_state_dict, prefix: str) -> None:
        """Nothing to re-arrange: the block holds references to the prefill modules' parameters (whose own hooks run) and
        owns none itself (reference ``moe_fused_tkg.py:449-450``)."""

    # ---- dispatch ---------------------------------------------------------------------------------------------------------
    def _mlp_op(self):
        em = self._experts
        if getattr(em, "enabled_hybrid_sharding", False) and hasattr(em, "mlp_op_tkg"):
            return em.mlp_op_tkg
        return em.mlp_op

    def _can_use_kernel(self, hidden_states: torch.Tensor) -> bool:
        """Conditions of the one-launch kernel (cf. reference ``_can_use_nki_kernel`` :142-180).  The reason of the last
        refusal is kept in ``_why_not`` and logged once."""
        from ... import ops

        def no(reason: str) -> bool:
            if self._why_not != reason:
                logger.info("MoE fused decode kernel not used: %s", reason)
                self._why_not = reason
            return False

        if self.config.moe_fused_kernel_enabled is False:
            return no("disabled by config.moe_fused_kernel_enabled")
        if self.training or torch.is_grad_enabled() and hidden_states.requires_grad:
            return no("training / autograd")
        if not _on_cuda(hidden_states):
            return no("cannot run on cpu")
        if self.config.quantized or self.config.is_mxfp4_compute:
            return no("quantized experts")
        cfg, op, r = self._experts.routed_experts_mlp_config, self._mlp_op(), self._router
        if not cfg.glu_mlp:
            return no("non-GLU experts")
        if cfg.bias:
            return no("expert biases")
        if type(r).__name__ != "RouterTopK" or r.act_fn not in ROUTER_ACT_FN_MAPPING or r.sequence_parallel_enabled:
            return no(f"router {type(r).__name__} / {getattr(r, 'act_fn', None)}")
        act = ops.moe_tkg.act_id(cfg.hidden_act, cfg.glu_type)
        if act is None:
            return no(f"activation {cfg.hidden_act}")
        ids = list(op.local_expert_ids)
        if ids != list(range(ids[0], ids[0] + len(ids))):
            return no("non-contiguous local experts")
        n = self._norm
        if n is not None and (not hasattr(n, "weight") or not any(hasattr(n, a) for a in ("eps", "variance_epsilon"))):
            return no("norm module without weight / eps")
        T = hidden_states.numel() // hidden_states.shape[-1]
        w_gu, w_dn, rw = op.gate_up_proj.weight, op.down_proj.weight, r.linear_router.weight
        x2 = hidden_states.reshape(T, -1)
        if not ops.moe_tkg.kernel_eligible(x2.contiguous(), rw, w_gu, w_dn, cfg.top_k):
            return no(f"shape / dtype (T={T}, dtype={hidden_states.dtype}, E={cfg.num_experts})")
        if n is not None and n.weight.dtype != torch.bfloat16:
            return no("norm weight dtype")
        return True

    def _moe_fused_tkg_kernel(self, hidden_states: torch.Tensor):
        from ... import ops

        cfg, op, r, n = self._experts.routed_experts_mlp_config, self._mlp_op(), self._router, self._norm
        shape = hidden_states.shape
        x2 = hidden_states.reshape(-1, shape[-1]).contiguous()
        eps = 0.0 if n is None else float(getattr(n, "eps", getattr(n, "variance_epsilon", 1e-6)))
        inf = float("inf")

        def lim(v, d):
            return d if v is None else float(v)

        out, logits, idx, _w = ops.moe_tkg.moe_block_tkg(
            x2, None if n is None else n.weight, r.linear_router.weight, r.linear_router.bias, op.gate_up_proj.weight,
            op.down_proj.weight, int(op.local_expert_ids[0]), cfg.top_k, eps, int(ROUTER_ACT_FN_MAPPING[r.act_fn]),
            bool(r.apply_act_fn_over_topk), bool(cfg.normalize_top_k_affinities), bool(cfg.early_expert_affinity_modulation),
            True, ops.moe_tkg.act_id(cfg.hidden_act, cfg.glu_type), float(cfg.hidden_act_scaling_factor), float(cfg.hidden_act_bias),
            (lim(cfg.gate_clamp_lower_limit, -inf), lim(cfg.gate_clamp_upper_limit, inf), lim(cfg.up_clamp_lower_limit, -inf),
             lim(cfg.up_clamp_upper_limit, inf)))
        return out.view(shape), logits, idx

    def _flat_path(self, x: torch.Tensor):
        h = self._norm(x) if self._norm is not None else x
        logits, aff, idx = self._router(h)
        aff = mappings.copy_to_tensor_model_parallel_region(aff) if torch.is_grad_enabled() else aff
        flat = h.reshape(-1, h.shape[-1])
        em = self._experts
        prev, em._decode_hint = em._decode_hint, True
        try:
            y = em.forward_all_experts(flat, aff, idx).view(h.shape)
        finally:
            em._decode_hint = prev
        return y, logits, idx, h

    def _reduce(self, y: torch.Tensor) -> torch.Tensor:
        """The delayed all-reduce: intermediate shards (TP), then expert shards (EP) — together the reference's world group."""
        y = mappings.reduce_from_tensor_model_parallel_region(y, self.tensor_parallel_group) if self.tensor_parallel_group is not None \
            else mappings.reduce_from_tensor_model_parallel_region(y)
        if ps.model_parallel_is_initialized() and ps.get_expert_model_parallel_size() > 1 and not self.training:
            y = mappings.reduce_from_tensor_model_parallel_region(y, ps.get_expert_model_parallel_group())
        return y

    def forward(self, hidden_states: torch.Tensor, residual: Optional[torch.Tensor] = None):
        """``hidden_states`` [B, S, H] or [S, B, H].  Returns ``(output, [router_logits], [expert_index], [residual])`` —
        ``residual`` (= ``hidden_states + residual``, the stream the next block adds to) only when one was passed in."""
        x = hidden_states if residual is None else hidden_states + residual
        if self._can_use_kernel(x):
            y, logits, idx = self._moe_fused_tkg_kernel(x)
            if self._shared is not None:
                y = y + self._shared(self._norm(x) if self._norm is not None else x)
        else:
            y, logits, idx, h = self._flat_path(x)
            if self._shared is not None:
                y = y + self._shared(h)
        y = self._reduce(y)
        out = (y,)
        if self.return_router_logits:
            out += (logits,)
        if self.return_expert_index:
            out += (idx,)
        if residual is not None:
            out += (x,)
        return out
