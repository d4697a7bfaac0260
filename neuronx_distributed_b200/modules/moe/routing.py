"""Token routers (reference ``modules/moe/routing.py``: ``RouterBase`` :20-152, ``RouterTopK`` :155,
``RouterSinkhorn`` :213, ``GroupLimitedRouter`` :316).  Router math runs in fp32; the tiny ``[T, E]``
GEMM is replicated on every TP rank and its weight is tagged for sequence-parallel gradient all-reduce."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from ...parallel_layers import mappings


class RouterBase(nn.Module):
    def __init__(self, num_experts: int, top_k: int, hidden_size: int, sequence_parallel_enabled: bool = False,
                 sequence_dimension: Optional[int] = None, dtype: torch.dtype = torch.float32, device=None,
                 bias: bool = False, act_fn: str = "softmax", tensor_model_parallel_group=None, jitter_eps: float = 0.0,
                 store_transposed_weights: bool = False, apply_act_fn_over_topk: bool = False):
        super().__init__()
        if not 0 < top_k <= num_experts:
            raise ValueError(f"invalid top_k={top_k} for num_experts={num_experts}")
        self.num_experts, self.top_k, self.hidden_size = num_experts, top_k, hidden_size
        self.sequence_parallel_enabled = sequence_parallel_enabled
        self.sequence_dimension = 0 if sequence_dimension is None else sequence_dimension
        self.act_fn, self.jitter_eps = act_fn, jitter_eps
        self.apply_act_fn_over_topk = apply_act_fn_over_topk
        self.tensor_parallel_group = tensor_model_parallel_group
        self.linear_router = nn.Linear(hidden_size, num_experts, bias=bias, dtype=dtype, device=device)
        for p in self.linear_router.parameters():
            setattr(p, "sequence_parallel_enabled", sequence_parallel_enabled)
        self.bias, self.store_transposed_weights = bias, store_transposed_weights
        if store_transposed_weights:
            # inference: a second copy ``[H, E]`` so the decode-time router GEMV streams the weight along its contiguous dim
            # (reference routing.py:96-104, filled from the checkpoint by ``preshard_hook``)
            self.weight_T = nn.Parameter(self.linear_router.weight.detach().t().contiguous().clone(), requires_grad=False)

    def preshard_hook(self, model_state_dict, prefix: str) -> None:
        """Full checkpoints carry only ``linear_router.weight``: derive ``weight_T`` from it."""
        if not self.store_transposed_weights:
            return
        base = prefix
        for suffix in ("linear_router.weight", "weight_T", "weight"):
            if base.endswith(suffix):
                base = base[: -len(suffix)]
                break
        src = base + "linear_router.weight"
        if src in model_state_dict:
            model_state_dict[base + "weight_T"] = model_state_dict[src].detach().t().contiguous().clone()

    def get_router_logits(self, hidden_states: torch.Tensor) -> torch.Tensor:
        if self.sequence_parallel_enabled:
            hidden_states = mappings.gather_from_sequence_parallel_region(
                hidden_states, self.sequence_dimension, to_model_parallel=False, process_group=self.tensor_parallel_group)
        x = hidden_states.reshape(-1, self.hidden_size)
        if self.training and self.jitter_eps > 0:
            x = x * torch.empty_like(x).uniform_(1.0 - self.jitter_eps, 1.0 + self.jitter_eps)
        if self.store_transposed_weights and not self.training:
            out = x.to(self.weight_T.dtype) @ self.weight_T
            return (out if self.linear_router.bias is None else out + self.linear_router.bias).float()
        w = self.linear_router.weight
        return F.linear(x.to(w.dtype), w, self.linear_router.bias).float()

    def apply_activation_fn(self, weights: torch.Tensor) -> torch.Tensor:
        logits = weights      # reference parameter names in the signature
        if self.act_fn == "softmax":
            return torch.softmax(logits, dim=-1, dtype=torch.float32)
        if self.act_fn == "sigmoid":
            return torch.sigmoid(logits.float())
        raise ValueError(f"unknown router activation {self.act_fn}")


class RouterTopK(RouterBase):
    """Standard top-k routing: returns ``(router_logits [T,E], expert_affinities [T,E], expert_index [T,k])``."""

    def forward(self, hidden_states: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        logits = self.get_router_logits(hidden_states)
        if self.apply_act_fn_over_topk:
            top_logits, idx = torch.topk(logits, self.top_k, dim=-1)
            aff_k = self.apply_activation_fn(top_logits)
            aff = torch.zeros_like(logits).scatter_(1, idx, aff_k)
        else:
            aff = self.apply_activation_fn(logits)
            _, idx = torch.topk(aff, self.top_k, dim=-1)
        return logits, aff.to(hidden_states.dtype), idx.detach().long()


class RouterSinkhorn(RouterBase):
    DEFAULT_SINKHORN_ITERS = 30

    """Top-1 routing balanced with Sinkhorn iterations during training (Megatron-style)."""

    DEFAULT_SINKHORN_ITERS = 30

    def __init__(self, *a, sinkhorn_iterations: Optional[int] = None, sinkhorn_tol: Optional[float] = None, **k):
        k.setdefault("act_fn", "sigmoid")
        super().__init__(*a, **k)
        if self.top_k != 1:
            raise NotImplementedError("RouterSinkhorn supports top_k=1 only")
        self.sinkhorn_iterations = sinkhorn_iterations if sinkhorn_iterations is not None else self.DEFAULT_SINKHORN_ITERS
        self.sinkhorn_tol = sinkhorn_tol

    @staticmethod
    def _sinkhorn(cost: torch.Tensor, iters: int, tol: Optional[float]) -> torch.Tensor:
        cost = torch.exp(cost.float())
        d0 = torch.ones(cost.shape[0], device=cost.device)
        d1 = torch.ones(cost.shape[1], device=cost.device)
        eps = 1e-8
        for _ in range(iters):
            d0_new = 1.0 / (d0.shape[0] * ((d1.unsqueeze(0) * cost).sum(1) + eps))
            d1 = 1.0 / (d1.shape[0] * ((d0_new.unsqueeze(1) * cost).sum(0) + eps))
            if tol is not None and (d0_new - d0).abs().mean() < tol:
                d0 = d0_new
                break
            d0 = d0_new
        return d1 * cost * d0.unsqueeze(1)

    def forward(self, hidden_states: torch.Tensor):
        logits = self.get_router_logits(hidden_states)
        aff = self.apply_activation_fn(logits)
        if self.training:
            with torch.no_grad():
                idx = self._sinkhorn(logits, self.sinkhorn_iterations, self.sinkhorn_tol).argmax(-1, keepdim=True)
        else:
            idx = aff.argmax(-1, keepdim=True)
        return logits, aff.to(hidden_states.dtype), idx.long()


class GroupLimitedRouter(RouterBase):
    """DeepSeek-V3 style: experts are split into ``n_group`` groups; each token may only pick experts from its
    ``topk_group`` best groups (group score = sum of the group's top-2 biased scores)."""

    def __init__(self, *a, n_group: int = 1, topk_group: int = 1, routed_scaling_factor: float = 1.0, **k):
        k.setdefault("act_fn", "sigmoid")
        super().__init__(*a, **k)
        assert self.num_experts % n_group == 0
        self.n_group, self.topk_group, self.routed_scaling_factor = n_group, topk_group, routed_scaling_factor
        self.e_score_correction_bias = nn.Parameter(torch.zeros(self.num_experts, dtype=torch.float32), requires_grad=False)

    def noaux_tc_top_k(self, scores: torch.Tensor):
        """Auxiliary-loss-free top-k (reference routing.py:391-413): add the learned per-expert correction bias (selection
        only), score every group by the sum of its two best experts, keep the ``topk_group`` best groups, pick the ``top_k``
        experts among them.  Returns ``(topk_idx, scores)`` — the affinities stay the UNbiased scores."""
        biased = scores + self.e_score_correction_bias.unsqueeze(0)
        T = biased.shape[0]
        g = biased.view(T, self.n_group, -1)
        group_scores = g.topk(min(2, g.shape[-1]), dim=-1).values.sum(-1)
        gidx = group_scores.topk(self.topk_group, dim=-1).indices
        gmask = torch.zeros_like(group_scores).scatter_(1, gidx, 1.0)
        emask = gmask.unsqueeze(-1).expand(T, self.n_group, g.shape[-1]).reshape(T, -1)
        masked = biased.masked_fill(emask == 0, float("-inf"))
        return masked.topk(self.top_k, dim=-1).indices, scores

    def forward(self, hidden_states: torch.Tensor):
        logits = self.get_router_logits(hidden_states)
        scores = self.apply_activation_fn(logits)
        idx, _ = self.noaux_tc_top_k(scores)
        aff = scores * self.routed_scaling_factor
        return logits, aff.to(hidden_states.dtype), idx.long()
