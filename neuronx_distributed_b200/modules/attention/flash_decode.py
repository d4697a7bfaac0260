"""Distributed flash-decoding: the KV cache of a GQA head that is replicated over several TP ranks is *sharded along the sequence*
inside that KV-replica group instead of being stored ``n`` times (reference ``examples/inference/modules/attention/
flashdecode_attention.py:190-235``, masks ``flashdecode/util.py:9-77``, cache shape ``kv_cache_manager.py:60-88``).

Per decode step, inside a group of ``n`` ranks that share the same KV heads:
  1. all-gather the queries of the group (every rank attends for all ``n·H_local`` query heads over ITS sequence shard),
  2. local partial attention → un-normalised output, row max ``m`` and row sum ``l`` (fp32),
  3. combine across the group with the log-sum-exp rule: all-reduce(max) of ``m``, rescale, reduce-scatter(sum) of the
     outputs and of ``l`` over the head axis — each rank ends with the finished output of its own heads.
Rank ``r`` of the group owns cache positions ``[r·L_local, (r+1)·L_local)``; the new token's K/V is written only by its owner.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def owner_and_offset(positions: torch.Tensor, l_local: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Global cache position → (owning rank in the group, offset inside that rank's shard)."""
    return positions // l_local, positions % l_local


def write_decode_sharded(k_cache: torch.Tensor, v_cache: torch.Tensor, k_new: torch.Tensor, v_new: torch.Tensor,
                         positions: torch.Tensor, rank_in_group: int) -> None:
    """k/v_cache ``[B, L_local, Hkv, D]`` (this rank's shard); k/v_new ``[B, 1, Hkv, D]``; only the owner writes (mask, no sync)."""
    l_local = k_cache.shape[1]
    owner, off = owner_and_offset(positions, l_local)
    mine = (owner == rank_in_group)
    b = torch.arange(k_cache.shape[0], device=k_cache.device)
    off = torch.where(mine, off, torch.zeros_like(off))
    keep = mine.view(-1, 1, 1).to(k_cache.dtype)
    k_cache[b, off] = keep * k_new[:, 0] + (1 - keep) * k_cache[b, off]
    v_cache[b, off] = keep * v_new[:, 0] + (1 - keep) * v_cache[b, off]


def flash_decode_attention(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, positions: torch.Tensor,
                           group, scale: Optional[float] = None) -> torch.Tensor:
    """q ``[B, 1, H_local, D]`` (this rank's query heads), caches ``[B, L_local, Hkv, D]`` (this rank's sequence shard of the
    shared KV heads), ``positions`` ``[B]`` = index of the newest token (attend to ≤ position).  Returns ``[B, 1, H_local, D]``."""
    n = dist.get_world_size(group)
    r = dist.get_rank(group)
    B, _, Hl, D = q.shape
    l_local, Hkv = k_cache.shape[1], k_cache.shape[2]
    scale = scale if scale is not None else 1.0 / math.sqrt(D)
    # 1. every rank needs the queries of the whole group
    if n > 1:
        parts = [torch.empty_like(q) for _ in range(n)]
        dist.all_gather(parts, q.contiguous(), group=group)
        q_all = torch.cat(parts, dim=2)                                  # [B, 1, n·Hl, D]
    else:
        q_all = q
    H = q_all.shape[2]
    # 2. partial attention over the local shard (fp32), masked to the positions this rank owns that are ≤ position
    g = H // Hkv
    kf = k_cache.float().repeat_interleave(g, dim=2)                    # [B, L_local, H, D]
    vf = v_cache.float().repeat_interleave(g, dim=2)
    s = torch.einsum("bhd,blhd->bhl", q_all[:, 0].float(), kf) * scale  # [B, H, L_local]
    gpos = r * l_local + torch.arange(l_local, device=q.device)
    valid = gpos[None, :] <= positions[:, None]                         # [B, L_local]
    s = s.masked_fill(~valid[:, None, :], float("-inf"))
    m = s.max(dim=-1).values                                            # [B, H]  (−inf if this shard has nothing yet)
    p = torch.exp(s - torch.where(torch.isinf(m), torch.zeros_like(m), m).unsqueeze(-1))
    p = torch.where(valid[:, None, :], p, torch.zeros_like(p))
    l = p.sum(-1)                                                       # [B, H]
    o = torch.einsum("bhl,blhd->bhd", p, vf)                            # un-normalised [B, H, D]
    # 3. combine across the group
    if n > 1:
        m_glob = m.clone()
        dist.all_reduce(m_glob, op=dist.ReduceOp.MAX, group=group)
        w = torch.where(torch.isinf(m), torch.zeros_like(m), torch.exp(m - m_glob))
        o, l = o * w.unsqueeze(-1), l * w
        o_parts = list(o.view(B, n, Hl, D).unbind(1))
        l_parts = list(l.view(B, n, Hl).unbind(1))
        o_mine, l_mine = torch.empty_like(o_parts[0]), torch.empty_like(l_parts[0])
        if dist.get_backend(group) == "gloo":                           # gloo has no reduce_scatter
            o_sum, l_sum = o.clone(), l.clone()
            dist.all_reduce(o_sum, group=group)
            dist.all_reduce(l_sum, group=group)
            o_mine, l_mine = o_sum.view(B, n, Hl, D)[:, r], l_sum.view(B, n, Hl)[:, r]
        else:
            dist.reduce_scatter(o_mine, [t.contiguous() for t in o_parts], group=group)
            dist.reduce_scatter(l_mine, [t.contiguous() for t in l_parts], group=group)
    else:
        o_mine, l_mine = o, l
    return (o_mine / l_mine.clamp(min=1e-30).unsqueeze(-1)).to(q.dtype).unsqueeze(1)
