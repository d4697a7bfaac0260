"""Blockwise (dropless) expert MLP.

Role of the reference's NKI blockwise kernels K6/K7/K9 (``modules/moe/blockwise.py:180-468,1037-1127``,
``expert_mlps_v2.py:691-917,1079-1206``).  Tokens are grouped by expert, every expert's group is padded to a multiple
of ``block_size`` and each block is multiplied by exactly one expert's weights:

    num_blocks            = ceil((T·k − (E−1)) / B) + (E−1)          (static upper bound)
    block_to_expert[b]    : expert owning block b
    token_position_to_id  : flat [num_blocks·B] token id feeding each block slot (−1 = padding)

On CUDA the two projections are grouped tcgen05 GEMMs (``ops.gemm.grouped_matmul`` → MODE 3/4 of
``csrc/gemm_sm100.cu``): the block's expert only shifts the TMA coordinate of the weight tile, the backward runs the
grouped dgrad and a per-expert segmented wgrad, and nothing syncs with the host.  On CPU it is a gathered einsum.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch


def get_num_blocks(total_tokens: int, top_k: int, num_experts: int, block_size: int) -> int:
    return max(1, math.ceil(max(total_tokens * top_k - (num_experts - 1), 0) / block_size) + (num_experts - 1))


def build_block_metadata(expert_index: torch.Tensor, num_experts: int, block_size: int
                         ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """From ``expert_index [T, k]`` build ``(block_to_expert [nb], token_position_to_id [nb·B], tokens_per_expert [E])``.
    Entirely on device (sort + cumsum), no host sync."""
    T, k = expert_index.shape
    nb = get_num_blocks(T, k, num_experts, block_size)
    flat_e = expert_index.reshape(-1)
    token_ids = torch.arange(T, device=expert_index.device).repeat_interleave(k)
    order = torch.argsort(flat_e, stable=True)
    se, st = flat_e[order], token_ids[order]
    counts = torch.bincount(flat_e, minlength=num_experts)
    blocks_per_e = (counts + block_size - 1) // block_size
    block_start = torch.cumsum(blocks_per_e, 0) - blocks_per_e                      # first block of each expert
    seg_start = torch.cumsum(counts, 0) - counts
    pos_in_e = torch.arange(se.numel(), device=se.device) - seg_start[se]
    slot = block_start[se] * block_size + pos_in_e
    tp2id = torch.full((nb * block_size,), -1, dtype=torch.long, device=se.device)
    tp2id[slot] = st
    b_idx = torch.arange(nb, device=se.device)
    ends = torch.cumsum(blocks_per_e, 0)
    block_to_expert = torch.searchsorted(ends, b_idx, right=True).clamp(max=num_experts - 1)
    return block_to_expert, tp2id, counts


def blockwise_expert_mlp(hidden: torch.Tensor, expert_affinities: torch.Tensor, expert_index: torch.Tensor,
                         experts, block_size: int, normalize: bool = False) -> torch.Tensor:
    """Dropless MoE MLP: ``out[t] = Σ_j aff[t, e_j] · MLP_{e_j}(hidden[t])``."""
    T, H = hidden.shape
    E = expert_affinities.shape[-1]
    from ... import ops

    b2e, tp2id, counts = build_block_metadata(expert_index, E, block_size)
    nb = b2e.numel()
    blocks_per_e = (counts + block_size - 1) // block_size
    seg_first = torch.cat([blocks_per_e.new_zeros(1), torch.cumsum(blocks_per_e, 0)])          # [E+1] block prefix sum
    valid = tp2id >= 0
    gather_ids = tp2id.clamp(min=0)
    x = hidden[gather_ids] * valid.unsqueeze(-1).to(hidden.dtype)                   # [nb·B, H]
    proj = experts.gate_up_proj if experts.glu_mlp else experts.up_proj
    w1, w2 = proj.weight, experts.down_proj.weight
    h = ops.gemm.grouped_matmul(x, w1, b2e, seg_first, block_size)
    h = experts.activation(h)
    y = ops.gemm.grouped_matmul(h.contiguous(), w2, b2e, seg_first, block_size)
    aff = expert_affinities[gather_ids, b2e.repeat_interleave(block_size)] * valid.to(expert_affinities.dtype)
    if normalize:
        denom = expert_affinities.gather(1, expert_index).sum(-1, keepdim=True).clamp(min=1e-9)
        aff = aff / denom[gather_ids, 0]
    y = y * aff.unsqueeze(-1).to(y.dtype)
    out = torch.zeros(T, H, dtype=y.dtype, device=y.device)
    out.index_add_(0, gather_ids, y * valid.unsqueeze(-1).to(y.dtype))
    return out
