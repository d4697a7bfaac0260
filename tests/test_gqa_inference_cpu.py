"""Inference GQA sharding for head counts that do not divide TP: REPLICATE_TO_TP_DEGREE and CONVERT_TO_MHA layouts reproduce the
dense attention block from an HF-style checkpoint (preshard hooks → per-rank shards)."""
import pytest
import torch
from torch import nn

from dist_utils import run_distributed


def test_head_layout_slot_maps():
    from neuronx_distributed_b200.inference.gqa import GQA, HeadLayout, determine_sharding_strategy, get_shardable_head_counts, remap_heads

    assert determine_sharding_strategy(32, 8) == GQA.REPLICATE_TO_TP_DEGREE
    assert determine_sharding_strategy(4, 3) == GQA.CONVERT_TO_MHA
    assert determine_sharding_strategy(32, 8, GQA.CONVERT_TO_MHA) == GQA.CONVERT_TO_MHA
    assert get_shardable_head_counts(32, 56, 8, GQA.REPLICATE_TO_TP_DEGREE) == (64, 32)
    assert get_shardable_head_counts(32, 56, 8, GQA.CONVERT_TO_MHA) == (64, 64)
    assert get_shardable_head_counts(8, 32, 32, GQA.REPLICATE_TO_TP_DEGREE) == (32, 32)
    L = HeadLayout.build(4, 6, 2)
    assert (L.q, L.kv) == (8, 4)
    assert L.q_slots() == [0, 1, 2, -1, 3, 4, 5, -1] and L.kv_slots() == [0, 0, 1, 1]
    M = HeadLayout.build(4, 6, 3)
    assert M.strategy == GQA.CONVERT_TO_MHA and (M.q, M.kv) == (8, 8)
    assert M.q_slots() == [0, 1, 2, 3, 4, 5, -1, -1] and M.kv_slots() == [0, 0, 1, 1, 2, 2, -1, -1]
    w = torch.arange(6 * 2 * 3, dtype=torch.float32).view(12, 3)                # 6 heads of 2 rows
    r = remap_heads(w, L.q_slots(), 6, 0)
    assert r.shape == (16, 3) and torch.equal(r[0:6], w[0:6]) and torch.equal(r[6:8], torch.zeros(2, 3)) and torch.equal(r[8:14], w[6:12])
    c = remap_heads(w.t().contiguous(), L.q_slots(), 6, 1)
    assert torch.equal(c, r.t())


class _Attn(nn.Module):
    def __init__(self, hidden, D, nq, nkv, tp, fused):
        super().__init__()
        from neuronx_distributed_b200.inference.gqa import GroupQueryAttention_O, GroupQueryAttention_QKV

        self.qkv_proj = GroupQueryAttention_QKV(hidden, D, nq, nkv, tp_degree=tp, gather_output=False, fused_qkv=fused)
        self.o_proj = GroupQueryAttention_O(hidden, D, nq, nkv, tp_degree=tp, input_is_parallel=True)
        self.D = D

    def forward(self, x):                                   # x [T, hidden]
        q, k, v = self.qkv_proj(x)
        T, D = x.shape[0], self.D
        q, k, v = q.view(T, -1, D), k.view(T, -1, D), v.view(T, -1, D)
        rep = q.shape[1] // k.shape[1]
        k, v = k.repeat_interleave(rep, 1), v.repeat_interleave(rep, 1)
        p = torch.softmax(torch.einsum("thd,shd->hts", q, k) / D ** 0.5, dim=-1)
        o = torch.einsum("hts,shd->thd", p, v).reshape(T, -1)
        return self.o_proj(o)


def _dense(x, sd, D, nq, nkv):
    T = x.shape[0]
    q = (x @ sd["attn.q_proj.weight"].t()).view(T, nq, D)
    k = (x @ sd["attn.k_proj.weight"].t()).view(T, nkv, D).repeat_interleave(nq // nkv, 1)
    v = (x @ sd["attn.v_proj.weight"].t()).view(T, nkv, D).repeat_interleave(nq // nkv, 1)
    p = torch.softmax(torch.einsum("thd,shd->hts", q, k) / D ** 0.5, dim=-1)
    o = torch.einsum("hts,shd->thd", p, v).reshape(T, -1)
    return o @ sd["attn.o_proj.weight"].t()


def _gqa_block(rank, world, nq, nkv, fused):
    from neuronx_distributed_b200.inference.sharding import shard_state_dict_for_rank
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    hidden, D = 32, 8
    g = torch.Generator().manual_seed(0)
    sd = {"attn.q_proj.weight": torch.randn(nq * D, hidden, generator=g) * 0.3, "attn.k_proj.weight": torch.randn(nkv * D, hidden, generator=g) * 0.3,
          "attn.v_proj.weight": torch.randn(nkv * D, hidden, generator=g) * 0.3, "attn.o_proj.weight": torch.randn(hidden, nq * D, generator=g) * 0.3}
    model = nn.ModuleDict({"attn": _Attn(hidden, D, nq, nkv, world, fused)})
    local = shard_state_dict_for_rank(model, dict(sd), rank, world)
    missing, unexpected = model.load_state_dict(local, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    x = torch.randn(5, hidden, generator=torch.Generator().manual_seed(1))
    got = model["attn"](x)
    torch.testing.assert_close(got, _dense(x, sd, D, nq, nkv), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("nq,nkv,fused", [(6, 2, False), (6, 2, True), (6, 3, False), (8, 8, False), (12, 1, False)])
def test_gqa_padded_replicated_block_matches_dense_tp4(nq, nkv, fused):
    run_distributed(_gqa_block, 4, nq, nkv, fused, timeout=180)
