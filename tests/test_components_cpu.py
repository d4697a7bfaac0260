"""Unit tests of the model-building components that the bigger model tests only touch indirectly: GQA-QKV with KV replication,
input-channel parallel conv, pad_model, RNG tracker, grad norm / clip, DistributedLogprob, the extra routers, shared experts,
token shuffling."""

import pytest
import torch
from torch import nn

from dist_utils import run_distributed


def _gqa(rank, world):
    """kv_heads (1) < tp (2): KV weights are replicated ×2; q/k/v equal the un-sharded projection; dK/dV weight grads are summed
    over the KV-shared group so both replicas carry the full gradient; the input gradient equals the dense model's."""
    from neuronx_distributed_b200.modules.qkv_linear import GQAQKVColumnParallelLinear
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    H, D, nq, nkv = 32, 8, 4, 1
    for fuse in (True, False):
        torch.manual_seed(0)
        lin = GQAQKVColumnParallelLinear(H, [nq * D, nkv * D], bias=False, gather_output=False, kv_size_multiplier=world,
                                         fuse_qkv=fuse, keep_master_weight=True, head_dim=D)
        mw = lin.master_weights
        wq, wk, wv = (mw[n].float() for n in ("q", "k", "v"))            # full (replicated for k/v) fp32 masters
        x = torch.randn(6, 2, H, generator=torch.Generator().manual_seed(1), requires_grad=True)
        q, k, v = lin(x)
        assert q.shape[-1] == nq * D // world and k.shape[-1] == nkv * D and v.shape[-1] == nkv * D
        torch.testing.assert_close(q, (x @ wq.t())[..., rank * nq * D // world:(rank + 1) * nq * D // world], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(k, x @ wk[: nkv * D].t(), rtol=1e-5, atol=1e-5)     # every replica = the single kv head
        torch.testing.assert_close(v, x @ wv[: nkv * D].t(), rtol=1e-5, atol=1e-5)
        (q.sum() + (k * 2).sum() + (v * 3).sum()).backward()
        gk = (lin.weight_qkv.grad[nq * D // world: nq * D // world + nkv * D] if fuse else lin.weight_k.grad)
        # d/dWk of 2·sum(k) summed over the 2 replicas (KV-shared group all-reduce) = 2 · 2 · Σ x
        want = 2.0 * world * x.detach().sum((0, 1)).expand(nkv * D, H)
        torch.testing.assert_close(gk, want, rtol=1e-4, atol=1e-4)
        # dL/dx equals the DENSE model's: the K/V head feeds `world` per-rank loss terms, i.e. L = Σq + world·(2Σk + 3Σv).
        # (Before the 1/multiplier correction of the dgrad path the replicated K/V contribution was counted world times.)
        dense_gx = (wq.sum(0) + world * (2.0 * wk[: nkv * D].sum(0) + 3.0 * wv[: nkv * D].sum(0))).expand_as(x)
        torch.testing.assert_close(x.grad, dense_gx, rtol=1e-4, atol=1e-4)
        ps.destroy_model_parallel(); ps.initialize_model_parallel(tensor_model_parallel_size=world)


def test_gqa_qkv_kv_replication_tp2():
    run_distributed(_gqa, 2, timeout=120)


def _conv_pad_rng(rank, world):
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.parallel_layers import random as prandom
    from neuronx_distributed_b200.parallel_layers.layers import InputChannelParallelConv2d, OutputChannelParallelConv2d
    from neuronx_distributed_b200.parallel_layers.pad import generate_padding_mask, get_number_of_extra_heads

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    # output-channel conv (not gathered) feeding an input-channel conv (input_is_parallel) == the dense two-conv stack
    torch.manual_seed(0)
    c1 = OutputChannelParallelConv2d(3, 8, 3, padding=1, gather_output=False, keep_master_weight=True)
    c2 = InputChannelParallelConv2d(8, 4, 3, padding=1, input_is_parallel=True, keep_master_weight=True)
    x = torch.randn(2, 3, 6, 6, generator=torch.Generator().manual_seed(2))
    y = c2(torch.relu(c1(x)))
    b1 = torch.cat([t for t in _all_gather_cat(c1.bias.detach(), world)])
    ref = nn.functional.conv2d(torch.relu(nn.functional.conv2d(x, c1.master_weight, b1, padding=1)), c2.master_weight, c2.bias, padding=1)
    torch.testing.assert_close(y, ref, rtol=1e-4, atol=1e-4)
    # head padding helpers
    assert get_number_of_extra_heads(6, 4) == 2 and get_number_of_extra_heads(8, 4) == 0
    from neuronx_distributed_b200.parallel_layers.pad import generate_global_padding_masks

    q_mask, kv_mask = generate_global_padding_masks(6, 8, 2, 2, 4)
    assert q_mask.tolist() == [True] * 6 + [False] * 2 and kv_mask.tolist() == [True, True]
    # per-rank mask (the reference's docstring example): 48 query heads padded to 64, 8 KV heads replicated 4x, TP=32 →
    # 2 local heads; the 6 query heads of a KV head fill replicas 0..2, replica 3 holds only padding
    for layout, rank_of in (("tile", lambda kv, rep: rep * 8 + kv), ("adjacent", lambda kv, rep: kv * 4 + rep)):
        real = 0
        for kv in range(8):
            got = [generate_padding_mask(48, 64, 8, 32, rank_of(kv, rep), kv_layout=layout).tolist() for rep in range(4)]
            assert got == [[True, True]] * 3 + [[False, False]], (layout, kv, got)
            real += sum(sum(g) for g in got)
        assert real == 48
    assert generate_padding_mask(48, 64, 8, 32, 31, hardware_type="trn2").tolist() == [False, False]
    assert generate_padding_mask(48, 64, 8, 32, 7, hardware_type="trn1").tolist() == [True, True]
    # RNG tracker: default generator identical inside the TP group, the model-parallel stream differs per tp rank
    prandom.model_parallel_manual_seed(123)
    a = torch.rand(4)
    with prandom.get_rng_tracker().fork():
        b = torch.rand(4)
    ga, gb = _all_gather_cat(a, world), _all_gather_cat(b, world)
    assert torch.equal(ga[0], ga[1]) and not torch.equal(gb[0], gb[1])


def _all_gather_cat(t, world):
    import torch.distributed as dist

    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t.contiguous())
    return out


def test_parallel_conv_stack_padding_helpers_rng_tp2():
    run_distributed(_conv_pad_rng, 2, timeout=120)


def _grads_logprob(rank, world):
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.parallel_layers.grads import clip_grad_norm, get_grad_norm
    from neuronx_distributed_b200.parallel_layers.layers import ColumnParallelLinear
    from neuronx_distributed_b200.parallel_layers.loss_functions import from_parallel_logits_to_logprobs

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    torch.manual_seed(0)
    lin = ColumnParallelLinear(8, 12, bias=False, gather_output=False, keep_master_weight=True)
    norm_w = nn.Parameter(torch.ones(8))                       # a replicated (non-TP) parameter
    x = torch.randn(5, 8, generator=torch.Generator().manual_seed(3))
    (lin(x * norm_w).pow(2).sum()).backward()
    # the Column layer's backward all-reduces dgrad over TP, so the replicated parameter already holds its full gradient on
    # every rank; get_grad_norm must count it once
    total = get_grad_norm([lin.weight, norm_w], norm_type=2)
    wf = lin.master_weight.clone().requires_grad_(True); nf = torch.ones(8, requires_grad=True)
    ((x * nf) @ wf.t()).pow(2).sum().backward()
    want = torch.sqrt(wf.grad.pow(2).sum() + nf.grad.pow(2).sum())
    torch.testing.assert_close(total, want, rtol=1e-4, atol=1e-4)
    before = lin.weight.grad.clone()
    clip_grad_norm([lin.weight, norm_w], max_norm=float(want) / 2)
    torch.testing.assert_close(lin.weight.grad, before * 0.5, rtol=1e-3, atol=1e-5)
    # vocab-parallel log-probs of the NEXT token (the helper shifts the targets, reference loss_functions.py:206-215)
    logits_full = torch.randn(4, 3, 16, generator=torch.Generator().manual_seed(5))
    tgt = torch.randint(0, 16, (4, 3), generator=torch.Generator().manual_seed(6))
    local = logits_full.chunk(world, -1)[rank].clone().requires_grad_(True)
    with pytest.raises(RuntimeError, match="inference=False"):            # the reference's default: scoring only
        from_parallel_logits_to_logprobs(local, tgt).sum().backward()
    lp = from_parallel_logits_to_logprobs(local, tgt, inference=False)
    ref = torch.log_softmax(logits_full, -1)[:, :-1].gather(-1, tgt[:, 1:].unsqueeze(-1)).squeeze(-1)
    torch.testing.assert_close(lp, ref, rtol=1e-5, atol=1e-5)
    lp.sum().backward()
    assert local.grad is not None and torch.isfinite(local.grad).all()


def test_grad_norm_clip_and_distributed_logprob_tp2():
    run_distributed(_grads_logprob, 2, timeout=120)


def _moe_parts(rank, world):
    from neuronx_distributed_b200.modules.moe import SharedExperts
    from neuronx_distributed_b200.modules.moe.routing import GroupLimitedRouter, RouterSinkhorn
    from neuronx_distributed_b200.parallel_layers import mappings
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    torch.manual_seed(0)
    T, H, E = 32, 16, 8
    x = torch.randn(T, H, generator=torch.Generator().manual_seed(1))
    # Sinkhorn router (training): top-1, assignments spread over more experts than plain argmax of the same logits
    r = RouterSinkhorn(E, 1, H).train()
    logits, aff, idx = r(x)
    assert idx.shape == (T, 1) and aff.shape == (T, E)
    assert idx.unique().numel() >= logits.argmax(-1).unique().numel()
    # group-limited router: every chosen expert lies inside the token's top `topk_group` groups
    g = GroupLimitedRouter(E, 2, H, n_group=4, topk_group=2)
    _, aff2, idx2 = g(x)
    scores = torch.sigmoid(g.get_router_logits(x).float())
    gs = scores.view(T, 4, 2).topk(2, -1).values.sum(-1)
    allowed = gs.topk(2, -1).indices
    assert all(int(e) // 2 in allowed[t].tolist() for t in range(T) for e in idx2[t])
    # shared experts: partial output summed over TP == dense SwiGLU MLP with the gathered weights
    torch.manual_seed(4)
    se = SharedExperts(H, 12, num_shared_experts=2, fused_gate_up_projection=True)
    y = mappings.reduce_from_tensor_model_parallel_region(se(x))
    import torch.distributed as dist
    gu = _all_gather_cat(se.gate_up_proj.weight.detach(), world); dn = _all_gather_cat(se.down_proj.weight.detach(), world)
    ref = 0
    for w_gu, w_dn in zip(gu, dn):
        gte, up = (x @ w_gu.t()).chunk(2, -1)
        ref = ref + (nn.functional.silu(gte) * up) @ w_dn.t()
    torch.testing.assert_close(y, ref, rtol=1e-4, atol=1e-4)
    # separate gate / up projections (the reference's default key names), plain and stored-transposed: load ONE full
    # [out, in] checkpoint into both through the sharder (preshard hook transposes) → identical outputs
    from neuronx_distributed_b200.inference.sharding import shard_state_dict_for_rank
    from neuronx_distributed_b200.modules.moe.shared_experts import ColumnParallelLinearTransposed, RowParallelLinearTransposed

    g = torch.Generator().manual_seed(11)
    full = {"gate_proj.weight": torch.randn(24, H, generator=g) * 0.3, "up_proj.weight": torch.randn(24, H, generator=g) * 0.3,
            "down_proj.weight": torch.randn(H, 24, generator=g) * 0.3}
    dense = (nn.functional.silu(x @ full["gate_proj.weight"].t()) * (x @ full["up_proj.weight"].t())) @ full["down_proj.weight"].t()
    outs = []
    for transposed in (False, True):
        se2 = SharedExperts(H, 12, num_shared_experts=2, transpose_weights=transposed).eval()
        assert set(dict(se2.named_parameters())) == {"gate_proj.weight", "up_proj.weight", "down_proj.weight"}
        if transposed:
            assert isinstance(se2.gate_proj, ColumnParallelLinearTransposed) and isinstance(se2.down_proj, RowParallelLinearTransposed)
            assert se2.gate_proj.weight.shape == (H, 24 // world) and se2.gate_proj.weight.partition_dim == 1
            assert se2.down_proj.weight.shape == (24 // world, H) and se2.down_proj.weight.partition_dim == 0
        se2.load_state_dict(shard_state_dict_for_rank(se2, full, rank, world))
        outs.append(mappings.reduce_from_tensor_model_parallel_region(se2(x, seq_len=T)))
        torch.testing.assert_close(outs[-1], dense, rtol=1e-4, atol=1e-4)
    # backward through the transposed layers
    se2.train()
    xg = x.clone().requires_grad_(True)
    mappings.reduce_from_tensor_model_parallel_region(se2(xg)).sum().backward()
    assert xg.grad is not None and se2.gate_proj.weight.grad.shape == se2.gate_proj.weight.shape
    # sequence-parallel mode: replicated weights; prefill needs no collective, decode slices the replicated weight by rank
    torch.manual_seed(5)
    for fused in (False, True):
        se3 = SharedExperts(H, 12, num_shared_experts=2, sequence_parallel_enabled=True, fused_gate_up_projection=fused).eval()
        names = ("gate_up_proj", "down_proj") if fused else ("gate_proj", "up_proj", "down_proj")
        for n in names:                                                       # replicated: make every rank hold rank 0's copy
            dist.broadcast(getattr(se3, n).weight.data, 0)
        assert se3.down_proj.weight.shape == (H, 24)                          # full, not sharded
        full_out = se3(x, seq_len=T)                                          # prefill: complete result locally
        tok = x[:1]
        part = se3(tok, seq_len=1)                                            # decode: this rank's partial sum
        torch.testing.assert_close(mappings.reduce_from_tensor_model_parallel_region(part), full_out[:1], rtol=1e-4, atol=1e-4)


def test_routers_and_shared_experts_tp2():
    run_distributed(_moe_parts, 2, timeout=120)


def _shuffle(rank, world):
    """Token shuffling over a DP sub-group: tokens really move between ranks, unshuffle is the exact inverse."""
    import torch.distributed as dist

    from neuronx_distributed_b200.modules.moe.token_shuffling import token_shuffle, token_unshuffle
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=1)          # dp = world
    ps.initialize_token_shuffle_group(world)
    x = torch.arange(16, dtype=torch.float32).view(8, 2) + 100 * rank
    xs, perm = token_shuffle(x, seed=7 + rank)
    assert xs.shape == x.shape
    got = [torch.empty_like(xs) for _ in range(world)]
    dist.all_gather(got, xs)
    assert (xs >= 100 * (1 - rank)).any() if rank == 0 else (xs < 100).any()      # received tokens of the other rank
    allx = torch.cat(got)
    assert sorted(allx[:, 0].tolist()) == sorted(torch.cat([torch.arange(0, 16, 2.0), torch.arange(0, 16, 2.0) + 100]).tolist())
    torch.testing.assert_close(token_unshuffle(xs, perm), x)


def test_token_shuffle_roundtrip_dp2():
    run_distributed(_shuffle, 2, timeout=60)
