"""More component-level tests: SPMDRank, post-partition hooks, pad_model, cumsum, LoRA on the GQA-QKV layer, the v0 trace API,
NxDParallelState (single-process build for any rank), ZeRO-1 DCP optimizer checkpoints, the EP-aware ZeRO-1 optimizer."""
import os

import torch
from torch import nn

from dist_utils import run_distributed


def _spmd_pad_lora(rank, world, tmp):
    from neuronx_distributed_b200.modules.lora import LoraConfig, get_lora_model
    from neuronx_distributed_b200.modules.qkv_linear import GQAQKVColumnParallelLinear
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.parallel_layers.layers import ColumnParallelLinear, RowParallelLinear, SPMDRank
    from neuronx_distributed_b200.parallel_layers.pad import pad_model
    from neuronx_distributed_b200.trainer import post_partition_hooks as pph
    from neuronx_distributed_b200.utils.tensor_utils import cumsum

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    # SPMDRank: the rank lives in a sharded weight; the pre-shard hook fabricates arange(world) so sharding hands out ids
    sr = SPMDRank(world)
    assert int(sr()) == rank and sr.rank.tensor_model_parallel and sr.rank.partition_dim == 0
    sd = {}
    sr.preshard_hook(sd, "spmd_rank.")
    assert sd["spmd_rank.rank"].tolist() == list(range(world))
    # post-partition hooks run once, in order, then clear
    seen = []
    pph.register_post_partition_hook(lambda m: seen.append(("a", type(m).__name__)))
    pph.register_post_partition_hook(lambda tag: seen.append(("b", tag)), args=("x",))
    pph.execute_all_hooks(nn.Linear(2, 2))
    pph.execute_all_hooks(nn.Linear(2, 2))
    assert seen == [("a", "Linear"), ("b", "x")]
    # pad_model: 3 heads on tp=2 → padded to 4; padded heads are zero so the attention block output is unchanged
    class Attn(nn.Module):
        def __init__(self):
            super().__init__()
            self.num_heads = 3
            self.q = ColumnParallelLinear(8, 3 * 4, bias=False, gather_output=False)
            self.o = RowParallelLinear(3 * 4, 8, bias=False, input_is_parallel=True)

    ps.destroy_model_parallel(); ps.initialize_model_parallel(tensor_model_parallel_size=1)
    torch.manual_seed(0)
    a = Attn()
    x = torch.randn(5, 8, generator=torch.Generator().manual_seed(1))
    want = a.o(a.q(x))
    pad_model(a, tp_degree=2, n_heads=3)
    assert a.num_heads == 4 and a.q.weight.shape == (16, 8) and a.o.weight.shape == (8, 16)
    torch.testing.assert_close(a.o(a.q(x)), want)
    # scope and hook of the reference: only instances of ``wrapped_classes`` (and what is below them) are padded, and the hook
    # sees every module in that scope together with padded_heads / heads
    class Two(nn.Module):
        def __init__(self):
            super().__init__()
            self.attn, self.other = Attn(), Attn()
            self.attn.split_size = 12

    two, seen = Two(), []

    def hook(mod, ratio):
        seen.append((type(mod).__name__, ratio))
        if hasattr(mod, "split_size"):
            mod.split_size = int(mod.split_size * ratio)

    class Marker(Attn):
        pass

    two.attn.__class__ = Marker
    pad_model(two, tp_degree=2, n_heads=3, wrapped_classes=[Marker], pad_hook_fn=hook)
    assert two.attn.q.weight.shape == (16, 8) and two.other.q.weight.shape == (12, 8) and two.attn.split_size == 16
    assert sorted(n for n, _ in seen) == ["ColumnParallelLinear", "Marker", "RowParallelLinear"] and all(abs(r - 4 / 3) < 1e-9 for _, r in seen)
    ps.destroy_model_parallel(); ps.initialize_model_parallel(tensor_model_parallel_size=world)
    # cumsum helper keeps dtype semantics
    t = torch.arange(10, dtype=torch.float32).view(2, 5)
    torch.testing.assert_close(cumsum(t, 1), t.cumsum(1))
    assert cumsum(torch.ones(4, dtype=torch.long)).tolist() == [1, 2, 3, 4]
    # LoRA over the fused GQA-QKV projection: zero-initialised B ⇒ identical outputs; only adapters train
    class Blk(nn.Module):
        def __init__(self):
            super().__init__()
            self.qkv_proj = GQAQKVColumnParallelLinear(16, [4 * 4, 2 * 4], bias=False, gather_output=False)

        def forward(self, x):
            return self.qkv_proj(x)

    torch.manual_seed(2)
    blk = Blk()
    xin = torch.randn(3, 2, 16, generator=torch.Generator().manual_seed(3))
    base = [t.clone() for t in blk(xin)]
    lm = get_lora_model(blk, LoraConfig(lora_rank=2, lora_alpha=4, target_modules=["qkv_proj"]))
    out = lm(xin)
    for a_, b_ in zip(out, base):
        torch.testing.assert_close(a_, b_)
    trainable = [n for n, p_ in lm.named_parameters() if p_.requires_grad]
    assert trainable and all("lora_" in n for n in trainable)
    sum(t.sum() for t in out).backward()
    assert any(p_.grad is not None for n, p_ in lm.named_parameters() if "lora_B" in n)
    # merge / unmerge into the fused [q_r; k_r; v_r] parameter: same outputs as the un-merged adapter, exact restore
    lora_layer = next(m for m in lm.modules() if type(m).__name__ == "LoraGQAQKVParallelLinear")
    with torch.no_grad():
        for b in (lora_layer.lora_B_q, lora_layer.lora_B_k, lora_layer.lora_B_v):
            b.weight.normal_(generator=torch.Generator().manual_seed(5 + rank))
    w0 = lora_layer.base_layer.weight_qkv.detach().clone()
    ref = [t.detach().clone() for t in lm(xin)]
    assert not all(torch.allclose(a_, b_) for a_, b_ in zip(ref, base))
    lora_layer.merge(safe_merge=True)
    assert lora_layer.merged and len(lora_layer.get_qkv(lora_layer.base_layer)) == 3
    for a_, b_ in zip(lm(xin), ref):
        torch.testing.assert_close(a_, b_, atol=1e-5, rtol=1e-5)
    lora_layer.unmerge()
    torch.testing.assert_close(lora_layer.base_layer.weight_qkv, w0, atol=1e-6, rtol=0)
    for a_, b_ in zip(lm(xin), ref):
        torch.testing.assert_close(a_, b_, atol=1e-5, rtol=1e-5)


def test_spmd_rank_hooks_pad_cumsum_lora_gqa(tmp_path):
    run_distributed(_spmd_pad_lora, 2, str(tmp_path), timeout=120)


def test_parallel_state_context_builds_any_rank_in_one_process():
    """NxDParallelState: construct rank 1's shard of a TP=2 layer in a single process (no process group of size 2)."""
    from neuronx_distributed_b200.trace.parallel_context import NxDParallelState
    from neuronx_distributed_b200.parallel_layers.layers import ColumnParallelLinear

    shards = {}
    for r in (0, 1):
        with NxDParallelState(world_size=2, rank=r, tensor_model_parallel_size=2):
            torch.manual_seed(0)
            lin = ColumnParallelLinear(6, 8, bias=False, gather_output=False, keep_master_weight=True)
            shards[r] = (lin.weight.detach().clone(), lin.master_weight.clone())
    full = shards[0][1]
    torch.testing.assert_close(torch.cat([shards[0][0], shards[1][0]]), full)


def _trace_v0(rank, world, tmp):
    from neuronx_distributed_b200.trace.trace import parallel_model_load, parallel_model_save, parallel_model_trace
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.parallel_layers.layers import ColumnParallelLinear, RowParallelLinear

    ps.initialize_model_parallel(tensor_model_parallel_size=world)

    class MLP(nn.Module):
        def __init__(self):
            super().__init__()
            self.up = ColumnParallelLinear(8, 16, bias=False, gather_output=False)
            self.down = RowParallelLinear(16, 8, bias=False, input_is_parallel=True)

        def forward(self, x):
            return self.down(torch.relu(self.up(x)))

    torch.manual_seed(0)
    holder = {}

    def build():
        holder["m"] = MLP().eval()
        return holder["m"], None

    x = torch.randn(4, 8, generator=torch.Generator().manual_seed(1))
    traced = parallel_model_trace(build, (x,), tp_degree=world)
    with torch.no_grad():
        torch.testing.assert_close(traced(x), holder["m"](x))
    d = os.path.join(tmp, f"rank{rank}")
    parallel_model_save(traced, d)
    meta = parallel_model_load(d)
    assert meta is not None and os.path.exists(os.path.join(d, "nxd_model_meta.pt"))


def test_v0_trace_save_load(tmp_path):
    run_distributed(_trace_v0, 2, str(tmp_path), timeout=120)


def _zero_dcp_and_ep(rank, world, tmp):
    """(a) ZeRO-1 optimizer state through the DCP writer survives a save → fresh optimizer → load → identical next step;
    (b) the EP-aware ZeRO-1 optimizer steps expert and non-expert parameters with their own sharding groups."""
    import torch.distributed as dist

    from neuronx_distributed_b200.modules.moe import ExpertMLPsV2, MoE, RoutedExpertsMLPOpsConfig, RouterTopK
    from neuronx_distributed_b200.optimizer import NeuronEPZero1Optimizer, NeuronZero1Optimizer
    from neuronx_distributed_b200.optimizer import zero_dcp_utils as dcp
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=1)          # dp = world

    def make():
        torch.manual_seed(0)
        m = nn.Sequential(nn.Linear(8, 16), nn.Tanh(), nn.Linear(16, 4))
        return m, NeuronZero1Optimizer(m.parameters(), torch.optim.AdamW, lr=1e-2, grad_clipping=True, max_norm=1.0)

    def step(m, o, seed):
        x = torch.randn(6, 8, generator=torch.Generator().manual_seed(seed + rank))
        o.zero_grad(); m(x).pow(2).mean().backward(); o.step()

    m1, o1 = make()
    step(m1, o1, 1)
    dcp.save_optim_state_dict(os.path.join(tmp, "optim"), o1.state_dict(), o1)
    dist.barrier()
    m2, o2 = make()
    m2.load_state_dict(m1.state_dict())
    for fg1, fg2 in zip(o1.flat_groups, o2.flat_groups):
        fg2.param_flat.copy_(fg1.param_flat)
    o2.load_state_dict(dcp.load_optim_state_dict(os.path.join(tmp, "optim"), o2))
    step(m1, o1, 2); step(m2, o2, 2)
    for p1, p2 in zip(m1.parameters(), m2.parameters()):
        torch.testing.assert_close(p1, p2, rtol=1e-6, atol=1e-7)
    # the reference's signatures: save(path, state_dict, aux_infos, dedup) / load(path, optimizer, aux_infos, dedup)
    aux = dcp.get_dcp_aux_infos(m1, o1)
    dcp.save_optim_state_dict(os.path.join(tmp, "optim_ref"), o1.state_dict(), aux, dedup=True)
    dist.barrier()
    m3, o3 = make()
    m3.load_state_dict(m1.state_dict())
    for fg1, fg3 in zip(o1.flat_groups, o3.flat_groups):
        fg3.param_flat.copy_(fg1.param_flat)
    o3.load_state_dict(dcp.load_optim_state_dict(os.path.join(tmp, "optim_ref"), o3, dcp.get_dcp_aux_infos(m3, o3), dedup=True))
    step(m1, o1, 3); step(m3, o3, 3)
    for p1, p3 in zip(m1.parameters(), m3.parameters()):
        torch.testing.assert_close(p1, p3, rtol=1e-6, atol=1e-7)
    # ---- EP-aware ZeRO-1
    ps.destroy_model_parallel()
    ps.initialize_model_parallel(tensor_model_parallel_size=1, expert_model_parallel_size=world)
    torch.manual_seed(0)
    E, k, H, I = 4, 2, 8, 16
    layer = MoE(RouterTopK(E, k, H), ExpertMLPsV2(RoutedExpertsMLPOpsConfig(normalize_top_k_affinities=True, num_experts=E, top_k=k, hidden_size=H,
                                                                              intermediate_size=I, capacity_factor=4.0)))
    opt = NeuronEPZero1Optimizer(layer.parameters(), torch.optim.AdamW, lr=1e-2, grad_clipping=True, max_norm=1.0)
    before = [p_.detach().clone() for p_ in layer.parameters()]
    x = torch.randn(10, 1, H, generator=torch.Generator().manual_seed(5 + rank))
    opt.zero_grad(); layer(x)[0].pow(2).mean().backward(); opt.step()
    moved = [not torch.equal(a, b.detach()) for a, b in zip(before, layer.parameters())]
    gn = opt.grad_norm() if callable(opt.grad_norm) else opt.grad_norm
    assert any(moved) and gn is not None and torch.isfinite(torch.as_tensor(gn))
    # non-expert parameters (router) stay identical across the EP/DP ranks after the step
    rw = layer.router.linear_router.weight.detach().clone()
    other = rw.clone(); dist.all_reduce(other)
    torch.testing.assert_close(other, rw * world)


def test_zero1_dcp_roundtrip_and_ep_zero1(tmp_path):
    run_distributed(_zero_dcp_and_ep, 2, str(tmp_path), timeout=180)


def _dcp_phase(rank, world, tmp, phase):
    """ZeRO-1 DCP checkpoint written at DP=2 and resumed at another DP degree: the re-sliced Adam state must give exactly
    the step the original run takes next (same global batch, split over more / fewer ranks)."""
    import torch.distributed as dist

    from neuronx_distributed_b200.optimizer import NeuronZero1Optimizer
    from neuronx_distributed_b200.optimizer import zero_dcp_utils as dcp
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=1)

    def make():
        torch.manual_seed(0)
        m = nn.Sequential(nn.Linear(8, 24), nn.Tanh(), nn.Linear(24, 5))
        return m, NeuronZero1Optimizer(m.parameters(), torch.optim.AdamW, lr=1e-2, grad_clipping=True, max_norm=1.0)

    def step(m, o, seed):
        x = torch.randn(8, 8, generator=torch.Generator().manual_seed(seed)).chunk(world)[rank]     # global batch of 8 rows
        o.zero_grad(); m(x).pow(2).mean().backward(); o.step()

    m, o = make()
    if phase == "save":
        step(m, o, 1); step(m, o, 2)
        dcp.save_optim_state_dict(os.path.join(tmp, "optim"), o.state_dict(), o)
        if rank == 0:
            torch.save(m.state_dict(), os.path.join(tmp, "model.pt"))
        step(m, o, 3)
        if rank == 0:
            torch.save(m.state_dict(), os.path.join(tmp, "after.pt"))
        dist.barrier()
        return
    m.load_state_dict(torch.load(os.path.join(tmp, "model.pt")))
    for fg in o.flat_groups:                                   # parameters live in the flat buffer; refresh the fp32 master shard
        lo, hi = fg.shard_range
        fg.master_shard.copy_(fg.param_flat[lo:hi])
    assert not os.path.isdir(os.path.join(tmp, "optim", f"zero1_rank_{rank:02d}_of_{world:02d}"))
    o.load_state_dict(dcp.load_optim_state_dict(os.path.join(tmp, "optim"), o))
    step(m, o, 3)
    want = torch.load(os.path.join(tmp, "after.pt"))
    for k, v in m.state_dict().items():
        torch.testing.assert_close(v, want[k], rtol=1e-5, atol=1e-6)


def test_zero1_dcp_resume_at_other_dp_degree(tmp_path):
    run_distributed(_dcp_phase, 2, str(tmp_path), "save", timeout=180)
    run_distributed(_dcp_phase, 4, str(tmp_path), "load", timeout=180)
    run_distributed(_dcp_phase, 1, str(tmp_path), "load", timeout=180)
