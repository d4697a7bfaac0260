"""Small public helpers that no other test names (found by a reference sweep of tests / examples against the package's public
definitions): each is exercised against its documented behaviour, so none of them is an unexercised shell."""
import os

import pytest
import torch
from torch import nn

from dist_utils import run_distributed


def _state(rank, world):
    """Process-group registry accessors and setters, message prefixes, the metadata store and the gloo pipeline groups."""
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.parallel_layers import random as prandom
    from neuronx_distributed_b200.parallel_layers import utils as pu

    ps.initialize_model_parallel(tensor_model_parallel_size=2, pipeline_model_parallel_size=2)
    tp_rank, pp_rank = rank % 2, rank // 2
    assert ps.get_tensor_model_parallel_ranks() == [pp_rank * 2, pp_rank * 2 + 1]
    assert ps.get_zero1_sharding_ranks() == [rank]                                     # dp = 1: the shard group is this rank
    msg = ps.rmsg("hello")
    assert msg.endswith("hello") and f"pp{pp_rank}" in msg and f"tp{tp_rank}" in msg
    assert ps.rmsg_ep("x").endswith("x") and "ep" in ps.rmsg_ep("x")
    # tensor-parallel duplicates: a replicated parameter counts once (on tp rank 0), a sharded one on every rank
    rep, shard = nn.Parameter(torch.ones(2)), nn.Parameter(torch.ones(2))
    pu.set_tensor_model_parallel_attributes(shard, True, 0, 1)
    assert pu.param_is_not_tensor_parallel_duplicate(shard) and pu.param_is_not_tensor_parallel_duplicate(rep) == (tp_rank == 0)
    # token-shuffle groups exist only once initialised; size 1 before
    with pytest.raises(AssertionError):
        ps.get_token_shuffle_group_size()                                              # not initialised yet (as in the reference)
    ps.initialize_token_shuffle_group(1)                                               # dp = 1 here: groups of one
    assert ps.get_token_shuffle_group_size() == 1 and [rank] in ps.get_token_shuffle_replica_groups()
    # speculative draft sub-groups of TP
    ps.initialize_speculative_draft_group(1)
    # python-object exchange needs the gloo mirror of the PP groups: idempotent
    ps.initialize_pp_gloo_groups()
    assert ps.get_pp_gloo_group() is not None
    assert ps.is_tcp_store_available() in (True, False)
    if ps.is_tcp_store_available():
        assert ps.get_tcp_store() is not None
    # manual overrides used by offline tools (checkpoint conversion runs one process that plays every rank)
    ps.set_expert_model_parallel_size(4); ps.set_expert_model_parallel_rank(3); ps.set_context_model_parallel_size(2)
    assert ps.get_expert_model_parallel_size() == 4 and ps.get_expert_model_parallel_rank() == 3
    assert ps.get_context_model_parallel_size() == 2
    ps.set_expert_model_parallel_size(None); ps.set_expert_model_parallel_rank(None); ps.set_context_model_parallel_size(None)
    assert ps.get_expert_model_parallel_size() == 1 and ps.get_context_model_parallel_size() == 1
    # RNG tracker class under its generic name
    assert prandom.RNGStatesTracker is type(prandom.get_rng_tracker()) or issubclass(type(prandom.get_rng_tracker()), prandom.RNGStatesTracker)
    tr = prandom.RNGStatesTracker()
    tr.add("stream", 123)
    with tr.fork("stream"):
        a = torch.rand(3)
    tr2 = prandom.RNGStatesTracker()
    tr2.add("stream", 123)
    with tr2.fork("stream"):
        b = torch.rand(3)
    assert torch.equal(a, b)
    with pytest.raises(Exception):
        tr.add("stream", 5)                                                            # a name can be registered once


def test_parallel_state_accessors_tp2_pp2():
    run_distributed(_state, 4, timeout=180)


def test_tensor_utils_and_casts():
    from neuronx_distributed_b200.parallel_layers import utils as pu

    t = torch.arange(24.0).view(2, 12)
    parts = pu.split_tensor_along_dim(t, 1, 3)
    assert [p.shape for p in parts] == [(2, 4)] * 3 and not parts[1].is_contiguous()
    assert all(p.is_contiguous() for p in pu.split_tensor_along_dim(t, 1, 3, contiguous_split_chunks=True))
    assert torch.equal(torch.cat(pu.split_tensor_along_second_dim(t, 4), 1), t)
    with pytest.raises(AssertionError):
        pu.ensure_divisibility(7, 2)
    pu.ensure_divisibility(8, 2)
    assert pu.cast_tensor(torch.ones(2)).dtype == torch.bfloat16 and pu.cast_tensor(torch.ones(2, dtype=torch.int32)).dtype == torch.int32
    assert pu.cast_tensor(torch.ones(2, dtype=torch.bfloat16), torch.bfloat16, torch.float32).dtype == torch.float32
    assert pu.is_torch_version_greater_than_2() and not pu.is_pjrt_device() and not pu.requires_init_pg_override()


def test_quantization_utils_roundtrip():
    from neuronx_distributed_b200.quantization import quantization_utils as qu
    from neuronx_distributed_b200.quantization.dequantize import mx_dequantize
    from neuronx_distributed_b200.quantization.microscaling.mx_torch import quantize_mx
    from neuronx_distributed_b200.quantization.quantization_config import get_float4x4_torch_dtype

    torch.manual_seed(0)
    w = torch.randn(6, 16)
    assert qu.qmax(torch.int8) == 127 and qu.qmax(torch.float8_e4m3fn) == 448
    q, s = qu.quantize_per_tensor(w, torch.int8)
    assert q.dtype == torch.int8 and s.numel() == 1 and (q.float() * s - w).abs().max() <= s.item() * 0.51
    q, s = qu.quantize_per_channel(w, torch.int8, axis=0)
    assert s.shape == (6, 1) and ((q.float() * s - w).abs() <= s * 0.51).all()
    # scales of torch's own quantised tensors
    qt = torch.quantize_per_tensor(w, 0.05, 0, torch.qint8)
    assert float(qu.extract_q_scale_per_tensor(qt)) == pytest.approx(0.05)
    qc = torch.quantize_per_channel(w, torch.full((6,), 0.1), torch.zeros(6, dtype=torch.long), 0, torch.qint8)
    assert qu.extract_q_scale_per_channel(qc).shape == (6, 1) and torch.allclose(qu.extract_q_scale_per_channel(qc), torch.full((6, 1), 0.1, dtype=torch.float64).to(qu.extract_q_scale_per_channel(qc).dtype))
    # blockwise: one scale per 2 x 8 block
    sc = torch.rand(3, 2) + 0.5
    qb = torch.randint(-5, 5, (6, 16), dtype=torch.int8)
    deq = qu.dequantize_blockwise(qb, sc, [0, 1], [2, 8], torch.float32)
    assert torch.allclose(deq[2:4, 8:], qb[2:4, 8:].float() * sc[1, 1])
    # MX de-quantisation entry point of the quantised layers agrees with the packer
    for kind in ("mxfp4", "mxfp8"):
        x = torch.randn(4, 64)
        packed, scale = quantize_mx(x, kind)
        back = mx_dequantize(packed, scale, kind, torch.float32)
        assert back.shape == x.shape and ((back - x).norm() / x.norm()) < (0.2 if kind == "mxfp4" else 0.05)
    assert get_float4x4_torch_dtype() in (torch.uint16, torch.float16)


def test_model_utils_helpers(tmp_path):
    from neuronx_distributed_b200.utils import cpu_mode, set_cpu_mode
    from neuronx_distributed_b200.utils import model_utils as mu
    from neuronx_distributed_b200.utils.medusa_utils import pad_path
    from neuronx_distributed_b200.utils.random import set_random_seed
    from neuronx_distributed_b200.utils.safetensors_utils import remove_shared_tensors
    from neuronx_distributed_b200.utils.sampling import create_sampler

    emb, head = nn.Embedding(8, 4), nn.Linear(4, 8, bias=False)
    head.weight = emb.weight
    model = nn.Sequential(emb, head)
    tied = mu.get_tied_parameters(model)
    assert any(set(g) == {"0.weight", "1.weight"} for g in (tied.values() if isinstance(tied, dict) else tied)), tied
    # parallel attributes survive a materialisation / re-initialisation round trip
    emb.weight.tensor_model_parallel, emb.weight.partition_dim = True, 0
    saved = mu.preserve_parallel_attributes(model)
    del emb.weight.tensor_model_parallel
    mu.restore_parallel_attributes(model, saved)
    assert emb.weight.tensor_model_parallel is True and emb.weight.partition_dim == 0
    with mu.init_on_device(torch.device("meta")):
        meta = nn.Linear(3, 3)
    assert meta.weight.device.type == "meta"
    mu.maybe_materalize_model(meta)
    assert meta.weight.device.type == "cpu" and meta.weight.shape == (3, 3)
    assert not mu.is_nxd_pipeline_model(model) and mu.is_nxdt_available() in (True, False) and not mu.is_nxdt_pretrained_model(model)
    assert mu.is_hf_accelerate_available() in (True, False)
    assert int(mu.get_platform_lnc()) == 1 and mu.LogicalNCConfig(2).name == "LNC_2"
    # seeds: python / numpy / torch move together
    import random as pyrandom

    set_random_seed(11); a = (pyrandom.random(), float(torch.rand(())))
    set_random_seed(11); b = (pyrandom.random(), float(torch.rand(())))
    assert a == b
    # safetensors cannot store aliases: one copy is kept, the alias map says who shared it
    sd = {"a": emb.weight.data, "b": emb.weight.data, "c": torch.zeros(2)}
    kept, aliases = remove_shared_tensors(sd)
    assert set(kept) | set(aliases) == {"a", "b", "c"} and len(kept) == 2 and list(aliases.values())[0] in kept
    assert pad_path([1, 2], 4) == [1, 2, -2, -2] and pad_path([1, 2, 3], 3) == [1, 2, 3]
    s = create_sampler(top_k=3, do_sample=True)
    assert s.top_k == 3 and s.do_sample
    from types import SimpleNamespace as NS

    s2 = create_sampler(NS(top_k=5, temperature=0.7, do_sample=True))
    assert (s2.top_k, s2.temperature) == (5, 0.7)
    prev = cpu_mode()
    set_cpu_mode(True)
    assert cpu_mode() is True
    set_cpu_mode(None if prev is not True else True)


def test_hlo_utils_extras(tmp_path):
    """Remaining program-surgery helpers: weight indices stored on a plan, per-weight layout transform lookup, metadata files."""
    import json

    from neuronx_distributed_b200.trace import hlo_utils as hu
    from neuronx_distributed_b200.trace.functions import trace
    from neuronx_distributed_b200.trace.nxd_model.utils import ts_convert_dict_to_ordered_list_type_list_tensor

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.w = nn.Parameter(torch.randn(8, 8, dtype=torch.bfloat16))
            self.v = nn.Parameter(torch.randn(8))

        def forward(self, x):
            return x @ self.w.float().t() + self.v

    m = M()
    ta = trace(m, (torch.randn(2, 8),))
    plan = ta.record_plan()
    idx = ta.weight_name_to_idx
    hu.add_weight_idx_attr_to_hlo(plan, idx, weight_names_to_skip={"v"})
    assert plan.meta[hu.TRANSPOSABLE_WEIGHT_IDX] == [idx["w"]]
    transformer, main = hu.extract_weight_layout_transform_hlo(ta)
    f = hu.get_wlt(transformer, "w")
    assert f is not None and any(t.dtype == torch.float32 for t in f(m.w.detach()))          # the hoisted bf16 → fp32 cast
    assert hu.get_wlt(transformer, "v") is None                                             # nothing hoisted for v
    meta = hu.prepare_metaneff_for_wlt_hlo(transformer)
    path = str(tmp_path / "meta.json")
    json.dump(meta, open(path, "w"), default=str)
    assert hu.read_metaneff(path)["outputs"] == json.loads(json.dumps(meta, default=str))["outputs"]
    assert os.path.isdir(hu.get_compiler_package_dir())
    with pytest.raises(NotImplementedError, match="no computations to renumber"):           # one value-id space per plan: says so
        hu.update_computation_id_and_name(plan)
    params = [("a", False), ("b", True)]
    vals, names = ts_convert_dict_to_ordered_list_type_list_tensor(params, {"b": [torch.zeros(1)], "a": [torch.ones(1)]}, 0)
    assert names == ["a", "b"] and float(vals[0][0]) == 1.0                                 # signature order, not dict order


def _opt_from_class(rank, world):
    import neuronx_distributed_b200 as nxd
    from neuronx_distributed_b200.trainer.trainer import initialize_optimizer_from_class

    cfg = nxd.neuronx_distributed_config(optimizer_config={"zero_one_enabled": True, "grad_clipping": True, "max_grad_norm": 1.0})
    model = nxd.initialize_parallel_model(cfg, lambda: nn.Linear(4, 4))
    opt = initialize_optimizer_from_class(cfg, torch.optim.AdamW, model.parameters(), model=model, lr=1e-2)
    x = torch.randn(3, 4)
    before = [p.detach().clone() for p in model.parameters()]
    model(x).pow(2).mean().backward()
    opt.step()
    assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters())) and opt.grad_norm is not None


def test_initialize_optimizer_from_class():
    run_distributed(_opt_from_class, 1, timeout=120)


def test_mock_distributed_schedule_bases_and_state_dict_adaptor():
    from neuronx_distributed_b200.pipeline import scheduler as S
    from neuronx_distributed_b200.quantization.quantization_config import MyEnumMeta, QuantizationType
    from neuronx_distributed_b200.quantization.quantization_layers import QuantizedParallelLinearLayerStateDictAdaptor as Ad
    from neuronx_distributed_b200.trace.mock_torchdist import MockDistributed
    from neuronx_distributed_b200.trace.trace import ParallelModel
    from neuronx_distributed_b200.utils.serialization import TensorStub
    from neuronx_distributed_b200.utils.tensor_replacement import TensorReplacementRegistry, enable_tensor_replacement

    # stand-in for torch.distributed while sharding checkpoints for N ranks in one process
    md = MockDistributed(world_size=8)
    md.init_process_group("nccl", rank=3, world_size=8)
    assert md.is_available() and md.is_initialized() and md.get_world_size() == 8 and md.get_rank() == 3
    g = md.new_group(ranks=[2, 3])
    assert md.get_world_size(g) == 2 and md.get_process_group_ranks(g) == [2, 3] and md.get_rank(g) == 1
    md.barrier(); md.destroy_process_group()
    assert not md.is_initialized()
    # schedule base classes: the deprecated lock-step name computes in 1F1B order; task records carry (mb, chunk)
    assert issubclass(S.Train1F1BSchedule, S.PipeSchedule) and issubclass(S.TrainSchedule, S.Train1F1BSchedule)
    assert S.TrainSchedule(4, 2, 0).compute_order() == S.Train1F1BSchedule(4, 2, 0).compute_order()
    t = S.ForwardStepTask(1, 0)
    assert isinstance(t, S.PipelineTask) and (t.mb, t.model_chunk) == (1, 0) and isinstance(S.ReduceGradsTask(), S.PostProcessTask)
    with pytest.raises(TypeError):
        S.PipeSchedule(4, 2, 0)                                                     # abstract
    # wrapper base of the traced models is a plain module container
    pm = ParallelModel()
    assert isinstance(pm, nn.Module) and list(pm.parameters()) == []
    # state-dict adaptor: plain layout and torch's packed qint8 layout answer the same questions
    w, b = torch.randint(-8, 8, (4, 3), dtype=torch.int8), torch.ones(4)
    sd = {"l.weight": w, "l.scale": torch.tensor([0.5]), "l.bias": b}
    assert torch.equal(Ad.get_weight_from_state_dict("l.", sd), w) and float(Ad.get_scale_from_state_dict("l.", sd)) == 0.5
    assert torch.equal(Ad.get_bias_from_state_dict("l.", sd), b)
    Ad.set_weight_to_state_dict("l.", w + 1, sd); Ad.set_bias_to_state_dict("l.", b * 2, sd)
    assert torch.equal(Ad.get_weight_from_state_dict("l.", sd), w + 1) and torch.equal(Ad.get_bias_from_state_dict("l.", sd), b * 2)
    qlin = torch.ao.nn.quantized.Linear(3, 4)
    qsd = {"q." + k: v for k, v in qlin.state_dict().items()}
    assert Ad.get_weight_from_state_dict("q.", qsd).shape == (4, 3) and Ad.get_scale_from_state_dict("q.", qsd).numel() >= 1
    # enum membership by raw value
    assert isinstance(QuantizationType, MyEnumMeta) and "per_tensor_symmetric" in QuantizationType and "nope" not in QuantizationType
    # replacement registry is a process-wide singleton that the hooks read
    lin = nn.Linear(2, 2)
    m = nn.Sequential(lin)
    reg = TensorReplacementRegistry.get()
    assert reg is TensorReplacementRegistry.get()
    enable_tensor_replacement(m, {"0": torch.full((1, 2), 7.0)})
    assert torch.equal(m(torch.zeros(1, 2)), torch.full((1, 2), 7.0)) and "0" in reg.replacements
    for h in reg.handles:
        h.remove()
    reg.handles.clear(); reg.replacements.clear(); reg.masks.clear()
    assert repr(TensorStub(3)) == "TensorStub(3)"


def _moe_helpers(rank, world):
    """Building blocks of ``ExpertMLPsV2`` that the reference exposes and callers reach directly."""
    from neuronx_distributed_b200.modules.moe import ExpertMLPsV2, RoutedExpertsMLPOpsConfig
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.parallel_layers.layers import SPMDRank

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    torch.manual_seed(0)
    E, k, H, I, T = 4, 2, 8, 16, 6
    em = ExpertMLPsV2(RoutedExpertsMLPOpsConfig(normalize_top_k_affinities=True, num_experts=E, top_k=k, hidden_size=H, intermediate_size=I))
    aff = torch.rand(T, E).softmax(-1)
    idx = aff.topk(k, -1).indices
    x = torch.randn(T, H)
    # all-experts set-up: (count, k-hot mask, affinities of the chosen experts only, tokens), optionally a subset of experts
    n, mask, a, xs = em.setup_all_experts(x, aff, idx)
    assert n == E and mask.shape == (T, E) and mask.sum(-1).tolist() == [k] * T and xs is x
    assert torch.allclose(a.sum(-1), torch.ones(T)) and ((a > 0) == (mask > 0)).all()
    n2, mask2, a2, _ = em.setup_all_experts(x, aff, idx, chosen_expert_indices=[0, 2])
    assert n2 == 2 and torch.equal(mask2, mask[:, [0, 2]]) and torch.equal(a2, a[:, [0, 2]])
    # masked affinities: computed from the router outputs unless the caller already has the full tensor
    assert torch.equal(em.maybe_get_expert_affinities_masked(idx, aff), a)
    marker = torch.zeros(1)
    assert em.maybe_get_expert_affinities_masked(idx, aff, expert_affinities_masked_full=marker) is marker
    # sequence-parallel router outputs are gathered over the TP group (rank r holds its own token shard)
    full = em.get_full_expert_affinities_masked(aff, idx)
    assert full.shape == (T * world, E) and torch.equal(full[rank * T:(rank + 1) * T], a)
    ga, gm, gi = em.get_sp_expert_masks_index(a, idx)
    assert ga.shape == (T * world, E) and gi.shape == (T * world, k) and torch.equal((ga > 0).to(torch.float64), gm)
    # block bookkeeping: which blocks hold at least one real token; the kernel-named mapping equals the default one
    from neuronx_distributed_b200.modules.moe.blockwise import build_block_metadata

    b2e, tp2id, _ = build_block_metadata(idx, E, 4)
    cond = em.get_block_conditions(4, b2e.numel(), tp2id)
    assert cond.dtype == torch.int32 and cond.tolist() == [int((tp2id.view(-1, 4)[b] != -1).any()) for b in range(b2e.numel())]
    m1 = em.get_blockwise_expert_and_token_mapping(T, b2e.numel(), None, idx, block_size=4)
    m2 = em.get_blockwise_expert_and_token_mapping_kernel(T, b2e.numel(), None, idx, block_size=4)
    assert all(torch.equal(u, v) for u, v in zip(m1, m2))
    # redundant experts: this rank hosts logical experts [1, 1, 3]; expert 1's token range is split between its two replicas,
    # and the second local copy of the same expert is switched off
    deg = torch.tensor([[0, 2, 0, 1]])                                            # one EP rank, replicas per expert
    start, end = em.allocate_token_blocks(deg, [8, 8, 8, 8])
    lm = torch.ones(8, 3)
    out = em.generate_local_expert_mask_with_redundancy(lm, torch.tensor([1, 1, 3]), start, end, E, 0)
    assert out[:, 0].sum() == 8 and out[:, 1].sum() == 0 and out[:, 2].sum() == 8
    # expert weights can be rebuilt for other groups (hybrid prefill / decode sharding)
    old = em.mlp_op
    new = em.initialize_mlp_op(ps.get_tensor_model_parallel_group(), None, is_prefill=False)
    assert new is em.mlp_op and new is not old and new.down_proj.weight.shape == old.down_proj.weight.shape
    assert em.get_spmd_rank() is None
    # the SPMD rank module carries this rank's expert ids as a (sharded) weight
    sr = SPMDRank(world_size=world)
    p = sr.initialize_expert_indices(E)
    assert p.shape == (1, E) and sr.get_local_expert_indices() is p and p.tolist() == [[0, 1, 2, 3]] and p.tensor_model_parallel


def test_expert_mlp_building_blocks_tp2():
    run_distributed(_moe_helpers, 2, timeout=180)


def test_schedule_ids_rng_states_observer_and_dtypes():
    from neuronx_distributed_b200.parallel_layers import random as prandom
    from neuronx_distributed_b200.pipeline import scheduler as S
    from neuronx_distributed_b200.quantization.observer import PerChannelAbsMaxObserver
    from neuronx_distributed_b200.quantization.quantization_config import QuantizedDtype

    sch = S.TrainInterleavedSchedule(4, 2, 2, 0)                                    # 4 micro-batches, 2 chunks, pp = 2, stage 0
    order = sch.compute_order()
    fwd = [(mb, c) for is_f, mb, c in order if is_f]
    assert [(sch.get_microbatch_id(i), sch.get_model_chunk_id(i)) for i in range(len(fwd))] == fwd
    assert sorted(fwd) == sorted((m, c) for m in range(4) for c in range(2))
    # RNG tracker state can be exported / restored (activation recompute replays the same dropout masks)
    tr = prandom.RNGStatesTracker()
    tr.add("mp", 7)
    snap = {k: v.clone() for k, v in tr.get_states().items()}
    with tr.fork("mp"):
        a = torch.rand(4)
    tr.set_states(snap)
    with tr.fork("mp"):
        b = torch.rand(4)
    assert torch.equal(a, b)
    # per-channel abs-max observer accumulates over calls and can be reset
    ob = PerChannelAbsMaxObserver(ch_axis=0)
    ob(torch.tensor([[1.0, -3.0], [0.5, 0.2]])); ob(torch.tensor([[2.0, 1.0], [-4.0, 0.1]]))
    assert ob.abs_max.flatten().tolist() == [3.0, 4.0]
    ob.reset_min_max_vals()
    ob(torch.tensor([[1.0, 0.0], [0.0, 0.5]]))
    assert ob.abs_max.flatten().tolist() == [1.0, 0.5]
    scale, zero = ob.calculate_qparams()
    assert scale.shape == (2,) and torch.allclose(scale, torch.tensor([1.0, 0.5]) / ob.quant_max) and not zero.any()
    assert QuantizedDtype.INT8.storage_dtype() is torch.int8 and QuantizedDtype.F8E4M3FN_X4.storage_dtype() is torch.uint32


def _builder_pieces(rank, world, tmp):
    """The separately callable steps of the first-generation builder and the runtime model's initialisation variants."""
    import neuronx_distributed_b200  # noqa: F401
    from neuronx_distributed_b200.parallel_layers import ColumnParallelLinear, RowParallelLinear
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.trace.model_builder import ModelBuilder

    ps.initialize_model_parallel(tensor_model_parallel_size=world)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.up = ColumnParallelLinear(8, 16, bias=False, gather_output=False)
            self.down = RowParallelLinear(16, 8, bias=False, input_is_parallel=True)
            self.register_buffer("cache", torch.zeros(2, 8))

        def forward(self, x):
            return self.down(torch.relu(self.up(x)))

    torch.manual_seed(0)
    net = Net()
    torch.manual_seed(1)
    full = {"up.weight": torch.randn(16, 8), "down.weight": torch.randn(8, 16)}
    mb = ModelBuilder(None, world, lambda: dict(full), model=net, use_cuda_graphs=False)       # reference positional order
    mb.add("main", net, [(torch.randn(2, 8),)])
    shard = mb.shard_weights(rank)
    assert shard["up.weight"].shape == (16 // world, 8) and torch.equal(shard["up.weight"], full["up.weight"].chunk(world, 0)[rank])
    assert torch.equal(mb.shard_weights_with_cache(rank)["down.weight"], full["down.weight"].chunk(world, 1)[rank])
    assert torch.equal(mb.shard_weights_with_cache(rank)["up.weight"], shard["up.weight"])      # second call: cached full checkpoint
    cast = ModelBuilder.cast_weights({"up.weight": torch.ones(16 // world, 8, dtype=torch.float64)}, net)
    assert cast["up.weight"].dtype == torch.float32
    init = mb.build_state_initializer()
    states = init()
    assert list(states[0]) == ["cache"] and states[0]["cache"].shape == (2, 8) and not states[0]["cache"].any()
    fl, pk = mb.build_flattener_map(), mb.build_packer()
    assert list(fl) == ["main_0"] and fl["main_0"]((1, 2)) == [1, 2] and pk("y") == "y"
    nxd_model = mb.build_nxd_model()
    # weights: `initialize` = set_weights + to_neuron in one call; `initialize_with_saved_weights` trusts the modules
    nxd_model.initialize_spmd_models(None, [mb.shard_weights(r) for r in range(world)], 0)
    x = torch.randn(2, 8, generator=torch.Generator().manual_seed(3))
    want = torch.relu(x @ full["up.weight"].t()) @ full["down.weight"].t()
    torch.testing.assert_close(nxd_model(x), want, rtol=1e-4, atol=1e-5)
    nxd2 = mb.build_nxd_model()
    nxd2.mock_initialization(True)
    assert nxd2.loaded_on_device
    nxd2.initialize_with_saved_weights(0)
    torch.testing.assert_close(nxd2(x), want, rtol=1e-4, atol=1e-5)
    mb.write_neff_to_file(nxd_model, os.path.join(tmp, f"r{rank}"))
    assert "main" in open(os.path.join(tmp, f"r{rank}", "programs.txt")).read()


def test_v1_builder_steps_and_runtime_initialisation_tp2(tmp_path):
    run_distributed(_builder_pieces, 2, str(tmp_path), timeout=180)


def test_lora_serving_config_files_and_lightning_bits(tmp_path):
    from neuronx_distributed_b200.lightning.accelerator import NeuronXLAAccelerator
    from neuronx_distributed_b200.lightning.logger import NeuronTensorBoardLogger
    from neuronx_distributed_b200.modules.lora.serving import LoraServingConfig

    cfg = LoraServingConfig(max_loras=3, max_lora_rank=8, target_modules=["q_proj"], lora_dtype=torch.bfloat16)
    f = str(tmp_path / "serving.json")
    cfg.to_json_file(f)
    back = LoraServingConfig.from_json_file(f)
    assert (back.max_loras, back.max_lora_rank, back.target_modules, back.lora_dtype) == (3, 8, ["q_proj"], torch.bfloat16)
    assert LoraServingConfig.from_json_file(str(tmp_path / "missing.json")) is None
    assert LoraServingConfig.from_json_file(f, max_loras=5).max_loras == 5                     # keyword overrides win

    class Registry(dict):
        def register(self, name, cls, description=""):
            self[name] = (cls, description)

    reg = Registry()
    NeuronXLAAccelerator.register_accelerators(reg)
    assert reg["b200"][0] is NeuronXLAAccelerator
    assert NeuronXLAAccelerator().get_device_stats(torch.device("cpu")) == {}                 # no CUDA here: nothing to report
    lg = NeuronTensorBoardLogger(str(tmp_path), name="run", version="v0")
    lg.log_metrics({"loss": 1.5}, step=1)
    assert lg.experiment is not None and lg.print_step() in (True, False)
    lg.log_graph(nn.Linear(2, 2))                                                             # accepted, nothing to draw


def _last_batch(rank, world, tmp):
    import neuronx_distributed_b200 as nxd
    from neuronx_distributed_b200.inference.gqa import GQA, BaseGroupQueryAttention
    from neuronx_distributed_b200.lightning import NeuronCheckpointIO, NeuronLTModule
    from neuronx_distributed_b200.models.llama import LlamaConfig, LlamaForCausalLM
    from neuronx_distributed_b200.modules.lora import LoraConfig, LoraModel
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.pipeline import NxDPPModel
    from neuronx_distributed_b200.scripts.checkpoint_converter import CheckpointConverterBase
    from neuronx_distributed_b200.utils.serialization import SerializationManager, TensorStub

    ps.initialize_model_parallel(tensor_model_parallel_size=1, pipeline_model_parallel_size=world)
    # head layout of a GQA block under a TP degree that exceeds the KV heads: KV heads are replicated up to the degree
    g = BaseGroupQueryAttention(64, 8, num_attention_heads=8, num_key_value_heads=2, tp_degree=1,
                                desired_sharding_strategy=GQA.REPLICATE_TO_TP_DEGREE)
    assert g.get_sharding_strategy() in tuple(GQA) and g.get_num_attention_heads() == 8 and g.get_num_key_value_heads() >= 2
    # analytic parameter count of the flagship model == what it actually holds when nothing is sharded
    cfg = LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, dtype=torch.float32, max_position_embeddings=16)
    if world == 1:
        lm = LlamaForCausalLM(cfg)
        assert lm.num_parameters_global() == sum(p.numel() for p in lm.parameters())
    # pipeline model: names of the layers it cuts at, children of the local partition
    from test_pipeline_cpu import Block, Toy

    torch.manual_seed(0)
    ppm = NxDPPModel(Toy(tie=False), transformer_layer_cls=Block, num_microbatches=2, output_loss_value_spec=True,
                     input_names=["input_ids", "labels"])
    assert ppm.get_model_layers() == [f"layers.{i}" for i in range(4)] and len(list(ppm.local_children())) == 1
    # LoRA: the base weights under their original names (what a plain checkpoint of the base model contains)
    net = nn.Sequential(nn.Linear(4, 4), nn.ReLU(), nn.Linear(4, 2))
    lora = LoraModel(net, LoraConfig(lora_rank=2, target_modules=["0"]))
    base = lora.base_state_dict()
    assert set(base) == {"0.weight", "0.bias", "2.weight", "2.bias"} and torch.equal(base["0.weight"], lora.module[0].base_layer.weight)
    # Lightning module: decayed / un-decayed optimizer groups
    ncfg = nxd.neuronx_distributed_config(pipeline_parallel_size=world) if world == 1 else None
    if ncfg is not None:
        mod = NeuronLTModule(ncfg, lambda: nn.Sequential(nn.Linear(4, 4), nn.LayerNorm(4)), torch.optim.AdamW)
        mod.setup()
        groups = mod.get_param_groups_by_weight_decay(0.1, no_decay=("bias", "1.weight"))
        assert [g["weight_decay"] for g in groups] == [0.1, 0.0] and len(groups[0]["params"]) == 1 and len(groups[1]["params"]) == 3
    # checkpoint plugin removes what it wrote
    io = NeuronCheckpointIO(save_load_xser=False)
    lin = nn.Linear(2, 2)
    io.save_checkpoint({"state_dict": lin, "global_step": 1}, f"{tmp}/ck/step_1")
    assert os.path.isdir(f"{tmp}/ck/step_1")
    import torch.distributed as dist

    dist.barrier()
    if rank == 0:
        io.remove_checkpoint(f"{tmp}/ck/step_1")
        assert not os.path.exists(f"{tmp}/ck/step_1")
    # converter: which tensors of a Megatron-style checkpoint are query / output projections
    conv = CheckpointConverterBase()
    from types import SimpleNamespace as NS

    assert conv.is_q_or_o_for_megatron(NS(model_style="megatron"), "layers.0.self_attn.o_proj.weight")
    assert not conv.is_q_or_o_for_megatron(NS(model_style="hf"), "layers.0.self_attn.o_proj.weight")
    assert not conv.is_q_or_o_for_megatron(NS(model_style="megatron"), "layers.0.mlp.down_proj.weight")
    # serializer: placeholders of a skeleton in tensor order; very deep objects fail with the object's class name
    sm = SerializationManager()
    skel, tens = sm.serialize((torch.ones(1), {"k": torch.zeros(2)}, 3))
    stubs = sm.extract_stubs(skel)
    assert [type(s) for s in stubs] == [TensorStub, TensorStub] and [s.index for s in stubs] == [0, 1] and len(tens) == 2

    class Deep:
        pass

    with pytest.raises(RuntimeError, match="Deep"):
        with sm.catch_and_raise_for_large_object(Deep()):
            raise RecursionError()


def test_remaining_public_methods_single_rank(tmp_path):
    run_distributed(_last_batch, 1, str(tmp_path), timeout=180)


def test_remaining_public_methods_pp2(tmp_path):
    run_distributed(_last_batch, 2, str(tmp_path), timeout=180)


def test_hooks_callback_progress_bar_and_ring_alias(tmp_path):
    from types import SimpleNamespace as NS

    from neuronx_distributed_b200.kernels import nki_ring_attn_func
    from neuronx_distributed_b200.lightning import NeuronHooksCallback
    from neuronx_distributed_b200.lightning.progress_bar import NeuronTQDMProgressBar

    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(4, 8), nn.Tanh(), nn.Linear(8, 2))
    pl = NS(model=model, global_step=0)
    x = torch.randn(3, 4)
    # keyword style: dump outputs / output gradients of the matching leaf modules at the listed steps
    cb = NeuronHooksCallback(str(tmp_path / "dump"), steps=[0], module_filter="2", dump_grads=True)
    cb.on_train_start(None, pl)
    model(x).sum().backward()
    cb.on_train_batch_end(None, pl)
    model(x).sum().backward()                                           # step 1: not listed, nothing written
    cb.detach()
    files = sorted(os.listdir(tmp_path / "dump"))
    assert files == ["step0_2_bwd.pt", "step0_2_fwd.pt"], files
    assert torch.equal(torch.load(tmp_path / "dump" / "step0_2_fwd.pt"), model(x).detach())
    # the reference's config object: (input, output) of the target layers every `hooks_interval` steps, norms only
    cfg = NS(hooks=True, hooks_dump_base_directory=str(tmp_path / "ref"), target_layers="0, 2", hooks_interval=1,
             enable_activation_dumps=True, enable_grad_dumps=True, dump_only_norms=True, dump_only_master_rank=True,
             master_print_model_layers=False)
    cb2 = NeuronHooksCallback(cfg)
    cb2.on_train_start(None, pl)
    model(x).sum().backward()
    assert set(cb2.activations_map) == {"0", "2"} and set(cb2.gradients_map) == {"0", "2"}
    cb2.on_train_batch_end(None, pl)
    cb2.detach()
    d = tmp_path / "ref" / "0" / "global_step_0"
    names = sorted(os.listdir(d))
    assert [n.split("_rank")[0] for n in names] == ["grad_output", "input", "output"] or len(names) >= 3, names
    got = torch.load(d / [n for n in names if n.startswith("input")][0])
    assert got.dim() == 0 and torch.allclose(got, x.norm())             # dump_only_norms
    assert not cb2.activations_map and not cb2.gradients_map            # flushed
    # progress bar: created at train start on the printing rank, advances per batch
    bar = NeuronTQDMProgressBar()
    bar.setup()
    bar.on_train_start(None, pl)
    bar.on_train_batch_end(None, pl)
    assert bar._bar is None or bar._bar.n == 1
    if bar._bar is not None:
        bar._bar.close()
    # ring-attention entry point under the reference's kernel name: [B, H, S_local, D] in and out (one rank = plain causal attention)
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29673")
        dist.init_process_group("gloo", rank=0, world_size=1)
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    if not ps.model_parallel_is_initialized():
        ps.initialize_model_parallel(tensor_model_parallel_size=1)
    q, k, v = (torch.randn(1, 2, 8, 4) for _ in range(3))
    out = nki_ring_attn_func(q, k, v, causal=True)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    ps.destroy_model_parallel(); dist.destroy_process_group()
