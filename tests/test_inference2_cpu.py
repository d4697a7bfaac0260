"""Inference builder v2 (``trace`` → ``compile`` → ``NxDModel``), runtime-model API, ``neuronx_distributed.trace`` import paths,
in-place checkpoint sharding helpers (TP + EP + quantised hooks), mock distributed."""
import pytest
import torch
from torch import nn

from dist_utils import run_distributed


class _Toy(nn.Module):
    """Module-level so the runtime model can embed it in ``save`` (pickle needs an importable class)."""

    def __init__(self):
        super().__init__()
        self.lin = nn.Linear(4, 3)
        self.register_buffer("cache", torch.zeros(2, 3))

    def forward(self, x, scale=None):
        y = self.lin(x)
        if scale is not None:
            y = y * scale
        self.cache[: min(2, x.shape[0])].copy_(y[:2])
        return y, y.sum(-1)


def test_functional_units_and_runtime_model(tmp_path):
    from neuronx_distributed_b200.trace.functions import append_default_compiler_flags, compile, compile_wlo, trace
    from neuronx_distributed_b200.trace.model_builder_utils import (CompilationArtifacts, ModelBuilderConstants, TraceArtifacts,
                                                                    WLOArtifacts, generate_key)
    from neuronx_distributed_b200.trace.model_builder_v2 import ModelBuilder
    from neuronx_distributed_b200.trace.nxd_model import NxDModel, StateInitializer
    from neuronx_distributed_b200.trace.nxd_model.utils import (generate_route_key_from_provided_args, get_dtype_enum,
                                                                get_dtype_from_enum, retrieve_artifact_from_model,
                                                                ts_convert_dict_to_ordered_list_type_tensor)

    torch.manual_seed(0)
    m = _Toy().eval()
    # --- trace(): validation rules --------------------------------------------------------------------------------------
    ta = trace(m, args=torch.randn(2, 4))
    assert isinstance(ta, TraceArtifacts) and [a.param_name for a in ta.provided_args] == ["x"]
    assert [(p.param_name, p.is_positional) for p in ta.model_params] == [("x", True), ("scale", False)]
    assert ta.state_names == ["cache"] and set(ta.weight_name_to_idx) == {"lin.weight", "lin.bias"}
    assert ta.output_spec == (((2, 3), "torch.float32"), ((2,), "torch.float32"))
    assert generate_key(ta, "k") == "k" and generate_key(ta).startswith(ModelBuilderConstants.DEFAULT_KEY_PREFIX + "_")
    assert generate_key(ta) == generate_key(trace(m, kwargs={"x": torch.zeros(2, 4)}))          # same route → same key
    for bad, exc in (((m, ), ValueError), ((None, torch.ones(1)), ValueError), ((m, [torch.ones(2, 4)]), ValueError),
                     ((m, (torch.ones(2, 4), torch.ones(1), torch.ones(1))), ValueError)):
        with pytest.raises(exc):
            trace(*bad)
    with pytest.raises(ValueError):
        trace(m, torch.ones(2, 4), {"x": torch.ones(2, 4)})                                     # given twice
    with pytest.raises(ValueError):
        trace(m, None, {"nope": torch.ones(1)})
    with pytest.raises(ValueError):
        trace(m, None, {"scale": torch.ones(1)})                                                # required x missing
    with pytest.raises(NotImplementedError):
        trace(lambda *a: a[0], torch.ones(1))
    with pytest.raises(ValueError):
        trace(lambda x, k=3: x, torch.ones(1))                                                  # non-None default
    with pytest.raises(NotImplementedError):
        trace(m, torch.ones(2, 4), spmd=False)
    with pytest.raises(RuntimeError):
        trace(m, torch.ones(2, 5))                                                              # the eager tracing run fails
    ca = compile(ta, None, str(tmp_path / "work"), "--no-cuda-graph -O1", "k")
    assert isinstance(ca, CompilationArtifacts) and not ca.captured and (tmp_path / "work" / "k.program.txt").exists()
    assert isinstance(compile_wlo(ta), WLOArtifacts) and b"captured_cuda_graph" in ca.get_neff_bytes()
    assert "--warmup=2" in append_default_compiler_flags("-O1") and append_default_compiler_flags("--warmup=5") == "--warmup=5"

    # --- builder → runtime model ----------------------------------------------------------------------------------------
    mb = ModelBuilder(m)
    with pytest.raises(ValueError):
        mb.compile()
    mb.trace(args=torch.randn(2, 4), tag="b2").trace(args=torch.randn(5, 4), tag="b5") \
      .trace(kwargs={"x": torch.randn(2, 4), "scale": torch.ones(1)}, tag="b2_scaled").trace(args=torch.randn(2, 4), tag="b2_alt")
    with pytest.raises(ValueError):
        mb.compile(priority_model_key="nope")
    with pytest.raises(ValueError):
        mb.compile(compiler_args={"b2": ""})
    nxd = mb.compile(priority_model_key="b5", compiler_workdir=str(tmp_path / "w2"))
    assert isinstance(nxd, NxDModel) and nxd.get_available_keys() == ["b2", "b5", "b2_scaled", "b2_alt"]
    x = torch.randn(2, 4)
    with pytest.raises(RuntimeError):
        nxd(x, model_name="b2")                                                                 # not initialised yet
    nxd.set_weights([m.state_dict()])
    nxd.to_neuron()
    assert nxd.loaded_on_neuron
    with pytest.raises(AssertionError):
        nxd(x)                                                                                  # b2 / b2_alt are ambiguous
    y, s = nxd(x, model_name="b2")
    torch.testing.assert_close(y, m.lin(x)) and torch.testing.assert_close(s, y.sum(-1))
    torch.testing.assert_close(nxd(torch.ones(5, 4))[0], m.lin(torch.ones(5, 4)))
    torch.testing.assert_close(nxd(x, scale=torch.full((1,), 2.0))[0], 2 * y)                   # kwargs routed + ordered
    torch.testing.assert_close(nxd(scale=torch.full((1,), 2.0), x=x)[0], 2 * y)
    with pytest.raises(AssertionError):
        nxd(torch.ones(5, 4), model_name="b2")                                                  # name contradicts the router
    with pytest.raises(KeyError):
        nxd(torch.ones(7, 4))
    with pytest.raises(KeyError):
        nxd(x, bogus=x)
    out = nxd([x], model_name="b2", forward_mode="ranked")
    assert len(out) == 2 and len(out[0]) == 1 and torch.equal(out[0][0], y)
    assert nxd([x], model_name="b2", forward_mode="ranked_to_cpu")[1][0].device.type == "cpu"
    # artefact getters, state / weight buffers, weight replacement
    assert nxd.get_hlo("b5")["inputs"] == (("x", (5, 4), "torch.float32"),) and nxd.get_metaneff("b2")["states"] == ["cache"]
    assert retrieve_artifact_from_model(nxd, "b2", "neff") == nxd.get_neff("b2")
    with pytest.raises(KeyError):
        nxd.get_hlo("nope")
    torch.testing.assert_close(nxd.read_from_neuron_buffer("cache", 0), y)                      # state written by the last call
    nxd.write_to_neuron_buffer(torch.ones(2, 3), "cache", 0)
    assert m.cache.eq(1).all()
    new = {k: v * 0 for k, v in m.state_dict().items()}
    nxd.replace_weights([new])
    assert nxd(x, model_name="b2")[0].abs().max() == 0 and m.lin.weight.abs().max() == 0        # in place: same tensors
    nxd.replace_weights([{"lin.weight": torch.eye(3, 4), "lin.bias": torch.zeros(3)}])
    # save / load round trip (module embedded; weights per rank)
    nxd.save(str(tmp_path / "saved"), save_weights=True)
    n2 = NxDModel.load(str(tmp_path / "saved"))
    n2.to_neuron()
    torch.testing.assert_close(n2(x, model_name="b2")[0], x[:, :3])
    assert sorted(n2.get_available_keys()) == sorted(nxd.get_available_keys())
    n3 = NxDModel.load(str(tmp_path / "saved"), model=_Toy())
    n3.to_neuron()
    torch.testing.assert_close(n3(x, model_name="b2")[0], x[:, :3])

    # helpers
    assert get_dtype_from_enum(get_dtype_enum(torch.bfloat16)) == torch.bfloat16
    with pytest.raises(ValueError):
        get_dtype_from_enum(999)
    assert generate_route_key_from_provided_args(ta.provided_args) == "x:(2, 4):torch.float32"
    vals, names = ts_convert_dict_to_ordered_list_type_tensor([("x", True), ("scale", False)], {"scale": 1}, 1)
    assert vals == [1] and names == ["x", "scale"]
    st = StateInitializer({"kv": [2, 3]}, {"kv": torch.bfloat16}, 1)()
    assert st[0]["kv"].shape == (2, 3) and st[0]["kv"].dtype == torch.bfloat16
    with pytest.raises(AssertionError):
        NxDModel(world_size=4, local_ranks_size=2)
    assert NxDModel(world_size=4, start_rank=2, local_ranks_size=2).start_rank == 2


def test_trace_namespace_v1_pieces():
    from neuronx_distributed_b200.trace import NxDModel, SPMDBucketModelScript
    from neuronx_distributed_b200.trace import hlo_utils
    from neuronx_distributed_b200.trace.mock_torchdist import mock_distributed
    from neuronx_distributed_b200.trace.model_builder import BaseModelInstance, JITWrapper, ModelContainer, get_hash_module
    from neuronx_distributed_b200.trace.spmd import BucketProgram, NxDModelExecutor, StateInitializer, default_bucket_kernel
    from neuronx_distributed_b200.trace.trace import (TensorParallelNeuronModel, collect_tp_neuron_models, create_local_weight,
                                                      create_local_weight_qkv, create_local_weight_with_expert_parallel,
                                                      find_unique_dtypes, generate_ranked_folder)

    lin = nn.Linear(4, 2)
    progs = [BucketProgram("k", lin, lambda mod, t: mod(t), (torch.zeros(n, 4),), use_cuda_graph=False) for n in (1, 3)]
    script = SPMDBucketModelScript(progs)
    x = torch.randn(3, 4)
    torch.testing.assert_close(script([x], torch.tensor(1)), lin(x))
    assert len(script.forward_ranked([[x], [x]], torch.tensor(1))) == 2
    with pytest.raises(ValueError):
        script([x], torch.tensor(5))
    inp, idx = default_bucket_kernel([x])
    assert inp[0] is x and int(idx) == 0
    nxd = NxDModel()
    for p in progs:
        nxd.add_program(p)
    torch.testing.assert_close(NxDModelExecutor(nxd)(x), lin(x))
    torch.testing.assert_close(collect_tp_neuron_models([lambda t: lin(t)])(x), lin(x))
    assert isinstance(collect_tp_neuron_models([lin]), TensorParallelNeuronModel)
    assert generate_ranked_folder(3, 0, 1) == "tp_3" and generate_ranked_folder(3, 1, 2) == "tp_3_bk_1"
    assert find_unique_dtypes(lin) == {torch.float32: 2}
    inst = BaseModelInstance(lambda: lin, {0: 1})
    assert inst.get()[0] is lin and ModelContainer(inst, [(x,)]).example_inputs == [(x,)]
    assert JITWrapper(lambda v: v + 1)(1) == 2 and len(get_hash_module("abc")) == 64
    # combined K/V allocation: K and V of a layer are views of one [2, …] buffer
    si = StateInitializer({"m.past_key_values.0": [2, 4], "m.past_key_values.1": [2, 4], "other": [3]},
                          {"m.past_key_values.0": torch.float32, "m.past_key_values.1": torch.float32, "other": torch.int32},
                          1, combine_kv_on_device=True)
    st = si()[0]
    comb = st["m.past_key_values.combined.0"]
    assert comb.shape == (2, 2, 4) and st["m.past_key_values.1"].data_ptr() == comb[1].data_ptr() and st["other"].dtype == torch.int32
    st["m.past_key_values.0"].fill_(7)
    assert comb[0].eq(7).all() and comb[1].eq(0).all()
    # sharding primitives with the reference's argument order
    w = torch.arange(24.0).view(12, 2)
    assert torch.equal(create_local_weight(1, 2, w, 0, 6, 1), w[6:])
    assert torch.equal(create_local_weight(1, 2, w, 0, 6, 2), torch.cat([w[3:6], w[9:12]]))     # stride 2: gate|up halves
    qkv = create_local_weight_qkv(1, 2, w, 0, 8, 2)
    assert torch.equal(qkv, torch.cat([w[4:8], w[9:10], w[11:12]]))
    e = torch.arange(2 * 4 * 6.0).view(2 * 2, 2, 6)
    got = create_local_weight_with_expert_parallel(0, 2, e, 2, 3, 1, [1, 3])
    assert torch.equal(got, e[[1, 3]][:, :, :3])
    f8 = torch.randn(4, 2, 6).to(torch.float8_e4m3fn)
    assert create_local_weight_with_expert_parallel(1, 2, f8, 2, 3, 1, [0]).dtype == torch.float8_e4m3fn
    # the layout-transformation hooks work on launch plans (tests/test_launch_plan_cpu.py); without a transformer they are identities
    wts = {"a": torch.ones(1)}
    assert hlo_utils.update_weight(wts) is wts
    with pytest.raises(FileNotFoundError):
        hlo_utils.read_hlo("x.pb")
    with mock_distributed(8) as d:
        import torch.distributed as td
        assert td.is_initialized() and td.get_world_size() == 8 and td.get_rank() == 0
        g = td.new_group([0, 2, 4, 6])
        assert td.get_world_size(g) == 4 and d.get_process_group_ranks(g) == [0, 2, 4, 6]
    import torch.distributed as td
    assert not isinstance(td, type(d))


def _shard_inplace(rank, world):
    """``get_sharded_checkpoint`` (in place, per rank) over TP layers, expert-parallel weights and quantised layers."""
    from neuronx_distributed_b200.modules.moe.moe_parallel_layers import ExpertFusedColumnParallelLinear
    from neuronx_distributed_b200.parallel_layers import ColumnParallelLinear, RowParallelLinear
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.quantization.quantization_layers import QuantizedColumnParallel
    from neuronx_distributed_b200.quantization.quantization_utils import quantize_pytorch_model_per_channel_symmetric
    from neuronx_distributed_b200.trace.trace import get_sharded_checkpoint, invoke_preshard_hook, preprocess_checkpoint

    ps.initialize_model_parallel(tensor_model_parallel_size=1, expert_model_parallel_size=world)   # dp = world, ep = world
    torch.manual_seed(0)
    experts = ExpertFusedColumnParallelLinear(4, 8, 6)
    assert experts.weight.shape == (4 // world, 8, 6) and getattr(experts.weight, "expert_model_parallel", world == 1)
    full = {"weight": torch.randn(4, 8, 6)}
    sd = dict(full)
    get_sharded_checkpoint(sd, experts, rank, world)
    torch.testing.assert_close(sd["weight"], full["weight"][experts.local_expert_ids])
    ps.destroy_model_parallel()

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    m = nn.Sequential()
    m.add_module("up", ColumnParallelLinear(8, 12, bias=True, gather_output=False, pad=False))
    m.add_module("q", QuantizedColumnParallel(8, 10, bias=False, gather_output=True, pad=True, quantization_type="per_channel_symmetric"))
    m.add_module("down", RowParallelLinear(12, 8, bias=False, input_is_parallel=True))
    ql = quantize_pytorch_model_per_channel_symmetric(nn.Sequential(nn.Linear(8, 10, bias=False))).state_dict()
    full = {"up.weight": torch.randn(12, 8), "up.bias": torch.randn(12), "down.weight": torch.randn(8, 12),
            "q.weight": ql["0.weight"], "q.scale": ql["0.scale"], "junk.weight": torch.zeros(1)}
    sd = dict(full)
    with pytest.warns(UserWarning, match="redundant"):
        get_sharded_checkpoint(sd, m, rank, world)
    assert "junk.weight" not in sd
    per = 12 // world
    torch.testing.assert_close(sd["up.weight"], full["up.weight"][rank * per:(rank + 1) * per])
    torch.testing.assert_close(sd["up.bias"], full["up.bias"][rank * per:(rank + 1) * per])
    torch.testing.assert_close(sd["down.weight"], full["down.weight"][:, rank * per:(rank + 1) * per])
    pad_rows = (10 + m.q.pad_size) // world                                    # preshard hook padded 10 → multiple of world
    assert sd["q.weight"].shape == (pad_rows, 8) and sd["q.scale"].shape == (pad_rows, 1) and sd["q.weight"].dtype == torch.int8
    m.load_state_dict(sd)
    bad = {"up.weight": torch.randn(10, 8)}
    with pytest.raises((RuntimeError, AssertionError)):
        get_sharded_checkpoint(bad, m, rank, world, is_cached=True)
    ck = {"q.weight": ql["0.weight"].clone(), "q.scale": ql["0.scale"].clone()}
    invoke_preshard_hook(m, ck, "")
    assert ck["q.weight"].shape[0] == 10 + m.q.pad_size
    preprocess_checkpoint(m, {"up.weight": full["up.weight"]})


def test_get_sharded_checkpoint_tp_ep_quantized():
    run_distributed(_shard_inplace, 2, timeout=120)


def test_autobucketing_spacing_and_routers():
    import torch

    from neuronx_distributed_b200.inference import autobucketing as ab

    assert ab.generate_buckets(128, 1024) == [128, 256, 512, 1024]
    assert ab.generate_buckets(128, 1500) == [128, 256, 512, 1024, 1500]
    assert ab.generate_buckets(128, 1100) == [128, 256, 512, 1100]            # 1024 is within sqrt(2) of 1100: dropped
    assert ab.generate_buckets(512, 512) == [512] and ab.pick_bucket([128, 256], 130) == 256
    buckets = [4, 8, 16]
    ids = torch.zeros(2, 16, dtype=torch.long)
    ids[0, :3], ids[1, :6] = 7, 9                                             # right padded: 3 and 6 real tokens
    mask = (ids != 0).long()
    seq_ids = torch.arange(2)
    out, idx = ab.context_encoding_router([ids, mask, seq_ids], buckets, "right", pad_token=0)
    assert idx == 1 and out[0].shape == (2, 8) and out[1].shape == (2, 8) and out[2] is seq_ids
    assert torch.equal(out[0], ids[:, :8])
    lids = torch.zeros(2, 16, dtype=torch.long)
    lids[0, -3:], lids[1, -6:] = 7, 9                                         # left padded
    out, idx = ab.get_context_encoder_bk()([lids, seq_ids], buckets, "left", 0)
    assert idx == 1 and torch.equal(out[0], lids[:, 8:16])
    # token generation: positions 3 and 7 → needs 8 cache slots → bucket 8; position 8 → bucket 16
    tok, pos = torch.ones(2, 1, dtype=torch.long), torch.tensor([[3], [7]])
    am = torch.ones(2, 16, dtype=torch.long)
    out, idx = ab.token_generation_router([tok, am, pos], buckets, "right")
    assert idx == 1 and out[1].shape == (2, 8) and out[0] is tok
    out, idx = ab.get_token_generation_bk()([tok, torch.tensor([[3], [8]])], buckets, "right")
    assert idx == 2
