"""Launch plans (``inference/launch_plan.py``): the model-code-free inference artefact and the plan passes that take the role
of the reference's HLO surgery (``trace/hlo_utils.py``) and TorchScript export (``trace/nxd_model/nxd_model.py:709-969``)."""
import os

import pytest
import torch
import torch.nn.functional as F
from torch import nn

from dist_utils import run_distributed


class _Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.l1, self.l2 = nn.Linear(16, 32), nn.Linear(32, 16)
        self.master = nn.Parameter(torch.randn(16, 16, dtype=torch.float64))      # used through a cast: hoistable
        self.register_buffer("cache", torch.zeros(4, 8, 16))
        self.table = torch.randn(8, 16)                                           # a plain attribute: anonymous constant

    def forward(self, x, slot):
        h = F.silu(self.l1(x))
        y = self.l2(h) + self.table[: x.shape[0]] + x @ self.master.float().t()
        mask = torch.tril(torch.ones(8, 8)).bool()                                # input-independent: hoistable
        self.cache.index_copy_(1, slot, y[:4].unsqueeze(1))
        att = (y @ y.t()).masked_fill(~mask, float("-inf")).softmax(-1)
        _unused = att * 3                                                          # dead code
        return att @ y, self.cache.sum()


def test_record_replay_passes_and_persistence(tmp_path):
    from neuronx_distributed_b200.inference import launch_plan as lp

    torch.manual_seed(0)
    m = _Block()
    x, slot = torch.randn(8, 16), torch.tensor([3])
    plan = lp.record(m, [x, slot], ["x", "slot"])
    s = plan.summary()
    assert s["inputs"] == [("x", (8, 16), "float32"), ("slot", (1,), "int64")] and s["baked_scalars"] == 0
    names = {c.name for c in plan.constants.values()}
    assert {"l1.weight", "l1.bias", "l2.weight", "l2.bias", "master", "cache"} <= names
    assert plan.state_names() == ["cache"]                                        # written in place → state, not a weight
    anon = [c for c in plan.constants.values() if not c.named]
    assert len(anon) == 1 and anon[0].shape == (8, 16)                            # self.table
    # replay == module, for new inputs too; state is shared with the module (same tensors)
    for xi, si in ((x, slot), (torch.randn(8, 16), torch.tensor([5]))):
        m.cache.zero_()
        a = plan(xi, si)
        got_cache = m.cache.clone()
        m.cache.zero_()
        b = m(xi, si)
        torch.testing.assert_close(a[0], b[0]) and torch.testing.assert_close(a[1], b[1])
        torch.testing.assert_close(got_cache, m.cache)
    torch.testing.assert_close(plan(slot=slot, x=x)[0], m(x, slot)[0])            # keyword call, any order
    with pytest.raises(lp.PlanError):
        plan(torch.randn(7, 16), slot)                                            # a plan is one bucket: shapes are fixed
    # usage map: l1.weight is read by the addmm (through its transpose view)
    use = plan.weight_usage()
    assert any(t.startswith("aten::addmm") for _, t in use["l1.weight"]) and use["cache"]
    assert not plan.calls_extension() and plan.kernel_weight_names() == []
    # DCE removes the dead multiply, keeps the in-place cache write
    n0 = len(plan.nodes)
    assert plan.dce() >= 1 and len(plan.nodes) < n0
    assert any(n.target.startswith("aten::index_copy_") for n in plan.nodes)
    m.cache.zero_()
    torch.testing.assert_close(plan(x, slot)[1], m(x, slot)[1])
    # layout transformer: casts / transposes of frozen weights and the input-independent mask leave the per-call plan
    transformer, main, tmap = plan.hoist_weight_only()
    hoisted = {n.target for n in transformer.nodes}
    assert "aten::tril.default" in hoisted and "aten::_to_copy.default" in hoisted and "aten::t.default" in hoisted
    assert tmap["master"][0] == "aten::_to_copy.default" and "cache" not in tmap
    assert not any(n.target.startswith(("aten::tril", "aten::ones", "aten::_to_copy")) for n in main.nodes)
    assert len(main.nodes) + len(transformer.nodes) == len(plan.nodes)
    main.apply_transformer(transformer)
    torch.testing.assert_close(main(x, slot)[0], m(x, slot)[0])
    # weights change in place → re-run the transformer; derived constants keep their addresses (CUDA-graph safe)
    derived = {i: t for i, t in main.tensors.items() if main.constants[i].name.startswith("_derived_")}
    with torch.no_grad():
        m.master.mul_(2.0)
    main.apply_transformer(transformer)
    assert all(main.tensors[i] is t for i, t in derived.items())
    torch.testing.assert_close(main(x, slot)[0], m(x, slot)[0])
    # skip list: the weight stays in the per-call plan
    t2, m2, tmap2 = plan.hoist_weight_only(skip={"master"})
    assert "master" not in tmap2 and any(c.name == "master" for c in m2.constants.values())
    # disk round trip of (per-call plan, transformer); constants are shared by name, derived ones are recomputed
    lp.save_plans(str(tmp_path / "art"), {"b": main, "__lt__b": transformer})
    plans, tensors, _ = lp.load_plans(str(tmp_path / "art"))
    assert not any(k.startswith("_derived_") for k in tensors)
    assert plans["b"].named_constants()["cache"] is plans["__lt__b"].named_constants().get("cache", plans["b"].named_constants()["cache"])
    plans["b"].apply_transformer(plans["__lt__b"])
    x2 = torch.randn(8, 16)
    torch.testing.assert_close(plans["b"](x2, slot)[0], m(x2, slot)[0])
    # JSON stability
    assert lp.LaunchPlan.from_json(plans["b"].to_json()).to_json() == plans["b"].to_json()


def test_hlo_utils_roles_on_plans(tmp_path):
    """``trace/hlo_utils.py`` entry points: marking, extraction, per-weight transforms, checkpoint transformation on disk."""
    from neuronx_distributed_b200.trace.functions import trace
    from neuronx_distributed_b200.trace import hlo_utils as hu
    from neuronx_distributed_b200.utils.safetensors_utils import load_state_dict_safetensors, save_state_dict_safetensors

    torch.manual_seed(0)
    m = _Block()
    x, slot = torch.randn(8, 16), torch.tensor([3])
    ta = trace(m, (x, slot))
    with pytest.raises(RuntimeError, match="Invalid weights"):
        hu.mark_weights_for_wlo(ta, {"nope"})
    hu.mark_weights_for_wlo(ta, {"l2.weight"})
    plan = ta.record_plan()
    assert plan.meta[hu.TRANSPOSABLE_WEIGHT_IDX] == sorted(i for n, i in ta.weight_name_to_idx.items() if n != "l2.weight")
    transformer, main = hu.extract_weight_layout_transform_hlo(ta)
    tmap = hu.get_layout_transform_map(transformer)
    assert "master" in tmap and "l2.weight" not in tmap and "l1.weight" in tmap            # skip list honoured
    assert hu.get_layout_transform_map(transformer, ta.weight_name_to_idx).keys() == {ta.weight_name_to_idx[n] for n in tmap
                                                                                      if n in ta.weight_name_to_idx}
    # per-weight transform callables (reference get_wlt_map): master → its fp32 copy (and the transposed view of it)
    wlt = hu.get_wlt_map(transformer)
    outs = wlt["master"](m.master.detach())
    assert any(o.dtype == torch.float32 and o.shape == (16, 16) for o in outs)
    # plans on disk, merged back together == the original program
    hu.write_hlo(str(tmp_path / "p" / "main.json"), main)
    assert len(hu.read_hlo(str(tmp_path / "p" / "main.json")).nodes) == len(main.nodes)
    merged = hu.append_layout_computation_to_hlo(main, transformer)
    torch.testing.assert_close(merged(x, slot)[0], m(x, slot)[0])
    # usage analysis
    use = hu.prepare_parameter_usage_map(plan, ["l1.weight", "cache"])
    assert set(use) == {"l1.weight", "cache"} and all(use.values())
    view_out = next(o for n in plan.nodes if n.target == "aten::t.default" for o in n.outs)
    assert hu.traceback_instruction_to_parameter(plan, view_out) in {"l1.weight", "l2.weight"}
    assert not hu.is_nki_kernel_called(plan) and hu.get_nki_kernel_weight_names(plan) == set()
    assert hu.get_input_order(ta) == ["x", "slot"] and hu.prepare_metaneff_for_wlt_hlo(transformer)["outputs"]
    # sharded checkpoint on disk → derived tensors next to it
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    save_state_dict_safetensors(sd, str(tmp_path / "tp0_sharded_checkpoint.safetensors"))
    done = hu.transform_weight_layout_on_cpu(transformer, None, 0, 1, str(tmp_path))
    on_disk = load_state_dict_safetensors(str(tmp_path / "tp0_derived.safetensors"))
    assert set(on_disk) == set(done[0]) and all(k.startswith("_derived_") for k in on_disk)
    # the builder-level switch (reference $NXD_LAYOUT_TRANSFORMATION_OPTIONS): unset → transform at load; CPU / device → serialise now
    from neuronx_distributed_b200.trace.model_builder import ModelBuilder as BuilderV1

    b1 = BuilderV1(tp_degree=1)
    assert b1.transform_weight_layout_with_overriden_option(str(tmp_path), transformer) is None
    os.remove(str(tmp_path / "tp0_derived.safetensors"))
    again = b1.transform_weight_layout_with_overriden_option(str(tmp_path), transformer, option=hu.NXD_LAYOUT_ON_CPU_AND_SERIALIZE)
    assert set(again[0]) == set(done[0]) and os.path.exists(str(tmp_path / "tp0_derived.safetensors"))
    os.environ[hu.NXD_LAYOUT_TRANSFORMATION_OPTIONS] = hu.NXD_LAYOUT_ON_DEVICE_AND_SERIALIZE
    try:
        os.remove(str(tmp_path / "tp0_derived.safetensors"))
        b1.transform_weight_layout_with_overriden_option(str(tmp_path), transformer)              # no GPU here: runs on the host
        redo = load_state_dict_safetensors(str(tmp_path / "tp0_derived.safetensors"))
        assert all(torch.equal(redo[k], on_disk[k]) for k in on_disk)
        os.environ[hu.NXD_LAYOUT_TRANSFORMATION_OPTIONS] = "SOMETHING_ELSE"
        with pytest.raises(ValueError):
            b1.transform_weight_layout_with_overriden_option(str(tmp_path), transformer)
    finally:
        del os.environ[hu.NXD_LAYOUT_TRANSFORMATION_OPTIONS]
    full = hu.update_weight(sd, transformer)
    want = m.master.detach().float()                               # what the per-call plan consumes is the transposed cast
    assert any(v.shape == want.shape and (torch.equal(v, want) or torch.equal(v, want.t()))
               for k, v in full.items() if k.startswith("_derived_"))
    # non-priority bucket: same pass, per-call plan stored back on the trace artifacts
    ta2 = trace(m, (torch.randn(8, 16), torch.tensor([1])))
    t2, main2 = hu.apply_layout_transformation(ta2, ta)
    assert ta2.record_plan() is main2 and "l2.weight" not in t2.meta["layout_transform_map"]
    hu.cleanup_after_layout_transformation(ta)
    assert hu.TRANSPOSABLE_WEIGHT_IDX not in plan.meta
    assert hu.get_executable_full_qualified_path("nvcc")


def test_plan_records_extension_kernels_and_plan_ops(monkeypatch):
    """Kernel calls go through ``ops._ext.ext()``: a recording sees them via the proxy; resource-bound entry points are
    refused outside a ``plan_op``; ``plan_op`` functions become single nodes that are called again at replay."""
    from neuronx_distributed_b200.inference import launch_plan as lp
    from neuronx_distributed_b200.ops import _ext
    from neuronx_distributed_b200.utils.plan_registry import plan_op

    calls = {"scale": 0, "res": 0}

    class FakeExt:
        @staticmethod
        def fake_scale(x, alpha):
            calls["scale"] += 1
            return torch.empty_like(x).copy_(x * alpha)             # allocations inside a kernel call are not recorded

        @staticmethod
        def nvls_fake(x, ptrs):
            return x

    monkeypatch.setattr(_ext, "_C", FakeExt)
    monkeypatch.setattr(_ext, "_TRIED", True)

    @plan_op("test.with_resources", pure=True)
    def with_resources(x, factor):
        calls["res"] += 1
        return _ext.ext().nvls_fake(x, [1 << 45]) * factor          # raw pointers stay inside the op

    def model(x):
        y = _ext.ext().fake_scale(x + 1, 2.0)
        return with_resources(y, 3)

    x = torch.randn(4)
    plan = lp.record(model, [x])
    assert [n.kind for n in plan.nodes] == ["op", "ext", "py"] and plan.nodes[1].target == "fake_scale"
    assert plan.calls_extension() and plan.meta["py_modules"] == [__name__]
    calls.update(scale=0, res=0)
    torch.testing.assert_close(plan(x), (x + 1) * 6)
    assert calls == {"scale": 1, "res": 1}
    assert _ext.ext() is FakeExt                                     # proxy removed after the recording

    def bad(x):
        return _ext.ext().nvls_fake(x, [1 << 45])

    with pytest.raises(lp.PlanError, match="process-local resources"):
        lp.record(bad, [x])
    assert _ext.ext() is FakeExt and not lp.recording()


class _Wrap(nn.Module):
    def __init__(self, model, which):
        super().__init__()
        self.m, self.which = model, which

    def forward(self, input_ids, aux):
        if self.which == "cte":
            return self.m.context_encoding(input_ids, aux)
        return self.m.token_generation(input_ids, aux)


def _portable_llama(rank, world, tmp):
    from neuronx_distributed_b200.trace.functions import compile as ncompile
    from neuronx_distributed_b200.trace.functions import compile_layout_transformer, compile_wlo, trace
    from neuronx_distributed_b200.trace.nxd_model import NxDModel, TorchScriptNxDModel, convert_nxd_model_to_torchscript_model
    from neuronx_distributed_b200.models.llama import LlamaConfig
    from neuronx_distributed_b200.models.llama_inference import LlamaForInference
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    cfg = LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, dtype=torch.float32, max_position_embeddings=32)
    torch.manual_seed(0)
    m = LlamaForInference(cfg, batch_size=2, max_seq_len=32).eval()
    ids = torch.randint(0, 64, (2, 8), generator=torch.Generator().manual_seed(1))
    last = torch.tensor([7, 5])
    want = m.generate(ids, 6, prompt_lens=last + 1)

    cte, tkg = _Wrap(m, "cte"), _Wrap(m, "tkg")
    ta_c, ta_t = trace(cte, (ids, last)), trace(tkg, (torch.zeros(2, 1, dtype=torch.long), torch.tensor([8, 6])))
    nxd = NxDModel(world_size=world)
    wlo = compile_wlo(ta_t, None, None, None, "tkg")
    nxd.add("cte", ta_c, ncompile(ta_c, None, None, "--plan", "cte")).add("tkg", ta_t, wlo)
    assert len(wlo.transformer.nodes) > 0 and compile_layout_transformer(wlo).transformers["tkg"][0] is wlo.transformer
    nxd.to_neuron()
    summ = wlo.plan.summary()
    if world > 1:
        assert summ["py_ops"] > 0                                   # TP collectives are plan ops, stored by group name

    def generate(model, n):
        tok = model(ids, last, model_name="cte")
        out, pos = [tok], last + 1
        for _ in range(n - 1):
            tok = model(tok.view(2, 1), pos, model_name="tkg")
            out.append(tok)
            pos = pos + 1
        return torch.stack(out, dim=1)

    assert torch.equal(generate(nxd, 6), want)
    ts = convert_nxd_model_to_torchscript_model(nxd)
    assert isinstance(ts, TorchScriptNxDModel) and torch.equal(generate(ts, 6), want)
    path = f"{tmp}/portable"
    ts.save(path)
    ps_names = {n for n, _ in m.named_parameters()}
    del ts, nxd, wlo, ta_c, ta_t, cte, tkg
    loaded = NxDModel.load(path)                                     # no model object, no model class
    loaded.to_neuron()
    assert sorted(loaded.get_available_keys()) == ["cte", "tkg"]
    assert torch.equal(generate(loaded, 6), want)
    # a loaded artefact still supports weight replacement by name: zero the lm_head → the sampler returns token 0
    name = next(n for n in loaded._named_state()[0] if "lm_head" in n)
    assert name.split("m.", 1)[1] in ps_names
    loaded.replace_weights([{name: torch.zeros_like(loaded._named_state()[0][name])}])
    assert int(generate(loaded, 2).abs().max()) == 0


def test_portable_llama_artifact_tp1(tmp_path):
    run_distributed(_portable_llama, 1, str(tmp_path))


def test_portable_llama_artifact_tp2(tmp_path):
    run_distributed(_portable_llama, 2, str(tmp_path))
