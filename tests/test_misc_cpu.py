"""Operators, sampling, capture/replacement, CP batch split, activation checkpoint wrapper, meta-device init, EP MoE."""
import pytest
import os

import torch
from torch import nn

from dist_utils import run_distributed


def _ops(rank, world):
    from neuronx_distributed_b200.operators import argmax, topk
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.utils.sampling import Sampler

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    torch.manual_seed(0)
    full = torch.randn(3, 5, 32)
    local = full.chunk(world, -1)[rank]
    assert torch.equal(argmax(local, dim=-1), full.argmax(-1))
    v, i = topk(local, 4, dim=-1)
    rv, ri = full.topk(4, -1)
    torch.testing.assert_close(v, rv)
    assert torch.equal(i, ri)
    s = Sampler(top_k=1, vocab_parallel=True)
    assert torch.equal(s.sample(local[:, 0]), full[:, 0].argmax(-1))
    s2 = Sampler(top_k=5, top_p=0.9, temperature=0.7, do_sample=True, vocab_parallel=True)
    g = torch.Generator().manual_seed(3)
    tok = s2.sample(local[:, 0], generator=g)
    assert tok.shape == (3,) and all(int(t) in rv_row.tolist() for t, rv_row in zip(tok, full[:, 0].topk(5, -1).indices))


def test_distributed_argmax_topk_sampling():
    run_distributed(_ops, 2, timeout=90)


def test_capture_and_replacement():
    from neuronx_distributed_b200.utils import tensor_capture as tc
    from neuronx_distributed_b200.utils import tensor_replacement as tr

    m = nn.Sequential(nn.Linear(4, 4), nn.ReLU(), nn.Linear(4, 2))
    x = torch.randn(3, 4)
    tc.enable_tensor_capture(m, ["0", "2"])
    y = m(x)
    cap = tc.get_captured_tensors()
    torch.testing.assert_close(cap["2.outputs"], y)
    assert "0.outputs" in cap
    tc.disable_tensor_capture()
    inj = torch.ones(3, 4)
    tr.enable_tensor_replacement(m, {"0": inj})
    y2 = m(x)
    torch.testing.assert_close(y2, m[2](torch.relu(inj)))
    tr.disable_tensor_replacement()
    torch.testing.assert_close(m(x), y)

    # reference-style API: inputs, nested outputs, manual tensors with a budget, ordering, hooks removed on disable
    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = nn.Linear(4, 4)

        def forward(self, h, scale=None):
            out = self.lin(h) * (1 if scale is None else scale)
            tc.register_tensor("pre_act", out)
            tc.register_tensor("pre_act", out + 1)                     # same name twice → suffixed key
            tc.register_tensor("over_budget", out)
            return {"hidden": torch.relu(out), "aux": (out, out * 2)}

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = Block(), nn.Linear(4, 2)

        def forward(self, h):
            return self.b(self.a(h, scale=torch.tensor(2.0))["hidden"])

    net = Net()
    assert tc.get_available_modules(net) == ["a", "a.lin", "b"]
    with pytest.raises(ValueError):
        tc.enable_tensor_capture(net, ["nope"])
    net = tc.enable_tensor_capture(net, ["b", "a"], max_tensors=2, capture_inputs=True)
    out = net(x)
    d = tc.get_captured_tensors_dict()
    assert list(d) == ["b.inputs.0", "b.outputs", "a.inputs.0", "a.inputs.kwargs.scale", "a.outputs.aux.0", "a.outputs.aux.1",
                       "a.outputs.hidden", "manual_pre_act", "manual_pre_act_1"], list(d)
    torch.testing.assert_close(d["b.outputs"], out)
    torch.testing.assert_close(d["a.outputs.aux.1"], 2 * d["a.outputs.aux.0"])
    torch.testing.assert_close(d["manual_pre_act_1"], d["manual_pre_act"] + 1)
    reg = tc.TensorRegistry.get_instance()
    assert reg.get_manual_tensor_count() == 2 and reg.get_monitored_tensor_count() == 7 and reg.get_total_tensor_count() == 9
    reg.model_info.manual_tensors["l0.moe_auto"] = torch.zeros(3); reg.model_info.manual_tensors["l1.moe_auto"] = torch.ones(3)
    assert reg.get_manual_tensors()["auto_moe_stats.expert_index"].shape == (2, 3)
    assert tc.get_captured_tensors(clear=True) and not tc.get_captured_tensors_dict()
    net = tc.disable_tensor_capture(net)
    net(x)
    assert not tc.get_captured_tensors_dict() and not reg.model_info.hooks

    # golden-side capture of an eager model: one file per (phase, step, tensor)
    import tempfile, os
    with tempfile.TemporaryDirectory() as td:
        net = tc.modify_hf_eager_model_for_tensor_capture(net, ["b"], tensor_capture_save_dir=td)
        net(x); net(x)
        files = sorted(os.listdir(td))
        assert files == ["captured_tensors_cte_step_1_module_b.outputs.pt", "captured_tensors_tkg_step_2_module_b.outputs.pt"], files
        torch.testing.assert_close(torch.load(os.path.join(td, files[0])), out)
        net = tc.restore_model(net)
        net(x)
        assert len(os.listdir(td)) == 2

    # replacement as trailing (tensor, mask) arguments: the same prepared model replaces nothing / one layer / part of a layer
    tr.RuntimeRegister.module_superset = ["0", "2"]
    m2, hooks = tr.modify_model_for_tensor_replacement(m)
    zeros0, zeros2 = torch.zeros(3, 4), torch.zeros(3, 2)
    off = torch.zeros((), dtype=torch.bool)
    torch.testing.assert_close(m2(x, zeros0, zeros2, off, off), y)
    torch.testing.assert_close(m2(x, inj, zeros2, ~off, off), y2)
    part = torch.zeros(3, 2, dtype=torch.bool); part[0] = True
    got = m2(x, zeros0, torch.full((3, 2), 7.0), off, part)
    assert (got[0] == 7).all() and torch.allclose(got[1:], y[1:])
    with pytest.raises(ValueError):
        m2(x, zeros0, off)
    assert not tr.RuntimeRegister._tr_runtime_list                      # cleared after every forward
    for h in hooks.values():
        h.remove()
    m.forward = m._nxd_tr_original_forward
    tr.RuntimeRegister.module_superset = []


def test_medusa_buffers():
    from neuronx_distributed_b200.utils.medusa_utils import generate_medusa_buffers

    b = generate_medusa_buffers([[0], [1], [0, 0], [0, 1], [1, 0]], topk=10)
    assert b["medusa_attn_mask"].shape == (1, 1, 6, 6)
    assert b["tree_indices"].tolist() == [0, 1, 2, 11, 12, 11]
    assert b["medusa_position_ids"].tolist() == [0, 1, 1, 2, 2, 2]
    assert b["retrieve_indices"].shape[1] == 3


def _cp_batch(rank, world):
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.utils.batch_utils import get_batch_on_this_context_parallel_rank

    ps.initialize_model_parallel(1, 1, context_parallel_size=world)
    ids = torch.arange(16).view(2, 8)
    out = get_batch_on_this_context_parallel_rank({"input_ids": ids, "labels": ids.clone()})
    assert out["input_ids"].shape == (2, 4)
    torch.testing.assert_close(out["input_ids"], ids[:, rank * 4:(rank + 1) * 4])
    want = torch.cat([ids[:, 1:], torch.full((2, 1), -100)], 1)[:, rank * 4:(rank + 1) * 4]
    torch.testing.assert_close(out["labels"], want)


def test_context_parallel_batch_split():
    run_distributed(_cp_batch, 2, timeout=60)


def test_activation_checkpoint_wrapper_keeps_names():
    from neuronx_distributed_b200.utils.activation_checkpoint import NxDCheckpointWrapper, apply_activation_checkpointing

    m = nn.Sequential(nn.Linear(4, 4), nn.Tanh(), nn.Linear(4, 4))
    ref = {k: v.clone() for k, v in m.state_dict().items()}
    apply_activation_checkpointing(m, check_fn=lambda mod: isinstance(mod, nn.Linear))
    assert isinstance(m[0], NxDCheckpointWrapper)
    assert set(m.state_dict().keys()) == set(ref.keys())
    x = torch.randn(2, 4, requires_grad=True)
    m(x).sum().backward()
    assert x.grad is not None
    m2 = nn.Sequential(nn.Linear(4, 4), nn.Tanh(), nn.Linear(4, 4))
    apply_activation_checkpointing(m2, check_fn=lambda mod: isinstance(mod, nn.Linear))
    m2.load_state_dict(ref)
    torch.testing.assert_close(m2(x), m(x))


def _meta(rank, world):
    import neuronx_distributed_b200 as nxd
    from neuronx_distributed_b200.models.llama import LlamaConfig, LlamaForCausalLM

    def init_fn(module, device):
        for p in module.parameters(recurse=False):
            nn.init.normal_(p, std=0.02)

    cfg = nxd.neuronx_distributed_config(tensor_parallel_size=world,
                                         model_init_config={"meta_device_init": True, "param_init_fn": init_fn, "sequential_move_factor": 11})
    mcfg = LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=4,
                       dtype=torch.float32, max_position_embeddings=16)
    model = nxd.initialize_parallel_model(cfg, lambda: LlamaForCausalLM(mcfg))
    ps_ = [p for p in model.parameters()]
    assert all(p.device.type != "meta" for p in ps_)
    w = model.module.lm_head.weight
    assert getattr(w, "tensor_model_parallel", False) and w.partition_dim == 0
    ids = torch.randint(0, 64, (2, 16))
    loss = model.run_train(input_ids=ids, labels=ids)
    assert torch.isfinite(loss)


def test_meta_device_init_tp2():
    run_distributed(_meta, 2, timeout=90)


def _ep(rank, world):
    from neuronx_distributed_b200.modules.moe import ExpertMLPsV2, MoE, RoutedExpertsMLPOpsConfig, RouterTopK
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=1, expert_model_parallel_size=2)
    assert ps.get_expert_model_parallel_size() == 2 and ps.get_data_parallel_size() == world
    torch.manual_seed(0)
    E, k, H, I = 4, 2, 16, 32
    cfg = RoutedExpertsMLPOpsConfig(normalize_top_k_affinities=True, num_experts=E, top_k=k, hidden_size=H, intermediate_size=I, capacity_factor=4.0)
    layer = MoE(RouterTopK(E, k, H), ExpertMLPsV2(cfg))
    w = layer.expert_mlps.mlp_op.down_proj.weight
    assert w.shape[0] == E // 2 and getattr(w, "expert_model_parallel", False)
    torch.manual_seed(10 + rank)             # different tokens on every DP rank
    x = torch.randn(12, 1, H, requires_grad=True)
    (y,) = layer(x)
    assert y.shape == x.shape
    y.pow(2).mean().backward()
    assert x.grad is not None and w.grad is not None and torch.isfinite(w.grad).all()
    # checkpoint-side rank tables (reference expert_mlps_v2.py:1501-1530): which logical experts every global rank hosts
    from neuronx_distributed_b200.modules.moe.expert_mlps_v2 import create_spmd_ranks

    sd = {}
    create_spmd_ranks(sd, "m.", world, E, ps.get_expert_model_parallel_group(), "spmd_rank")
    assert sd["m.spmd_rank.rank"].tolist() == [0, 1] and sd["m.spmd_rank.local_expert_indices"].tolist() == [[0, 1], [2, 3]]
    create_spmd_ranks(sd, "m.", world, E, ps.get_expert_model_parallel_group(), "spmd_rank", expert_distribution=[[0, 3], [1, 2]])
    assert sd["m.spmd_rank.local_expert_indices"].tolist() == [[0, 3], [1, 2]]
    # a residual passed to the layer is added first and handed back (the stream the next block adds to)
    layer.eval()
    with torch.no_grad():
        res = torch.randn_like(x)
        y2, stream = layer(x.detach(), residual=res)
        (y3,) = layer(x.detach() + res)
    torch.testing.assert_close(y2, y3)
    torch.testing.assert_close(stream, x.detach() + res)


def test_expert_parallel_all_to_all_training():
    run_distributed(_ep, 2, timeout=90)


def test_small_utils(tmp_path):
    """Medusa step helpers, serialization helpers, label shift, checkpoint_wrapper, duplicate-tensor check, timeline,
    autocast casting helpers, shared-weight helpers."""
    import json
    import logging
    from collections import namedtuple

    from neuronx_distributed_b200.parallel_layers.utils import (cast_if_autocast_enabled, get_padding_length,
                                                                indices_split_along_dim, move_all_tensor_to_cpu,
                                                                verify_casted_dtype)
    from neuronx_distributed_b200.utils.activation_checkpoint import NxDCheckpointWrapper, checkpoint_wrapper
    from neuronx_distributed_b200.utils.batch_utils import shift_labels
    from neuronx_distributed_b200.utils.logger import PackagePathFilter
    from neuronx_distributed_b200.utils.medusa_utils import (evaluate_posterior, generate_candidates, generate_medusa_buffers,
                                                            update_inference_inputs)
    from neuronx_distributed_b200.utils.model_utils import (analyze_shared_parameters, has_fake_tensors, preserve_shared_weights,
                                                           recursive_filter, retie_shared_weights)
    from neuronx_distributed_b200.utils.safetensors_utils import check_for_duplicate_tensors
    from neuronx_distributed_b200.utils.serialization import compress_to_string, is_instance_namedtuple, uncompress_from_string
    from neuronx_distributed_b200.utils.timeline import DistributedTimeline, Event

    # Medusa: tree of 6 nodes over 3 heads × top-3; the model "confirms" the path 7 → 100 → 103 and rejects 106
    b = generate_medusa_buffers([(0,), (1,), (0, 0), (0, 1), (1, 0), (0, 0, 0)], topk=3)
    H, K = 3, 3
    cart, tree = generate_candidates(torch.arange(100, 100 + H * K).view(H, 1, 1, K), torch.tensor([[[7, 8]]]),
                                     b["tree_indices"], b["retrieve_indices"])
    assert tree.tolist() == [[7, 100, 101, 103, 104, 103, 106]] and cart[0].tolist() == [7, 100, 103, 106]
    ver = torch.zeros(cart.shape[0], cart.shape[1], 1, dtype=torch.long)
    ver[:, 0, 0], ver[:, 1, 0], ver[:, 2, 0] = 100, 103, 999
    best, n_acc = evaluate_posterior(ver, cart)
    assert int(best) == 0 and int(n_acc) == 2
    ids, lg, mlg, new_tok, sel = update_inference_inputs(torch.tensor([[1, 2, 3]]), cart, best, n_acc, b["retrieve_indices"], None,
                                                        ver, torch.zeros(H, cart.shape[0], cart.shape[1], K, dtype=torch.long), 0)
    assert ids.tolist() == [[1, 2, 3, 7, 100, 103]] and new_tok == 3 and sel.tolist() == [3, 4, 6]
    assert lg.shape == (1, 1, 1) and int(lg) == 999 and mlg.shape == (H, 1, 1, K)
    best0, n0 = evaluate_posterior(torch.full_like(ver, 5), cart)                      # nothing matches → path 0, length 0
    assert int(best0) == 0 and int(n0) == 0

    P = namedtuple("P", "a b")
    assert uncompress_from_string(compress_to_string({"x": (1, [2]), "t": torch.ones(2)}))["x"] == (1, [2])
    from neuronx_distributed_b200.utils.serialization import SerializationManager, TensorMeta

    sm = SerializationManager()
    skel, tens = sm.serialize({"a": torch.ones(2, 3, requires_grad=True), "b": [torch.zeros(4, dtype=torch.int64), 5]})
    metas = sm.tensor_metas(tens)
    assert metas == [TensorMeta(0, torch.float32, torch.Size([2, 3]), True, torch.device("cpu")),
                     TensorMeta(1, torch.int64, torch.Size([4]), False, torch.device("cpu"))]
    back = sm.deserialize(skel, [torch.empty(m.shape, dtype=m.dtype, device=m.device) for m in metas])   # receiver side
    assert back["a"].shape == (2, 3) and back["b"][1] == 5
    from neuronx_distributed_b200.trace.nxd_model.nxd_model import JITWrapper

    skel3, tens3, metas3 = sm.serialize({"t": torch.ones(2)}, return_stub_list=True)                   # the reference's 3-tuple
    assert len(tens3) == 1 and metas3[0].shape == (2,) and sm.deserialize(stubbed_obj=skel3, tensors=tens3)["t"].shape == (2,)
    # sampler built from a config object (reference constructor): inverse-CDF top-k sampling, greedy for top_k == 1
    from types import SimpleNamespace as NS

    from neuronx_distributed_b200.utils.sampling import Sampler

    logits = torch.tensor([[0.0, 5.0, 1.0, 4.9], [3.0, 0.0, 0.0, 0.0]])
    greedy = Sampler(NS(on_device_sampling=True, hf_config=NS(do_sample=True, num_beams=1, top_k=1)))
    assert greedy.sample(token_logits=logits).tolist() == [1, 0] and greedy.on_device_sampling
    top2 = Sampler(NS(on_device_sampling=False, hf_config=NS(do_sample=True, num_beams=1, top_k=2)))
    draws = {int(top2.sample(logits, generator=torch.Generator().manual_seed(i))[0]) for i in range(40)}
    assert draws == {1, 3}                                                      # only the two best candidates are ever drawn
    with pytest.raises(Exception, match="not supported"):
        Sampler(NS(on_device_sampling=True, hf_config=NS(do_sample=False, num_beams=1, top_k=1)))
    # hook registry with the reference's argument names and return values
    from neuronx_distributed_b200.trainer.post_partition_hooks import PostPartitionHooks

    hk = PostPartitionHooks()

    def filter_to_local_parameter_group(groups, model=None):
        return (groups, type(model).__name__)

    hk.register_post_partition_hook(callable_function=lambda a, b=0: a + b, func_args=(1,), func_kwargs={"b": 2})
    hk.register_post_partition_hook(filter_to_local_parameter_group, (["g"],))
    assert [h["name"] for h in hk.hooks] == ["<lambda>", "filter_to_local_parameter_group"]
    assert hk.execute_all_hooks(model=torch.nn.Linear(1, 1)) == [3, (["g"], "Linear")] and hk.hooks == []
    with pytest.raises(ValueError):
        hk.register_post_partition_hook("not callable")
    jw = JITWrapper(lambda xs: [x * 2 for x in xs])
    assert isinstance(jw, torch.nn.Module) and jw([torch.ones(1)])[0].item() == 2
    assert is_instance_namedtuple(P(1, 2)) and not is_instance_namedtuple((1, 2))
    sb = shift_labels({"input_ids": torch.arange(6).view(2, 3), "labels": torch.arange(6).view(2, 3)})
    assert sb["labels"].tolist() == [[1, 2, -100], [4, 5, -100]] and sb["input_ids"].tolist() == [[0, 1, 2], [3, 4, 5]]

    calls = []
    lin = nn.Linear(4, 4)
    w = checkpoint_wrapper(lin, checkpoint_fn=lambda mod, *a, **k: (calls.append(1), mod(*a, **k))[1])
    assert isinstance(w, NxDCheckpointWrapper) and set(w.state_dict()) == {"weight", "bias"}
    w(torch.randn(2, 4, requires_grad=True)).sum().backward()
    assert calls == [1] and lin.weight.grad is not None

    t = torch.randn(4, 4)
    ck = {"a": t, "b": t, "c": torch.randn(2)}
    assert set(check_for_duplicate_tensors(dict(ck))) == {"a", "b", "c"}               # default: warn only
    assert set(check_for_duplicate_tensors(dict(ck), remove_duplicate_tensors=True)) == {"a", "c"}
    with pytest.raises(RuntimeError):
        check_for_duplicate_tensors({"x": t[:2], "y": t[2:]}, remove_duplicate_tensors=True)

    path = tmp_path / "trace.json"
    tl = DistributedTimeline(str(path))
    tl.mark_event_start("fwd"); tl.mark_event_end("fwd"); tl.mark_step_end()
    tl.mark_event_start("bwd"); tl.mark_event_end("bwd"); tl.mark_step_end()
    events = json.loads(path.read_text().rstrip().rstrip(",") + "]")
    assert [(e["name"], e["ph"]) for e in events] == [("fwd", "B"), ("fwd", "E"), ("bwd", "B"), ("bwd", "E")] and tl.step == 2
    assert Event("x", 0).start == -1
    off = DistributedTimeline(None)
    off.mark_event_start("x"); off.mark_step_end()                                      # disabled: no-ops

    rec = logging.LogRecord("n", logging.INFO, __file__, 1, "m", None, None)
    assert PackagePathFilter().filter(rec) and not os.path.isabs(rec.relativepath)

    assert get_padding_length(30, 8) == 2 and get_padding_length(32, 8) == 0
    assert indices_split_along_dim(torch.zeros(8, 3), 0, 2, 4).tolist() == [4, 5]
    assert indices_split_along_dim(None, 0, 0, 2) is None
    x32 = torch.ones(2)
    assert cast_if_autocast_enabled(x32)[0].dtype == torch.float32
    with torch.autocast("cpu", dtype=torch.bfloat16):
        a, (bb, c), d = cast_if_autocast_enabled(x32, (x32.double(), torch.ones(1, dtype=torch.long)), "s")
        assert a.dtype == torch.bfloat16 and bb.dtype == torch.float64 and c.dtype == torch.long and d == "s"
        verify_casted_dtype((a, {"k": a}))
        with pytest.raises(AssertionError):
            verify_casted_dtype(x32)
    assert move_all_tensor_to_cpu({"a": [x32]})["a"][0] is x32

    class Tied(nn.Module):
        def __init__(self):
            super().__init__()
            self.e, self.h = nn.Embedding(10, 4), nn.Linear(4, 10, bias=False)
            self.h.weight = self.e.weight

    m = Tied()
    assert analyze_shared_parameters(m) == [["e.weight", "h.weight"]]
    with preserve_shared_weights(m):
        m.h.weight = nn.Parameter(torch.zeros(10, 4))
    assert m.h.weight is m.e.weight
    m.h.weight = nn.Parameter(torch.zeros(10, 4))
    retie_shared_weights(m, [["e.weight", "h.weight"]])
    assert m.h.weight is m.e.weight
    assert has_fake_tensors(nn.Linear(2, 2, device="meta")) and not has_fake_tensors(m)
    assert recursive_filter({"a": x32, "b": [torch.zeros(1, device="meta"), 3]}, lambda t_: not t_.is_meta) == {"a": x32, "b": [3]}


def test_nvls_protocol_model_catches_epoch_reset_and_passes_monotonic():
    """tools/sim_nvls_protocol.py: random rank interleavings of AG / RS call sequences with a region re-layout in the middle.
    Monotonic epochs never produce a stale read or a premature overwrite; the round-2 bug (epochs reset while the reused region
    keeps old flags) is caught."""
    import importlib.util
    import os

    import pytest

    spec = importlib.util.spec_from_file_location("sim_nvls_protocol", os.path.join(os.path.dirname(__file__), "..", "tools", "sim_nvls_protocol.py"))
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    calls = ["ag", "rs", "ag", "ag", "rs", "ag", "rs", "rs"]
    for seed in range(25):
        for reuse, reset in ((True, False), (False, False), (False, True)):
            sim.simulate(3, calls, 2, 4, reuse, reset, seed)
    with pytest.raises(sim.StaleRead):
        for seed in range(10):
            sim.simulate(3, calls, 2, 4, True, True, seed)


def test_publish_pull_protocol_model():
    """Model of ``nvls_publish_kernel`` + pull readers (embedding gather, all-to-all, ``publish`` views): alternating halves are
    sufficient for reads that finish before the rank's next publish; a single buffer is not; views read after the rank has arrived
    at its next publish are not safe either (hence ``pull_attention`` re-publishes K/V in backward)."""
    import importlib.util
    import os

    import pytest

    spec = importlib.util.spec_from_file_location("sim_nvls_protocol", os.path.join(os.path.dirname(__file__), "..", "tools", "sim_nvls_protocol.py"))
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    for seed in range(40):
        sim.simulate_publish(2 + seed % 3, 2 + seed % 5, 1 + seed % 3, seed)
    with pytest.raises(sim.StaleRead):
        for seed in range(10):
            sim.simulate_publish(3, 5, 2, seed, single_buffer=True)
    with pytest.raises(sim.StaleRead):
        for seed in range(40):
            sim.simulate_publish(3, 6, 2, seed, reads_outlive=1)
