"""Pipeline engine: schedule invariants (pure data) + 1F1B / interleaved runs on gloo vs non-pipelined."""
import pytest
import torch
from torch import nn

from dist_utils import run_distributed
from neuronx_distributed_b200.pipeline import scheduler as S
from neuronx_distributed_b200.pipeline.partition import create_partitions


def _simulate(schedules):
    """Execute all stages' task lists with rendezvous p2p semantics; returns per-stage compute orders."""
    progs = [[t for step in sch.steps() for t in step] for sch in schedules]
    n = len(progs)
    pc = [0] * n
    done_f, done_b = [set() for _ in range(n)], [set() for _ in range(n)]
    chan = {}  # (src, dst, kind, mb, chunk) -> posted
    guard = 0
    while any(pc[s] < len(progs[s]) for s in range(n)):
        guard += 1
        assert guard < 100000, "deadlock in schedule simulation"
        moved = False
        for s in range(n):
            if pc[s] >= len(progs[s]):
                continue
            t = progs[s][pc[s]]
            if isinstance(t, S.ReduceGradsTask):
                pc[s] += 1; moved = True; continue
            key = (t.mb, t.model_chunk)
            if isinstance(t, S.ForwardStepTask):
                done_f[s].add(key); pc[s] += 1; moved = True
            elif isinstance(t, S.BackwardStepTask):
                assert key in done_f[s]
                done_b[s].add(key); pc[s] += 1; moved = True
            elif isinstance(t, S.ForwardPostprocessTask):
                chan[("f", s, key)] = True; pc[s] += 1; moved = True       # buffered send
            elif isinstance(t, S.BackwardPostprocessTask):
                chan[("b", s, key)] = True; pc[s] += 1; moved = True
            elif isinstance(t, S.ForwardPreprocessTask):
                first = s == 0 and t.model_chunk == 0
                src = (s - 1) % n
                src_key = (t.mb, t.model_chunk if s > 0 else t.model_chunk - 1)
                if first or chan.pop(("f", src, src_key), None):
                    pc[s] += 1; moved = True
            elif isinstance(t, S.BackwardPreprocessTask):
                src = (s + 1) % n
                src_key = (t.mb, t.model_chunk if s < n - 1 else t.model_chunk + 1)
                if chan.pop(("b", src, src_key), None):
                    pc[s] += 1; moved = True
        assert moved, f"deadlock: pcs={pc}"
    return done_f, done_b


@pytest.mark.parametrize("pp,mb", [(2, 1), (2, 4), (4, 4), (4, 8), (8, 32), (16, 32)])
def test_1f1b_schedule_completes(pp, mb):
    schs = [S.Train1F1BSchedule(mb, pp, r) for r in range(pp)]
    f, b = _simulate(schs)
    for r in range(pp):
        assert f[r] == {(m, 0) for m in range(mb)} == b[r]
        order = schs[r].compute_order()
        # never more than (pp - r) forwards in flight
        inflight = mx = 0
        for is_f, *_ in order:
            inflight += 1 if is_f else -1
            mx = max(mx, inflight)
        assert mx <= min(pp - r, mb)
        assert list(schs[r].steps())[-1] == [S.ReduceGradsTask()]


@pytest.mark.parametrize("pp,mb,chunks", [(2, 2, 2), (2, 4, 2), (4, 8, 2), (4, 4, 3), (4, 16, 4)])
def test_interleaved_schedule_completes(pp, mb, chunks):
    schs = [S.TrainInterleavedSchedule(mb, chunks, pp, r) for r in range(pp)]
    f, b = _simulate(schs)
    want = {(m, c) for m in range(mb) for c in range(chunks)}
    for r in range(pp):
        assert f[r] == want == b[r]


def test_interleaved_requires_divisible_microbatches():
    with pytest.raises(ValueError):
        S.TrainInterleavedSchedule(3, 2, 2, 0)


def test_create_partitions():
    assert create_partitions(8, 4) == [1, 3, 5]
    assert create_partitions(10, 4) == [1, 3, 6]      # remainder goes to later stages: sizes 2,2,3,3
    with pytest.raises(ValueError):
        create_partitions(2, 4)
    # the reference's convention: (pipeline_parallel_size, layer names) → names to cut after (what ``pipeline_cuts`` takes)
    names = [f"model.layers.{i}" for i in range(10)]
    assert create_partitions(4, names) == ["model.layers.1", "model.layers.3", "model.layers.6"]
    assert create_partitions(1, names) == []


def test_trace_model_reference_calling_convention():
    """``pipeline.trace.trace_model(model, args=…, kwargs=…, leaf_modules=[class names], autowrap_obj_methods=…)``."""
    from neuronx_distributed_b200.pipeline.trace import trace_model

    class Helper:
        def scale(self, x):
            return x * 2 if x.sum() > 0 else x                        # data-dependent: cannot be traced through

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.blk, self.out, self.h = Block(8), nn.Linear(8, 4), Helper()

        def forward(self, x, mask=None, scale=1.0):
            y = self.blk(x) * scale
            return self.out(self.h.scale(y))

    m = M()
    gm = trace_model(m, args=[torch.zeros(1, 8)], leaf_modules=["Block"], autowrap_obj_methods={m.h: ["scale"]})
    targets = [(n.op, str(n.target)) for n in gm.graph.nodes]
    assert ("call_module", "blk") in targets and not any("fc1" in t for _, t in targets)          # Block kept as one node
    assert [t for op, t in targets if op == "placeholder"] == ["x"]                              # mask / scale are constants
    assert any(op == "call_function" and "scale" in t for op, t in targets)                      # the wrapped method is opaque
    x = torch.randn(3, 8)
    torch.testing.assert_close(gm(x), m(x))
    gm2 = trace_model(m, kwargs={"x": torch.zeros(1, 8), "mask": None}, leaf_modules=[Block], autowrap_obj_methods={m.h: ["scale"]})
    assert [n.target for n in gm2.graph.nodes if n.op == "placeholder"] == ["x"]
    with pytest.raises(ValueError, match="does not name the class"):
        trace_model(m, args=[x], leaf_modules=["NoSuchBlock"])


class Block(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(h, 2 * h), nn.Linear(2 * h, h)

    def forward(self, x):
        return x + self.fc2(torch.tanh(self.fc1(x)))


class Toy(nn.Module):
    def __init__(self, v=32, h=16, layers=4, tie=False):
        super().__init__()
        self.emb = nn.Embedding(v, h)
        self.layers = nn.ModuleList([Block(h) for _ in range(layers)])
        self.head = nn.Linear(h, v, bias=False)
        if tie:
            self.head.weight = self.emb.weight

    def forward(self, input_ids, labels):
        x = self.emb(input_ids)
        for l in self.layers:
            x = l(x)
        logits = self.head(x)
        return torch.nn.functional.cross_entropy(logits.view(-1, logits.shape[-1]), labels.view(-1))


def _pp_worker(rank, world, pp, vpp, tie, out_path):
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.pipeline import NxDPPModel

    ps.initialize_model_parallel(tensor_model_parallel_size=1, pipeline_model_parallel_size=pp)
    torch.manual_seed(0)
    ref = Toy(tie=tie)
    torch.manual_seed(0)
    model = Toy(tie=tie)
    ids = torch.randint(0, 32, (8, 6), generator=torch.Generator().manual_seed(1))
    ref_loss = ref(ids, ids)
    ref_loss.backward()
    ppm = NxDPPModel(model, transformer_layer_cls=Block, num_microbatches=4, virtual_pipeline_size=vpp,
                     output_loss_value_spec=True, input_names=["input_ids", "labels"],
                     broadcast_and_average_loss=True)
    loss = ppm.run_train(input_ids=ids, labels=ids)
    torch.testing.assert_close(loss.float(), ref_loss.detach().float(), rtol=1e-4, atol=1e-5)
    refp = dict(ref.named_parameters(remove_duplicate=False))
    n = 0
    for name, p in ppm.local_named_parameters():
        assert name in refp, name
        if p.grad is not None:
            torch.testing.assert_close(p.grad, refp[name].grad, rtol=1e-3, atol=1e-5)
            n += 1
    assert n > 0
    sd = ppm.local_state_dict()
    assert all(k in dict(ref.state_dict()) for k in sd)
    ev = ppm.run_eval(input_ids=ids, labels=ids)
    assert ev is not None
    # reference-named accessors / translators
    assert ppm.get_current_stage(0) == rank and ppm.is_last_stage() == (rank == pp - 1 and True)
    assert ppm.is_last_pp_rank_last_model_chunk(vpp - 1) == (rank == pp - 1)
    assert len(list(ppm.get_batch_iterator({"input_ids": ids, "labels": ids}))) == 4
    assert type(ppm.create_schedule(train=True)).__name__.startswith("Train")
    local_sd = ppm.local_stage_modules.state_dict(prefix="local_stage_modules.")
    origin = ppm.translate_local_state_dict_to_origin_state_dict(local_sd)
    assert set(origin) == set(sd)
    back = ppm.translate_origin_state_dict_to_local_state_dict({**ref.state_dict()})
    assert set(back) == set(local_sd)
    chunks = ppm.construct_state_dict_per_model_chunk(ref.state_dict(), strict=True)
    assert len(chunks) == vpp and all(set(c) == set(st.module.state_dict()) for c, st in zip(chunks, ppm.stages))
    assert list(ppm.buffers()) == list(ppm.local_buffers()) and [n for n, _ in ppm.local_named_children()] == [str(i) for i in range(vpp)]
    ppm.clear_minibatch_state()
    assert not ppm._act and not ppm._losses
    t = torch.ones(3, requires_grad=True)
    out = (t * 2).sum()
    NxDPPModel.custom_backward(out, None)
    assert t.grad.tolist() == [2, 2, 2]
    # two-step construction: declare cuts, then partition (no transformer_layer_cls)
    torch.manual_seed(0)
    m2 = NxDPPModel(Toy(tie=tie), num_microbatches=4, output_loss_value_spec=True, broadcast_and_average_loss=True)
    assert not m2.partitioned
    m2.trace(kwargs={"input_ids": ids, "labels": ids}, leaf_modules=[Block])
    if pp == 2 and vpp == 1:
        m2.cut_pipeline_stage("layers.1")
        m2.partition()
        torch.testing.assert_close(m2.run_train(input_ids=ids, labels=ids).float(), ref_loss.detach().float(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("pp,vpp,tie", [(2, 1, False), (2, 2, False), (2, 1, True)])
def test_pipeline_matches_single_process(tmp_path, pp, vpp, tie):
    run_distributed(_pp_worker, pp, pp, vpp, tie, str(tmp_path / "o.pt"), timeout=90)


def _ring_worker(rank, world):
    from neuronx_distributed_b200.modules.attention.ring import block_attention, ring_attention
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(1, 1, context_parallel_size=world)
    torch.manual_seed(0)
    B, S, H, D = 2, 16, 4, 8
    q, k, v = [torch.randn(B, S, H, D) for _ in range(3)]
    qr, kr, vr = [t.clone().requires_grad_(True) for t in (q, k, v)]
    ref, _ = block_attention(qr, kr, vr, True, D ** -0.5)
    ref.sum().backward()
    sl = slice(rank * S // world, (rank + 1) * S // world)
    ql, kl, vl = [t[:, sl].clone().requires_grad_(True) for t in (q, k, v)]
    o = ring_attention(ql, kl, vl, True)
    torch.testing.assert_close(o, ref[:, sl].detach(), rtol=1e-4, atol=1e-5)
    o.sum().backward()
    for a, b in ((ql, qr), (kl, kr), (vl, vr)):
        torch.testing.assert_close(a.grad, b.grad[:, sl], rtol=1e-4, atol=1e-5)


def test_ring_attention_context_parallel():
    run_distributed(_ring_worker, 4, timeout=90)


def _delayed(rank, world):
    """Delayed tracing: no ``input_names`` → partitioning waits for the first batch, whose keyword arguments name the traced
    inputs; the result equals the eagerly traced model.  Helpers ``get_delay_tracing`` / ``check_delay_tracing`` and the
    signature analysis of ``get_concrete_args`` follow the reference's rules."""
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.pipeline import NxDPPModel
    from neuronx_distributed_b200.pipeline.trace import get_concrete_args
    from neuronx_distributed_b200.utils.model_utils import check_delay_tracing, get_delay_tracing

    ps.initialize_model_parallel(tensor_model_parallel_size=1, pipeline_model_parallel_size=world)
    ids = torch.randint(0, 32, (8, 6), generator=torch.Generator().manual_seed(1))
    torch.manual_seed(0)
    eager = NxDPPModel(Toy(tie=False), transformer_layer_cls=Block, num_microbatches=4, output_loss_value_spec=True,
                       input_names=["input_ids", "labels"], broadcast_and_average_loss=True)
    torch.manual_seed(0)
    lazy = NxDPPModel(Toy(tie=False), transformer_layer_cls=Block, num_microbatches=4, output_loss_value_spec=True,
                      broadcast_and_average_loss=True, _delay_tracing=True)
    assert eager.partitioned and not lazy.partitioned and get_delay_tracing(lazy) is True and get_delay_tracing(eager) is False
    a = eager.run_train(input_ids=ids, labels=ids)
    b = lazy.run_train(input_ids=ids, labels=ids)                     # traces here: input names = this call's keywords
    assert lazy.partitioned and lazy.input_names == ["input_ids", "labels"] and get_delay_tracing(lazy) is False
    torch.testing.assert_close(a, b)
    for (n1, p1), (n2, p2) in zip(eager.local_named_parameters(), lazy.local_named_parameters()):
        assert n1 == n2
        if p1.grad is not None:
            torch.testing.assert_close(p1.grad, p2.grad)
    # the helper functions
    assert get_delay_tracing({"pipeline_config": {"_delay_tracing": True}}) is True
    assert get_delay_tracing({"pipeline_config": {"other": 1}}) is None and get_delay_tracing("text") is None
    assert check_delay_tracing({"pipeline_config": {"use_model_wrapper": True}})
    assert not check_delay_tracing({"pipeline_config": {"use_model_wrapper": True, "input_names": ["x"]}})
    assert not check_delay_tracing({"pipeline_config": {"use_model_wrapper": False}}) and not check_delay_tracing({"pipeline_config": {}})

    # signature analysis: everything the call does not supply stays a constant (in signature order)
    class Sig(nn.Module):
        def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                    labels=None, use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None,
                    cache_position=None):
            return None

    concrete = get_concrete_args(Sig(), None, (1, 2, "a"), dict(use_cache=True, output_hidden_states=True, return_dict=True))
    assert list(concrete) == ["past_key_values", "inputs_embeds", "labels", "output_attentions", "cache_position"]
    import pytest

    with pytest.raises(ValueError, match="does not have input"):
        get_concrete_args(Sig(), ["nope"])


def test_delayed_tracing_and_signature_analysis_pp2():
    run_distributed(_delayed, 2, timeout=180)


class _Emb(nn.Module):
    def __init__(self, v, h):
        super().__init__()
        self.emb = nn.Embedding(v, h)

    def forward(self, input_ids):
        return self.emb(input_ids)


class _Head(nn.Module):
    def __init__(self, v, h):
        super().__init__()
        self.proj = nn.Linear(h, v, bias=False)

    def forward(self, x):
        return self.proj(x)


def _manual_pp(rank, world, tie_mode):
    """Manually partitioned pipeline (ordered layer list, loss function on the last stage): equals the unpartitioned model, and
    an embedding / head weight shared across the first and the last stage stays shared — by being the same Parameter
    (``"identity"``) or by ``PipelineStageModule.mark_weight_sharing`` on two separate tensors (``"marked"``)."""
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.pipeline import NxDPPModel
    from neuronx_distributed_b200.pipeline.manual_pipe_stage import WEIGHT_SHARING_ATTR_NAME, PipelineStageModule

    ps.initialize_model_parallel(tensor_model_parallel_size=1, pipeline_model_parallel_size=world)
    V, H = 32, 16

    def build():
        torch.manual_seed(0)
        layers = [_Emb(V, H), Block(H), Block(H), _Head(V, H)]
        if tie_mode == "identity":
            layers[-1].proj.weight = layers[0].emb.weight
        elif tie_mode == "marked":
            with torch.no_grad():
                layers[-1].proj.weight.copy_(layers[0].emb.weight)          # separate tensors, equal values
            PipelineStageModule.mark_weight_sharing([(layers[0], "emb.weight"), (layers[-1], "proj.weight")], "embeddings")
        return layers

    def loss_fn(logits, labels):
        return torch.nn.functional.cross_entropy(logits.view(-1, V), labels.view(-1))

    ids = torch.randint(0, V, (8, 6), generator=torch.Generator().manual_seed(1))
    ref_layers = build()
    ref = nn.Sequential(*ref_layers)
    ref_loss = loss_fn(ref(ids), ids)
    ref_loss.backward()
    if tie_mode == "marked":                                               # what sharing means for separate tensors: summed gradients
        g = ref_layers[0].emb.weight.grad + ref_layers[-1].proj.weight.grad
        ref_layers[0].emb.weight.grad, ref_layers[-1].proj.weight.grad = g, g.clone()
    stage = PipelineStageModule(build(), layer_names=["emb", "b0", "b1", "head"])
    if tie_mode != "none":
        assert stage.weight_sharing_groups() == [["0.emb.weight", "3.proj.weight"]]
        assert stage.shared_across_stages(2) == [[(0, "layers.emb.emb.weight"), (1, "layers.head.proj.weight")]]
        if tie_mode == "marked":
            assert getattr(stage.all_layers[0], WEIGHT_SHARING_ATTR_NAME) == {"embeddings": "emb.weight"}
            with pytest.raises(RuntimeError, match="already exists"):
                PipelineStageModule.mark_weight_sharing([(stage.all_layers[0], "emb.weight")], "embeddings")
    else:
        assert stage.weight_sharing_groups() == []
    ppm = NxDPPModel(stage, manual_pp_partition=True, manual_pp_loss_fn=loss_fn, num_microbatches=4, input_names=["input_ids"],
                     broadcast_and_average_loss=True)
    assert len(ppm.shared_weight_groups) == (0 if tie_mode == "none" else 1)
    loss = ppm.run_train(input_ids=ids, labels=ids)
    torch.testing.assert_close(loss.float(), ref_loss.detach().float(), rtol=1e-4, atol=1e-5)
    refp = {n: p for n, p in ref.named_parameters(remove_duplicate=False)}
    checked = 0
    for name, p in ppm.local_named_parameters():                           # names of the user's module: all_layers.<i>.<path>
        for ref_name in ("0.emb.weight", "3.proj.weight"):
            if name.endswith(ref_name):
                torch.testing.assert_close(p.grad, refp[ref_name].grad, rtol=1e-3, atol=1e-5)
                checked += 1
    assert checked == 1                                                    # each rank holds one end of the pipeline


@pytest.mark.parametrize("tie_mode", ["none", "identity", "marked"])
def test_manual_partition_with_shared_weights_pp2(tie_mode):
    run_distributed(_manual_pp, 2, tie_mode, timeout=180)


def _pyobj(rank, world):
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.pipeline import comm

    ps.initialize_model_parallel(tensor_model_parallel_size=1, pipeline_model_parallel_size=world)
    meta = {"from": rank, "shape": (2, 3), "dtype": torch.bfloat16}
    if rank == 0:
        comm.send_python_object(meta, True)                                   # reference flags: send_next=True
        assert comm.recv_python_object(False, method="gloo") == {"from": 1, "shape": (2, 3), "dtype": torch.bfloat16}
    else:
        assert comm.recv_python_object(True)["from"] == 0                     # recv_prev=True
        comm.send_python_object(meta, send_next := False)                     # back to the previous rank
    if rank == 0:
        comm.send_python_object("by rank", 1)                                 # explicit global rank still works
    else:
        assert comm.recv_python_object(0) == "by rank"
    assert comm.MAX_RETRY == 3 and comm.MAX_LENGTH == 2 ** 20
    # blocking tensor helpers of the reference: the receiver allocates from the metadata it was sent first
    t = torch.arange(6, dtype=torch.float32).view(2, 3) + rank
    if rank == 0:
        comm.send_python_object(comm.TensorMeta(0, t.dtype, t.shape, True), True)
        assert comm.send(t, send_next=True) is t
        back = comm.recv_from(comm.TensorMeta(0, torch.float32, torch.Size([2, 3]), False), recv_prev=False)
        assert torch.equal(back, t * 2) and not back.requires_grad
    else:
        meta = comm.recv_python_object(True)
        got = comm.recv_from(meta, recv_prev=True, tracing=False)
        assert torch.equal(got, torch.arange(6, dtype=torch.float32).view(2, 3)) and got.requires_grad
        comm.send(got.detach() * 2, send_next=False)
    # one batched group of isend / irecv per schedule step
    b = comm.P2PBatch(ps.get_pipeline_model_parallel_group())
    peer = 1 - rank
    b.send(torch.full((4,), float(rank)), peer)
    buf = b.recv(comm.TensorMeta(0, torch.float32, torch.Size([4]), False), peer)
    recv_works, send_works = b.launch()
    for w in recv_works + send_works:
        w.wait()
    assert buf.tolist() == [float(peer)] * 4 and b.launch() == ([], [])


def test_python_object_exchange_between_stages():
    run_distributed(_pyobj, 2, timeout=120)


def _manual_tp_pp(rank, world):
    """Manual partition with tensor-parallel layers inside the stages (TP=2 x PP=2): loss and the shared embedding / head
    gradient equal the unpartitioned, unsharded model."""
    from neuronx_distributed_b200.parallel_layers import ColumnParallelLinear, ParallelEmbedding, RowParallelLinear
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.pipeline import NxDPPModel
    from neuronx_distributed_b200.pipeline.manual_pipe_stage import PipelineStageModule

    ps.initialize_model_parallel(tensor_model_parallel_size=2, pipeline_model_parallel_size=2)
    tp_rank = ps.get_tensor_model_parallel_rank()
    V, H = 32, 16

    class TPEmb(nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = ParallelEmbedding(V, H)

        def forward(self, input_ids):
            return self.emb(input_ids)

    class TPBlock(nn.Module):
        def __init__(self):
            super().__init__()
            self.up = ColumnParallelLinear(H, 2 * H, bias=False, gather_output=False, keep_master_weight=True)
            self.down = RowParallelLinear(2 * H, H, bias=False, input_is_parallel=True, keep_master_weight=True)

        def forward(self, x):
            return x + self.down(torch.tanh(self.up(x)))

    class TPHead(nn.Module):
        def __init__(self):
            super().__init__()
            self.proj = ColumnParallelLinear(H, V, bias=False, gather_output=True, keep_master_weight=True)

        def forward(self, x):
            return self.proj(x)

    torch.manual_seed(0)
    layers = [TPEmb(), TPBlock(), TPBlock(), TPHead()]

    def loss_fn(logits, labels):
        return torch.nn.functional.cross_entropy(logits.view(-1, V), labels.view(-1))

    ids = torch.randint(0, V, (8, 6), generator=torch.Generator().manual_seed(1))
    # dense reference from the master weights
    full_emb = torch.randn(V, H, generator=torch.Generator().manual_seed(5)) * 0.5          # same full table on every rank
    with torch.no_grad():
        layers[0].emb.weight.copy_(full_emb.chunk(2, 0)[tp_rank])
    emb_w, head_w = full_emb.clone().requires_grad_(True), layers[3].proj.master_weight.clone().requires_grad_(True)
    blocks = [(l.up.master_weight.clone().requires_grad_(True), l.down.master_weight.clone().requires_grad_(True)) for l in layers[1:3]]
    x = emb_w[ids]
    for up, down in blocks:
        x = x + torch.tanh(x @ up.t()) @ down.t()
    ref_loss = loss_fn(x @ head_w.t(), ids)
    ref_loss.backward()
    stage = PipelineStageModule(layers, layer_names=["emb", "b0", "b1", "head"])
    ppm = NxDPPModel(stage, manual_pp_partition=True, manual_pp_loss_fn=loss_fn, num_microbatches=4, input_names=["input_ids"],
                     broadcast_and_average_loss=True)
    loss = ppm.run_train(input_ids=ids, labels=ids)
    torch.testing.assert_close(loss.float(), ref_loss.detach().float(), rtol=1e-4, atol=1e-5)
    checked = 0
    for name, p in ppm.local_named_parameters():
        if name.endswith("1.up.weight"):
            torch.testing.assert_close(p.grad, blocks[0][0].grad.chunk(2, 0)[tp_rank], rtol=1e-3, atol=1e-5); checked += 1
        if name.endswith("2.down.weight"):
            torch.testing.assert_close(p.grad, blocks[1][1].grad.chunk(2, 1)[tp_rank], rtol=1e-3, atol=1e-5); checked += 1
        if name.endswith("3.proj.weight"):
            torch.testing.assert_close(p.grad, head_w.grad.chunk(2, 0)[tp_rank], rtol=1e-3, atol=1e-5); checked += 1
        if name.endswith("0.emb.weight"):
            torch.testing.assert_close(p.grad, emb_w.grad.chunk(2, 0)[tp_rank], rtol=1e-3, atol=1e-5); checked += 1
    assert checked == 2, checked                                       # two of the four layers live on every pipeline rank


def test_manual_partition_with_tensor_parallel_stages_tp2_pp2():
    run_distributed(_manual_tp_pp, 4, timeout=240)
