"""Reference-named entry points added for API parity: attention entry-point layouts, rotary overrides, functional GQA-QKV,
gloo collectives with the xm-style signatures, Lightning accelerator / precision plugin, pipeline tracer helpers, misc."""
import math

import pytest
import torch
from torch import nn

from dist_utils import run_distributed


def test_attention_entry_points_and_overrides():
    from neuronx_distributed_b200.kernels import NKIAttnFunc, get_flash_attn_kernels, get_seq_tile_size, nki_flash_attn_func
    from neuronx_distributed_b200.kernels.kernel_utils import cast, get_seed, permute, torch_to_nki_dtype
    from neuronx_distributed_b200.overrides.transformer_overrides import apply_rotary_pos_emb, rotate_half

    torch.manual_seed(0)
    B, H, S, D = 2, 4, 16, 8
    q, k, v = (torch.randn(B, H, S, D) for _ in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True)
    # transpose_nki_inputs=False: everything [B,H,S,D];  default (True): q,k pre-transposed to [B,H,D,S], v untouched
    torch.testing.assert_close(nki_flash_attn_func(q, k, v, transpose_nki_inputs=False), ref, rtol=1e-5, atol=1e-5)
    qt, kt, _ = permute(q, k, v)
    torch.testing.assert_close(nki_flash_attn_func(qt, kt, v), ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(NKIAttnFunc.apply(qt, kt, v, 1 / math.sqrt(D), False), torch.nn.functional.scaled_dot_product_attention(q, k, v),
                               rtol=1e-5, atol=1e-5)
    qg = q.clone().requires_grad_(True)
    nki_flash_attn_func(qg, k, v, transpose_nki_inputs=False).sum().backward()
    assert qg.grad is not None
    with pytest.raises(AssertionError):
        nki_flash_attn_func(q, k, v, dropout_p=0.1, transpose_nki_inputs=False)
    assert callable(get_flash_attn_kernels()[0])
    assert get_seq_tile_size(8192) == 2048 and get_seq_tile_size(2048) == 1024 and get_seq_tile_size(512) == 1024
    assert get_seed(0.0, "cpu") is None and get_seed(0.1, "cpu").dtype == torch.int32
    assert cast(q, k, v)[0] is q and torch_to_nki_dtype(torch.bfloat16) == torch.bfloat16
    with pytest.raises(ValueError):
        torch_to_nki_dtype(torch.int64)
    # rotary: the [B,H,D,S] flash layout gives the transposed result of the plain layout; == HF formula
    inv = 1.0 / (10000 ** (torch.arange(0, D, 2).float() / D))
    emb = torch.cat([torch.outer(torch.arange(S).float(), inv)] * 2, -1)
    cos, sin = emb.cos()[None].expand(B, -1, -1), emb.sin()[None].expand(B, -1, -1)
    qa, ka = apply_rotary_pos_emb(q, k, cos, sin, None, False)
    qb, kb = apply_rotary_pos_emb(qt, kt, cos, sin, None, True, True)
    torch.testing.assert_close(qb.transpose(-1, -2), qa) and torch.testing.assert_close(kb.transpose(-1, -2), ka)
    x1, x2 = q[..., : D // 2], q[..., D // 2:]
    torch.testing.assert_close(qa, q * cos[:, None] + torch.cat((-x2, x1), -1) * sin[:, None])
    pos = torch.arange(S)[None].expand(B, -1)
    torch.testing.assert_close(apply_rotary_pos_emb(q, k, emb.cos(), emb.sin(), pos, False)[0], qa)
    assert torch.equal(rotate_half(qt, True, True), rotate_half(q).transpose(-1, -2))


def _tp_names(rank, world):
    import torch.distributed as dist

    from neuronx_distributed_b200.modules.qkv_linear import (GQAQKVColumnParallelLinear, GQAQKVLinearWithAsyncCommunication,
                                                             gqa_qkv_linear_with_async_allreduce)
    from neuronx_distributed_b200.modules.qkv_linear_utils import check_requires_grad, check_use_bias
    from neuronx_distributed_b200.operators.topk import get_topk_implementation
    from neuronx_distributed_b200.parallel_layers import comm, mappings
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.parallel_layers.layers import (BaseParallelLinear, ColumnParallelLinear, Conv2dWithInputGradAllReduce,
                                                                 OutputChannelParallelConv2d, RowParallelLinear,
                                                                 conv2d_with_weight_grad_allreduce)
    from neuronx_distributed_b200.parallel_layers.random import XLARNGStatesTracker, get_xla_rng_tracker

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    torch.manual_seed(0)
    # functional GQA-QKV == the module (separate and fused weights), forward and backward
    for fuse in (False, True):
        torch.manual_seed(1)
        layer = GQAQKVColumnParallelLinear(16, [16, 8], bias=True, gather_output=False, kv_size_multiplier=1, fuse_qkv=fuse)
        x = torch.randn(3, 2, 16, requires_grad=True)
        want = layer(x)
        if fuse:
            got = gqa_qkv_linear_with_async_allreduce(x, None, None, None, None, None, None, True, False, 1, layer.weight_qkv,
                                                      layer.bias_qkv, True, layer.q_output_size_per_partition,
                                                      layer.kv_output_size_per_partition)
            assert check_use_bias(layer.weight_qkv, True, None, None, layer.bias_qkv) and check_requires_grad(layer.weight_qkv, True, None)
        else:
            got = GQAQKVLinearWithAsyncCommunication.apply(x, layer.weight_q, layer.weight_k, layer.weight_v, layer.bias_q,
                                                           layer.bias_k, layer.bias_v, True, False)
            assert check_use_bias(None, False, layer.weight_q, layer.bias_q, None)
        for a, b in zip(got, want):
            torch.testing.assert_close(a, b)
        g1 = torch.autograd.grad(sum(t.sum() for t in got), x, retain_graph=True)[0]
        g2 = torch.autograd.grad(sum(t.sum() for t in want), x)[0]
        torch.testing.assert_close(g1, g2)
    # row-parallel slice_indices: using all local columns == plain call; a subset == zeroing the other inputs
    row = RowParallelLinear(8 * world, 6, bias=False, input_is_parallel=True, reduce_output=False)
    assert isinstance(row, BaseParallelLinear) and isinstance(ColumnParallelLinear(4, 4 * world), BaseParallelLinear)
    xi = torch.randn(5, 8)
    torch.testing.assert_close(row(xi, torch.arange(8)), row(xi))
    sub = torch.tensor([1, 4, 6])
    xz = torch.zeros_like(xi); xz[:, sub] = xi[:, sub]
    torch.testing.assert_close(row(xi[:, sub], sub), row(xz))
    # conv functional form == the layer's own autograd function
    conv = OutputChannelParallelConv2d(3, 4 * world, 3, padding=1, bias=True)
    img = torch.randn(2, 3, 6, 6, requires_grad=True)
    y1 = conv2d_with_weight_grad_allreduce(img, conv.weight, conv.bias, (1, 1), (1, 1), True)
    y2 = Conv2dWithInputGradAllReduce.apply(img, conv.weight, conv.bias, (1, 1), (1, 1), (1, 1), 1, True, ps.get_tensor_model_parallel_group())
    torch.testing.assert_close(y1, y2)
    gi = torch.autograd.grad(y1.sum(), img)[0]
    full = [torch.zeros_like(gi) for _ in range(world)]
    dist.all_gather(full, gi)
    assert all(torch.allclose(f, full[0]) for f in full)                      # dgrad was all-reduced over TP
    # gloo collectives with the out-parameter / xm-style signatures
    t = torch.arange(4.0 * world).view(2 * world, 2) + rank
    out = torch.empty(2, 2)
    comm.gloo_reduce_scatter(out, t, "sum")
    want = sum(torch.arange(4.0 * world).view(2 * world, 2) + r for r in range(world))[2 * rank:2 * rank + 2]
    torch.testing.assert_close(out, want)
    torch.testing.assert_close(comm.cpu_reduce_scatter("sum", t, scale=0.5, scatter_dim=0, shard_count=world), want * 0.5)
    swapped = mappings.nonzero_partition_dim_swap(lambda x, dim: x.narrow(dim, 0, 1))
    assert swapped(torch.arange(6).view(2, 3), 1).tolist() == [[0], [3]]
    assert isinstance(get_xla_rng_tracker(), XLARNGStatesTracker)
    uns, srt, stages = get_topk_implementation()
    v = torch.tensor([[1.0, 5.0, 3.0, 4.0]])
    assert srt(v, 2, dim=1).indices.tolist() == [[1, 3]] and sorted(uns(v, 2).values[0].tolist()) == [4.0, 5.0] and stages == 1
    with pytest.raises(AssertionError):
        get_topk_implementation(True, stages=2)


def test_functional_names_tp2():
    run_distributed(_tp_names, 2, timeout=120)


def test_lightning_accelerator_precision_and_misc(tmp_path):
    from neuronx_distributed_b200.lightning import NeuronXLAAccelerator, NeuronXLAPrecisionPlugin
    from neuronx_distributed_b200.lightning.launcher import _NeuronXLALauncher
    from neuronx_distributed_b200.optimizer.zero_dcp_utils import get_dcp_aux_infos
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.parallel_layers.checkpointing import ensure_directory_exists
    from neuronx_distributed_b200.pipeline.model import mark_timeline
    from neuronx_distributed_b200.pipeline.partition import PipelineIO, adding_live_obj_for_previous_stages, iterate_graph_model_outputs
    from neuronx_distributed_b200.pipeline.trace import (NxDTracer, TorchTracerWrapper, get_concrete_args, get_tracer_class,
                                                         patch_obj_method)
    from neuronx_distributed_b200.scripts.yaml_converter import convert_yaml_to_json, load_yaml_file
    from neuronx_distributed_b200.trainer.checkpoint_storage import is_slow_down_error, wait_decrementing_with_jitter
    from neuronx_distributed_b200.trainer.trainer import filter_to_local_parameter_group
    from neuronx_distributed_b200.utils.timeline import DistributedTimeline

    acc = NeuronXLAAccelerator()
    assert acc.parse_devices("0,1") == [0, 1] and acc.parse_devices(3) == 3 and len(acc.get_parallel_devices(2)) == 2
    with pytest.raises(ValueError):
        acc.parse_devices(0)
    assert acc.auto_device_count() == torch.cuda.device_count() and acc.is_available() == torch.cuda.is_available()
    p = nn.Parameter(torch.ones(2))
    opt = torch.optim.SGD([p], lr=0.5)
    p.grad = torch.ones(2)
    plug = NeuronXLAPrecisionPlugin(mixed_precision_enabled=True)
    called = []
    plug.optimizer_step(opt, None, lambda: called.append(1))
    assert called == [1] and p.data.tolist() == [0.5, 0.5]
    with plug.forward_context():
        assert torch.is_autocast_enabled("cpu") or torch.cuda.is_available()
    assert _NeuronXLALauncher().launch(lambda a, b=1: a + b, 1, b=2) == 3

    lin = nn.Linear(2, 2)
    info = get_dcp_aux_infos(lin, torch.optim.Adam(lin.parameters()))
    assert info["optim_pid_to_pnames"] == {0: "weight", 1: "bias"} and info["optim_pid_to_params"][0] is lin.weight

    # optimizer param groups restricted to what is materialised on this (pipeline) rank
    meta_p, real_p = nn.Parameter(torch.empty(2, device="meta")), nn.Parameter(torch.ones(2))
    local = nn.Parameter(torch.zeros(2))
    o2 = torch.optim.SGD([meta_p, real_p], lr=0.1)
    filter_to_local_parameter_group(o2, nn.Module())
    assert o2.param_groups[0]["params"] == [real_p] or (len(o2.param_groups[0]["params"]) == 1 and o2.param_groups[0]["params"][0] is real_p)
    holder = nn.Module()
    holder.meta_device_parameter_map = {meta_p: local}
    o3 = torch.optim.SGD([meta_p, real_p], lr=0.1)
    filter_to_local_parameter_group(o3, holder)
    assert len(o3.param_groups[0]["params"]) == 1 and o3.param_groups[0]["params"][0] is local

    ensure_directory_exists(str(tmp_path / "a" / "b" / "f.pt"))
    assert (tmp_path / "a" / "b").is_dir()
    tl = DistributedTimeline(str(tmp_path / "t.json"))
    with mark_timeline(tl, "ev"):
        pass
    assert tl.current_rank_events["ev"].end > 0
    assert is_slow_down_error(Exception("<Code>SlowDown</Code>")) and not is_slow_down_error(Exception("AccessDenied"))
    w = wait_decrementing_with_jitter(10)
    assert 1 <= w(1) <= 10 and 1 <= w(5) <= 2 and w(100) == 1.0
    assert ps.PG_Group_Logic.LOGIC2(tp=2, dp=2, pp=1).tp_groups() == [[0, 1], [2, 3]]
    assert ps.get_logic_chosen(1, None, 8) is ps.PG_Group_Logic.LOGIC1

    y = tmp_path / "cfg.yaml"
    y.write_text("model:\n  num_layers: 4\n  num_attention_heads: 8\n  hidden_size: 64\n  num_kv_heads: 2\n  moe:\n    num_experts: 16\n")
    import json
    out = convert_yaml_to_json(str(y), str(tmp_path / "c.json"))
    assert json.load(open(out)) == {"num_hidden_layers": 4, "num_attention_heads": 8, "hidden_size": 64, "num_key_value_heads": 2,
                                    "num_local_experts": 16}
    assert load_yaml_file(str(tmp_path / "missing.yaml")) is None

    class Blk(nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = nn.Linear(4, 4)

        def forward(self, x):
            return torch.relu(self.lin(x))

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = Blk(), Blk()

        def helper(self, x):
            return x * 2

        def forward(self, x, mask=None):
            return self.b(self.helper(self.a(x)))

    net = Net()
    assert get_tracer_class(net) is TorchTracerWrapper and get_tracer_class(net, "torch") is TorchTracerWrapper
    with pytest.raises(ValueError):
        get_tracer_class(net, "nope")
    assert get_concrete_args(net, ["x"]) == {"mask": None} and get_concrete_args(net, None, args=[1]) == {"mask": None}
    with pytest.raises(ValueError):
        get_concrete_args(net, ["nope"])
    tracer = TorchTracerWrapper(leaf_modules=["Blk"], autowrap_modules=(), autowrap_functions=())
    assert isinstance(tracer, NxDTracer)
    with patch_obj_method({net: ["helper"]}):
        graph = tracer.trace(net, concrete_args=get_concrete_args(net, ["x"]))
    ops = [(n.op, str(n.target)) for n in graph.nodes]
    assert ("call_module", "a") in ops and ("call_module", "b") in ops and not any(o == "call_module" and "lin" in t for o, t in ops)
    assert any(o == "call_function" and "helper" in t for o, t in ops)        # the method stayed opaque
    assert "helper" not in vars(net)                                           # …and was restored
    ins, outs = [{"x": 1}, {}, {}], [{"h": 1}, {}, {}]
    adding_live_obj_for_previous_stages(ins, outs, "h", 2)
    assert isinstance(ins[1]["h"], PipelineIO) and isinstance(outs[1]["h"], PipelineIO) and "h" in ins[2]
    with pytest.raises(RuntimeError):
        adding_live_obj_for_previous_stages([{}], [{}], "zzz", 0)
    assert list(iterate_graph_model_outputs(((1, 2),))) == [1, 2] and list(iterate_graph_model_outputs((7,))) == [7]
    assert "input_idx_0" in repr(PipelineIO("n", 0))
