"""Operators, sampling, capture/replacement, CP batch split, activation checkpoint wrapper, meta-device init, EP MoE."""
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
    cfg = RoutedExpertsMLPOpsConfig(num_experts=E, top_k=k, hidden_size=H, intermediate_size=I, capacity_factor=4.0)
    layer = MoE(RouterTopK(E, k, H), ExpertMLPsV2(cfg))
    w = layer.expert_mlps.mlp_op.down_proj.weight
    assert w.shape[0] == E // 2 and getattr(w, "expert_model_parallel", False)
    torch.manual_seed(10 + rank)             # different tokens on every DP rank
    x = torch.randn(12, 1, H, requires_grad=True)
    (y,) = layer(x)
    assert y.shape == x.shape
    y.pow(2).mean().backward()
    assert x.grad is not None and w.grad is not None and torch.isfinite(w.grad).all()


def test_expert_parallel_all_to_all_training():
    run_distributed(_ep, 2, timeout=90)
