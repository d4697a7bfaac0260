"""Decode MoE block: the kernel's numerics oracle (``ops.moe_tkg.moe_block_tkg_reference``) against the composed module path
for the routing / activation variants the one-launch kernel implements, and the ``MoEFusedTKG`` wrapper contract."""
import torch

from dist_utils import run_distributed


def _build(hidden_act="silu", glu_type="glu", router_act="softmax", over_topk=False, normalize=True, pre_scale=False, clamps=False,
           E=8, k=2, H=32, inter=48, dtype=torch.float32):
    from neuronx_distributed_b200.modules.moe.expert_mlps_v2 import ExpertMLPsV2
    from neuronx_distributed_b200.modules.moe.moe_configs import RoutedExpertsMLPOpsConfig
    from neuronx_distributed_b200.modules.moe.routing import RouterTopK
    from neuronx_distributed_b200.modules.rms_norm import RMSNorm

    kw = {}
    if clamps:
        kw = dict(gate_clamp_upper_limit=0.8, gate_clamp_lower_limit=-0.7, up_clamp_upper_limit=0.9, up_clamp_lower_limit=-0.6)
    cfg = RoutedExpertsMLPOpsConfig(num_experts=E, top_k=k, hidden_size=H, intermediate_size=inter, hidden_act=hidden_act,
                                    glu_mlp=True, glu_type=glu_type, capacity_factor=None, normalize_top_k_affinities=normalize,
                                    early_expert_affinity_modulation=pre_scale, hidden_act_scaling_factor=1.702 if glu_type == "swiglu" else 1.0,
                                    hidden_act_bias=1.0 if glu_type == "swiglu" else 0.0, **kw)
    torch.manual_seed(0)
    router = RouterTopK(E, k, H, dtype=dtype, act_fn=router_act, apply_act_fn_over_topk=over_topk, bias=True)
    with torch.no_grad():
        router.linear_router.bias.normal_(std=0.1)
    experts = ExpertMLPsV2(cfg, dtype=dtype)
    norm = RMSNorm(H, eps=1e-5)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
    return cfg, router, experts, norm


def _worker(rank, world):
    from neuronx_distributed_b200.modules.moe.moe_configs import MoEFusedTKGConfig
    from neuronx_distributed_b200.modules.moe.moe_fused_tkg import MoEFusedTKG
    from neuronx_distributed_b200.ops import moe_tkg
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    variants = [dict(), dict(router_act="sigmoid"), dict(over_topk=True), dict(over_topk=True, router_act="sigmoid", normalize=False),
                dict(pre_scale=True), dict(glu_type="swiglu", hidden_act="sigmoid", clamps=True), dict(hidden_act="gelu"),
                dict(hidden_act="gelu_new", normalize=False), dict(k=4, E=16)]
    for v in variants:
        cfg, router, experts, norm = _build(**v)
        fused = MoEFusedTKG(router, experts, None, norm, return_router_logits=True, return_expert_index=True).eval()
        x = torch.randn(1, 5, cfg.hidden_size)                                   # [S=1, B=5, H]: five decode tokens
        with torch.no_grad():
            want, want_logits, want_idx = fused(x)                               # composed path (CPU) incl. the TP all-reduce
            op = experts.mlp_op
            c = [cfg.gate_clamp_lower_limit, cfg.gate_clamp_upper_limit, cfg.up_clamp_lower_limit, cfg.up_clamp_upper_limit]
            c = [(-float("inf") if i % 2 == 0 else float("inf")) if val is None else val for i, val in enumerate(c)]
            out, logits, idx, w = moe_tkg.moe_block_tkg(
                x.reshape(-1, cfg.hidden_size), norm.weight, router.linear_router.weight, router.linear_router.bias,
                op.gate_up_proj.weight, op.down_proj.weight, int(op.local_expert_ids[0]), cfg.top_k, norm.variance_epsilon,
                0 if router.act_fn == "softmax" else 1, router.apply_act_fn_over_topk, cfg.normalize_top_k_affinities,
                cfg.early_expert_affinity_modulation, True, moe_tkg.act_id(cfg.hidden_act, cfg.glu_type),
                cfg.hidden_act_scaling_factor, cfg.hidden_act_bias, tuple(c))
            if world > 1:                                                        # the op returns this rank's partial sum
                torch.distributed.all_reduce(out)
        torch.testing.assert_close(logits, want_logits.float(), rtol=1e-5, atol=1e-5, msg=lambda m: f"{v}: {m}")
        assert torch.equal(idx.sort(-1).values, want_idx.sort(-1).values), v
        torch.testing.assert_close(out.view_as(want), want, rtol=2e-4, atol=2e-5, msg=lambda m: f"{v}: {m}")
        assert w.shape == (5, cfg.top_k) and (not cfg.normalize_top_k_affinities or torch.allclose(w.sum(-1), torch.ones(5), atol=1e-5))
    # wrapper contract: reference constructor order, residual stream, CPU → flat path with the reason recorded
    cfg, router, experts, norm = _build()
    f2 = MoEFusedTKG(router, experts, MoEFusedTKGConfig(), 0, None, post_attention_layernorm=norm).eval()
    assert f2.post_attention_layernorm is norm and f2.router is router and f2.expert_mlps is experts and f2.shared_experts is None
    x, res = torch.randn(1, 3, cfg.hidden_size), torch.randn(1, 3, cfg.hidden_size)
    with torch.no_grad():
        y, stream = f2(x, res)
        y0, = f2(x + res)
    torch.testing.assert_close(y, y0) and torch.testing.assert_close(stream, x + res)
    assert f2._why_not == "cannot run on cpu" and not list(f2.parameters())      # holds references, owns nothing
    f2.config.moe_fused_kernel_enabled = False
    assert not f2._can_use_kernel(x) and "disabled" in f2._why_not


def _eligibility(rank, world):
    """Every rule of ``_can_use_kernel`` after the device check, walked on the host (device predicate and the op-level shape /
    dtype check patched): a supported block says yes, each unsupported variant names its reason."""
    from neuronx_distributed_b200 import ops
    from neuronx_distributed_b200.modules.moe import moe_fused_tkg as mod
    from neuronx_distributed_b200.modules.moe.moe_configs import MoEFusedTKGConfig
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    mod._on_cuda = lambda t: True
    seen = []
    ops.moe_tkg.kernel_eligible = lambda x, rw, wgu, wdn, k: (seen.append((tuple(x.shape), tuple(rw.shape), tuple(wgu.shape),
                                                                             tuple(wdn.shape), k)) or True)
    cfg, router, experts, norm = _build()
    f = mod.MoEFusedTKG(router, experts, MoEFusedTKGConfig(), 0, None, post_attention_layernorm=norm).eval()
    x = torch.randn(1, 3, cfg.hidden_size)
    with torch.no_grad():
        assert not f._can_use_kernel(x) and f._why_not == "norm weight dtype"                      # fp32 norm weight on this host build
        norm.weight.data = norm.weight.data.bfloat16()
        assert f._can_use_kernel(x), f._why_not
    (xs, rws, gus, dns, k), = seen[-1:]
    assert xs == (3, cfg.hidden_size) and rws == (cfg.num_experts, cfg.hidden_size) and k == cfg.top_k
    assert gus == (cfg.num_experts, cfg.hidden_size, 2 * dns[1]) and dns == (cfg.num_experts, cfg.intermediate_size, cfg.hidden_size)
    router.act_fn = "tanh"
    assert not f._can_use_kernel(x) and "router" in f._why_not
    router.act_fn = "softmax"
    f.train()
    assert not f._can_use_kernel(x) and "training" in f._why_not


def test_kernel_eligibility_rules_walked_on_host():
    run_distributed(_eligibility, 1)


def test_moe_block_tkg_reference_matches_composed_path_tp1():
    run_distributed(_worker, 1)


def test_moe_block_tkg_reference_matches_composed_path_tp2():
    run_distributed(_worker, 2)


def _serving(rank, world):
    """Mixtral behind the serving wrapper: token generation through the fused decode block (forced on; on CPU it runs the
    kernel's fp32 oracle) gives the same tokens as the MoE layer's own dispatch, TP=2."""
    from neuronx_distributed_b200.models.llama_inference import LlamaForInference
    from neuronx_distributed_b200.models.mixtral import MixtralConfig, MixtralForCausalLM
    from neuronx_distributed_b200.modules.moe.moe_fused_tkg import MoEFusedTKG
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    cfg = MixtralConfig(vocab_size=64, hidden_size=32, intermediate_size=48, num_hidden_layers=2, num_attention_heads=4,
                        num_key_value_heads=2, num_local_experts=4, num_experts_per_tok=2, dtype=torch.float32,
                        max_position_embeddings=32)
    torch.manual_seed(0)
    m = LlamaForInference(cfg, batch_size=2, max_seq_len=32, lm_cls=MixtralForCausalLM).eval()
    ids = torch.randint(0, 64, (2, 8), generator=torch.Generator().manual_seed(1))
    want = m.generate(ids, 6)
    assert "_tkg_blocks" in m.__dict__ and not any("tkg" in n for n, _ in m.named_modules())      # consulted, not in the module tree
    calls = []
    orig = MoEFusedTKG._moe_fused_tkg_kernel

    def counted(self, x):
        calls.append(tuple(x.shape))
        return orig(self, x)

    MoEFusedTKG._can_use_kernel = lambda self, x: True
    MoEFusedTKG._moe_fused_tkg_kernel = counted
    m.kv.reset()
    got = m.generate(ids, 6)
    assert torch.equal(got, want), (got, want)
    assert len(calls) == 2 * 5 and all(c == (1, 2, 32) for c in calls)                          # 2 MoE layers × 5 decode steps


def test_serving_uses_the_fused_moe_decode_block():
    run_distributed(_serving, 2, timeout=240)


def test_kernel_index_arithmetic_emulation():
    """``tools/emulate_moe_tkg.py`` replays the kernel's work decomposition and pointer arithmetic on the host (ragged sizes, an
    expert-parallel slice, both affinity modes): every weight element of an active expert is read exactly once and the result
    equals the oracle."""
    import importlib.util
    import os

    import numpy as np

    from neuronx_distributed_b200.ops import moe_tkg

    spec = importlib.util.spec_from_file_location("emulate_moe_tkg", os.path.join(os.path.dirname(__file__), "..", "tools", "emulate_moe_tkg.py"))
    em = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(em)
    torch.manual_seed(0)
    for T, H, E, I, K, e0, El, pre in ((5, 264, 8, 136, 3, 0, 8, True), (4, 136, 12, 72, 2, 4, 4, False)):
        x, rw = torch.randn(T, H), torch.randn(E, H) * 0.3
        wgu, wdn = torch.randn(El, H, 2 * I) * 0.2, torch.randn(El, I, H) * 0.2
        out, _, idx, w = moe_tkg.moe_block_tkg_reference(x, None, rw, None, wgu, wdn, e0, K, round_logits=False, pre_scale=pre)
        assert int(((idx >= e0) & (idx < e0 + El)).sum()) > 0
        y = em.run(x.numpy(), idx.numpy(), w.numpy(), wgu.numpy(), wdn.numpy(), e0, pre, lambda g, u: g / (1 + np.exp(-g)) * u,
                   grid_warps=7)
        np.testing.assert_allclose(y, out.numpy(), rtol=1e-4, atol=1e-4)
