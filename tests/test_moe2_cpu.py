"""MoE: expert-MLP execution modes and their dispatch, padding masks, early affinity modulation, expert bias, selective
loading, explicit-metadata blockwise entry points (+ the PyTorch training oracle), MXFP4 decode block, hybrid process groups,
config validation."""
import json
import types

import pytest
import torch

from dist_utils import run_distributed


def _dense_ref(x, a_masked, experts, early=False):
    """out[t] = Σ_e a[t,e]·MLP_e(x[t]) with the experts' (possibly biased) weights; single rank."""
    w1, w2 = experts.gate_up_proj.weight, experts.down_proj.weight
    b1, b2 = experts.gate_up_proj.bias, experts.down_proj.bias
    out = torch.zeros_like(x)
    for e in range(w1.shape[0]):
        xin = x * a_masked[:, e:e + 1] if early else x
        h = xin @ w1[e] + (0 if b1 is None else b1[e])
        y = experts.activation(h) @ w2[e] + (0 if b2 is None else b2[e])
        out += y * ((a_masked[:, e:e + 1] > 0).float() if early else a_masked[:, e:e + 1])
    return out


def _modes(rank, world):
    from neuronx_distributed_b200.modules.moe import ExpertMLPsV2, RoutedExpertsMLPOpsConfig, RouterTopK
    from neuronx_distributed_b200.modules.moe.blockwise import (BlockwiseMatmulArgs, BlockwiseMatmulNKIFunc,
                                                                TorchBlockwiseTraining, augment_inputs_for_padded_blockwise_matmul,
                                                                blockwise_matmul, build_block_metadata,
                                                                can_use_blockwise_matmul_nki, check_blockwise_mm_kernel_compatibility,
                                                                KernelAvailabilityError)
    from neuronx_distributed_b200.modules.moe.expert_mlps_v2 import can_use_find_index_kernel, duplicate_and_replace_prefixes
    from neuronx_distributed_b200.modules.moe.model_utils import ACTFunc, GLUType, get_kernel_activation_func_id
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=1)
    torch.manual_seed(0)
    E, k, H, I, T = 8, 2, 16, 32, 40
    x = torch.randn(T, H)
    router = RouterTopK(E, k, H)
    _, aff, idx = router(x)
    aff = aff.detach()

    for kw in (dict(), dict(bias=True), dict(early_expert_affinity_modulation=True),
               dict(glu_type="swiglu", hidden_act="sigmoid", hidden_act_scaling_factor=1.702, hidden_act_bias=1.0,
                    gate_clamp_upper_limit=7.0, up_clamp_upper_limit=7.0, up_clamp_lower_limit=-7.0)):
        cfg = RoutedExpertsMLPOpsConfig(normalize_top_k_affinities=True, num_experts=E, top_k=k, hidden_size=H, intermediate_size=I, **kw)
        em = ExpertMLPsV2(cfg).eval()
        if cfg.bias:
            with torch.no_grad():
                em.mlp_op.gate_up_proj.bias.normal_(); em.mlp_op.down_proj.bias.normal_()
        mask = em.get_expert_mask(idx, E)
        assert mask.dtype == torch.float64 and mask.sum() == T * k
        am = em.get_expert_affinities_masked(aff, mask, True)
        torch.testing.assert_close(am.sum(-1), torch.ones(T))
        ref = _dense_ref(x, am, em.mlp_op, early=cfg.early_expert_affinity_modulation)
        with torch.no_grad():
            for name, y in (("all", em.forward_all_experts(x, aff, idx)), ("selective", em.forward_selective_loading(x, aff, idx)),
                            ("blockwise", em.forward_blockwise(x, aff, idx)), ("capacity", em.forward_capacity_factor(x, aff, idx)),
                            ("dispatch-prefill", em(x, aff, idx, seq_len=T)), ("dispatch-train", em.train()(x, aff, idx, seq_len=T))):
                torch.testing.assert_close(y, ref, rtol=1e-4, atol=1e-4, msg=lambda m: f"{kw} {name}: {m}")
            em.eval()
            # decode: 2 tokens × top-2 of 8 experts → selective loading; must equal the dense result of those tokens
            torch.testing.assert_close(em(x[:2], aff[:2], idx[:2], seq_len=1), ref[:2], rtol=1e-4, atol=1e-4)
            # padding mask: padded tokens produce zeros, real tokens are unchanged (every mode that takes a mask)
            pm = torch.ones(T); pm[::3] = 0
            for y in (em.forward_all_experts(x, aff, idx, padding_mask=pm), em.forward_blockwise(x, aff, idx, None, pm),
                      em.forward_capacity_factor(x, aff, idx, pm), em(x, aff, idx, seq_len=T, padding_mask=pm.view(1, T))):
                assert y[::3].abs().max() == 0
                torch.testing.assert_close(y[pm.bool()], ref[pm.bool()], rtol=1e-4, atol=1e-4)
            m2, a2 = em.mask_padding_tokens(mask, am, pm.view(2, T // 2))
            assert m2[::3].sum() == 0 and a2[::3].sum() == 0 and em.mask_padding_tokens(mask, am, None)[0] is mask

    # capacity factor: dropping really drops (tokens beyond capacity get no contribution from that expert)
    cfg = RoutedExpertsMLPOpsConfig(normalize_top_k_affinities=True, num_experts=E, top_k=k, hidden_size=H, intermediate_size=I, capacity_factor=0.5)
    em = ExpertMLPsV2(cfg).eval()
    y = em(x, aff, idx, seq_len=T)
    assert (y.abs().sum(-1) == 0).any() or not torch.allclose(y, _dense_ref(x, em._topk_affinities(aff, idx), em.mlp_op), atol=1e-4)
    cfg_full = RoutedExpertsMLPOpsConfig(normalize_top_k_affinities=True, num_experts=E, top_k=k, hidden_size=H, intermediate_size=I, capacity_factor=100.0)
    ExpertMLPsV2.validate_routed_experts_configs(cfg_full)
    assert cfg_full.capacity_factor is None                                  # ≥ E/k cannot drop → full capacity
    with pytest.raises(ValueError):
        ExpertMLPsV2.validate_routed_experts_configs(RoutedExpertsMLPOpsConfig(normalize_top_k_affinities=True, num_experts=4, top_k=5))
    with pytest.raises(ValueError):
        ExpertMLPsV2.validate_routed_experts_configs(RoutedExpertsMLPOpsConfig(hidden_act="nope"))

    # explicit-metadata entry points and the PyTorch training oracle (forward AND hand-written backward vs autograd)
    cfg = RoutedExpertsMLPOpsConfig(normalize_top_k_affinities=True, num_experts=E, top_k=k, hidden_size=H, intermediate_size=I)
    em = ExpertMLPsV2(cfg)
    B = 8
    b2e, tp2id, counts = build_block_metadata(idx, E, B)
    b2e_m, tp2id_m = em.get_blockwise_expert_and_token_mapping(T, b2e.numel(), None, idx, block_size=B)
    assert torch.equal(b2e_m, b2e) and torch.equal(tp2id_m, tp2id)
    am = em._topk_affinities(aff, idx)
    w1 = em.mlp_op.gate_up_proj.weight.detach().clone().requires_grad_(True)
    w2 = em.mlp_op.down_proj.weight.detach().clone().requires_grad_(True)
    xa, aa = x.clone().requires_grad_(True), am.clone().requires_grad_(True)
    y_or = TorchBlockwiseTraining.apply(xa, aa, tp2id, b2e, w1, w2)
    y_or.pow(2).sum().backward()
    w1b, w2b = w1.detach().clone().requires_grad_(True), w2.detach().clone().requires_grad_(True)
    xb, ab = x.clone().requires_grad_(True), am.clone().requires_grad_(True)
    args = BlockwiseMatmulArgs(xb, ab, w1b, w2b, tp2id, b2e, B)
    y_fn = blockwise_matmul(args)
    y_fn.pow(2).sum().backward()
    torch.testing.assert_close(y_fn, y_or, rtol=1e-4, atol=1e-4)
    for got, want in ((xb.grad, xa.grad), (ab.grad, aa.grad), (w1b.grad, w1.grad), (w2b.grad, w2.grad)):
        torch.testing.assert_close(got, want, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(BlockwiseMatmulNKIFunc.apply(x, am, w1, w2, tp2id, b2e, B), y_or, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(em.torch_blockwise_matmul_inference(x, am, idx), y_or, rtol=1e-4, atol=1e-4)
    # kernel-side enums of the reference's entry points (re-exported by blockwise as there) and the kernel resolver
    from neuronx_distributed_b200.modules.moe import blockwise as bw
    from neuronx_distributed_b200.modules.moe.nki_import import NKIImport, import_nki, import_nki_beta2

    post = blockwise_matmul(BlockwiseMatmulArgs(x, am, w1, w2, tp2id, b2e, B, expert_affinities_scaling_mode=bw.ExpertAffinityScaleMode.POST_SCALE,
                                                skip_dma=bw.SkipMode(True, False), block_sharding_strategy=bw.BlockShardStrategy.PING_PONG,
                                                kernel_act_fn=bw.ActFnType.SiLU))
    torch.testing.assert_close(post, y_or, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(blockwise_matmul(BlockwiseMatmulArgs(x, am, w1, w2, tp2id, b2e, B, expert_affinities_scaling_mode=1)), post)
    none = blockwise_matmul(BlockwiseMatmulArgs(x, am, w1, w2, tp2id, b2e, B, expert_affinities_scaling_mode=bw.ExpertAffinityScaleMode.NO_SCALE))
    ones = blockwise_matmul(BlockwiseMatmulArgs(x, (am != 0).to(am.dtype), w1, w2, tp2id, b2e, B))
    torch.testing.assert_close(none, ones)                                  # NO_SCALE = plain sum over the chosen experts
    pre = blockwise_matmul(BlockwiseMatmulArgs(x, am, w1, w2, tp2id, b2e, B, expert_affinities_scaling_mode="pre_scale"))
    assert not torch.allclose(pre, post) and bw.ExpertAffinityScaleMode.coerce(2) is bw.ExpertAffinityScaleMode.PRE_SCALE
    assert bw.ActivationFunction is bw.ActFnType and bw.torch_to_nki_dtype(torch.bfloat16) == "bf16"
    with pytest.raises(ValueError):
        bw.torch_to_nki_dtype(torch.complex64)
    fn, err = import_nki(NKIImport("blockwise_mlp_from_metadata", module_name="modules.moe.blockwise"))
    assert fn is bw.blockwise_mlp_from_metadata and err is None
    fn, err = import_nki_beta2(NKIImport("no_such_kernel", module_name="moe.moe_cte.bwmm_shard_on_block", is_kernel=False))
    assert fn is None and "no_such_kernel" in err
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                                     # device entry points warn when the extension is not built
        comps, train = bw.initialize_nki_components(), bw.initialize_training_kernels()
    assert comps["moe_cte"] is bw.blockwise_expert_mlp and comps["affinity_scale_mode"] is bw.ExpertAffinityScaleMode
    assert train["blockwise_mm_training"] is TorchBlockwiseTraining and set(train) >= {"blockwise_mm_bwd"}
    o, h, ids, a = augment_inputs_for_padded_blockwise_matmul(torch.zeros(T, H), x, tp2id, am)
    assert o.shape == (T + 1, H) and h[-1].abs().sum() == 0 and ids.min() >= 0 and (ids == T).sum() == (tp2id < 0).sum() and a.shape == (T + 1, E)
    assert not can_use_blockwise_matmul_nki(H, I, 128, device=torch.device("cpu"))
    check_blockwise_mm_kernel_compatibility(1024, 128, 512)
    with pytest.raises(KernelAvailabilityError):
        check_blockwise_mm_kernel_compatibility(1024, 100, 512)
    assert can_use_find_index_kernel(128, 64, 2) and not can_use_find_index_kernel(0, 64, 2)
    sd = {"a.mlp_op.w": 1, "b.other": 2}
    duplicate_and_replace_prefixes("mlp_op.", "mlp_op_tkg.", sd)
    assert sd["a.mlp_op_tkg.w"] == 1 and len(sd) == 3
    assert GLUType.validate(None) is GLUType.GLU and GLUType.validate("swiglu") is GLUType.SWIGLU
    assert ACTFunc.validate("gelu") is ACTFunc.GELU and ACTFunc.SILU.value == 0 and ACTFunc.validate(None) is ACTFunc.SIGMOID
    assert get_kernel_activation_func_id(ACTFunc.SILU, GLUType.GLU) == 0 and get_kernel_activation_func_id(ACTFunc.SIGMOID, GLUType.SWIGLU) == 3
    with pytest.raises(ValueError):
        get_kernel_activation_func_id(ACTFunc.GELU, GLUType.GLU)
    with pytest.raises(ValueError):
        GLUType.validate("geglu")


def test_expert_mlp_modes_and_blockwise_entry_points():
    run_distributed(_modes, 1, timeout=150)


def _mx_and_groups(rank, world):
    import torch.distributed as dist

    from neuronx_distributed_b200.modules.moe import ExpertMLPsV2, MoE, RoutedExpertsMLPOpsConfig, RouterTopK, SharedExperts
    from neuronx_distributed_b200.modules.moe import moe_process_group as mpg
    from neuronx_distributed_b200.modules.moe.moe_fused_tkg import MoEFusedTKG
    from neuronx_distributed_b200.modules.moe.moe_fused_tkg_mx import (MoEFusedTKGMX, mxfp4_moe_block_tkg_wrapper,
                                                                       pack_expert_weight_mxfp4)
    from neuronx_distributed_b200.modules.moe.token_shuffling import all_to_all_for_shuffle
    from neuronx_distributed_b200.modules.rms_norm import RMSNorm
    from neuronx_distributed_b200.parallel_layers import mappings
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    torch.manual_seed(0)
    E, k, H, I, T = 4, 2, 64, 128, 3
    cfg = RoutedExpertsMLPOpsConfig(normalize_top_k_affinities=True, num_experts=E, top_k=k, hidden_size=H, intermediate_size=I)
    router, experts, shared, norm = RouterTopK(E, k, H), ExpertMLPsV2(cfg), SharedExperts(H, 32), RMSNorm(H)
    x = torch.randn(T, 1, H, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = MoEFusedTKG(router, experts, shared, norm).eval()(x)[0]
        mx = MoEFusedTKGMX(router, experts, shared, norm).eval()
        assert mx.gate_up_x4.dtype == torch.uint16 and mx.gate_up_x4.shape == (E, 2 * I // world, H // 4)
        assert mx.down_scale.dtype == torch.uint8 and mx.down_scale.shape == (E, H, I // world // 32)
        got = mx(x)[0]
        err = float((got - want).abs().max() / want.abs().max())
        assert err < 0.25, err                                              # fp4 weights: a few % typical, 25 % worst element
        y2, res = mx(x, residual=torch.ones_like(x))
        torch.testing.assert_close(res, x + 1)
        assert {"gate_up_x4", "gate_up_scale", "down_x4", "down_scale"} <= set(mx.state_dict())
        if world == 1:
            # functional form == module (same packed weights, softmax router, normalised top-k, no shared experts)
            mx_ns = MoEFusedTKGMX(router, experts, None, norm).eval()
            out, logits = mxfp4_moe_block_tkg_wrapper(
                x.reshape(T, H), norm.weight, router.linear_router.weight, mx.gate_up_x4, mx.down_x4,
                expert_gate_up_weights_scale=mx.gate_up_scale, expert_down_weights_scale=mx.down_scale, eps=norm.variance_epsilon, top_k=k,
                router_act_fn="softmax", norm_topk_prob=True)
            torch.testing.assert_close(out, mx_ns(x)[0].reshape(T, H), rtol=2e-3, atol=2e-3)
            assert logits.shape == (T, E)
            out_all, _ = mxfp4_moe_block_tkg_wrapper(
                x.reshape(T, H), norm.weight, router.linear_router.weight, mx.gate_up_x4, mx.down_x4,
                expert_gate_up_weights_scale=mx.gate_up_scale, expert_down_weights_scale=mx.down_scale, eps=norm.variance_epsilon, top_k=k,
                router_act_fn="softmax", norm_topk_prob=True, is_all_expert=True)
            torch.testing.assert_close(out_all, out, rtol=1e-3, atol=1e-3)
    w = torch.randn(2, 64, 8)
    p, s = pack_expert_weight_mxfp4(w)
    assert p.shape == (2, 8, 16) and s.shape == (2, 8, 2)

    # SP shared experts inside the MoE layer: prefill runs them on the local sequence shard (no collective)
    if world > 1:
        torch.manual_seed(3)
        sh_sp = SharedExperts(H, 32, sequence_parallel_enabled=True, fused_gate_up_projection=True).eval()
        sh_tp = SharedExperts(H, 32, fused_gate_up_projection=True).eval()
        for n in ("gate_up_proj", "down_proj"):
            dist.broadcast(getattr(sh_sp, n).weight.data, 0)
        full_gu, full_dn = sh_sp.gate_up_proj.weight.data, sh_sp.down_proj.weight.data
        from neuronx_distributed_b200.parallel_layers.utils import create_local_weight
        sh_tp.gate_up_proj.weight.data.copy_(create_local_weight(full_gu, 0, full_gu.shape[0] // world, 2, rank=rank, world_size=world))
        sh_tp.down_proj.weight.data.copy_(create_local_weight(full_dn, 1, full_dn.shape[1] // world, 1, rank=rank, world_size=world))
        r2 = RouterTopK(E, k, H, sequence_parallel_enabled=True)
        xs = torch.randn(8, 2, H, generator=torch.Generator().manual_seed(5)).chunk(world, 0)[rank]
        with torch.no_grad():
            y_sp = MoE(r2, experts, shared_experts=sh_sp, sequence_parallel_enabled=True).eval()(xs)[0]
            y_tp = MoE(r2, experts, shared_experts=sh_tp, sequence_parallel_enabled=True).eval()(xs)[0]
        torch.testing.assert_close(y_sp, y_tp, rtol=1e-4, atol=1e-4)

    # hybrid prefill/decode process groups
    mpg.init_tensor_expert_parallel_moe_process_groups(tkg_tp_degree=world, tkg_ep_degree=1, cte_tp_degree=1, cte_ep_degree=world)
    assert dist.get_world_size(mpg.get_moe_tp_ep_group(prefill=False)) == world
    assert dist.get_world_size(mpg.get_moe_ep_group(prefill=True)) == world
    assert mpg.get_moe_group_ranks(True).tp_ranks == [[r] for r in range(world)]
    mpg.destroy_moe_model_parallel()
    try:
        mpg.get_moe_ep_group()
        raise SystemExit("expected an assertion after destroy")
    except AssertionError:
        pass

    # the shuffle exchange on its own is self-inverse
    ps.initialize_token_shuffle_group(ps.get_data_parallel_size())
    t = torch.arange(8.0).view(4, 2) + 10 * rank
    torch.testing.assert_close(all_to_all_for_shuffle(all_to_all_for_shuffle(t)), t)


def test_mx_decode_block_sp_shared_experts_and_groups_tp2():
    run_distributed(_mx_and_groups, 2, timeout=150)


def test_mx_decode_block_single_rank():
    run_distributed(_mx_and_groups, 1, timeout=150)


def test_moe_config_validator(tmp_path):
    from neuronx_distributed_b200.modules.moe.moe_config_validator import MoeConfigValidator
    from neuronx_distributed_b200.modules.moe.moe_configs import to_torch_dtype

    def make(source, dropless, cf, glu=True, act="silu", hf=None):
        moe = types.SimpleNamespace(dropless=dropless, capacity_factor=cf, glu_mlp=glu)
        path = tmp_path / "config.json"
        path.write_text(json.dumps(hf if hf is not None else {"hidden_act": act}))
        return types.SimpleNamespace(model_source=source, model=types.SimpleNamespace(moe=moe, model_config=str(path), activation=act))

    c = make("hf", True, 2.0)
    MoeConfigValidator(c).validate_moe_config()
    assert c.model.moe.capacity_factor == 0.0                                 # dropless forces capacity 0
    with pytest.raises(ValueError):
        MoeConfigValidator(make("hf", True, 0.0, act="gelu")).validate_moe_config()
    with pytest.raises(ValueError):
        MoeConfigValidator(make("megatron", True, 0.0, act="gelu")).validate_moe_config()
    MoeConfigValidator(make("megatron", True, 0.0, act="swiglu")).validate_moe_config()
    with pytest.raises(ValueError):
        MoeConfigValidator(make("hf", True, 0.0, glu=False)).validate_moe_config()
    with pytest.raises(ValueError):
        MoeConfigValidator(make("hf", False, 0.0)).validate_moe_config()      # dropping needs a positive factor
    MoeConfigValidator(make("hf", False, 1.25)).validate_moe_config()
    MoeConfigValidator(make("hf", True, 0.0, hf={"model_type": "dbrx", "ffn_config": {"ffn_act_fn": {"name": "silu"}}})).validate_moe_config()
    with pytest.raises(ValueError):
        MoeConfigValidator(make("hf", True, 0.0, hf={"model_type": "dbrx", "ffn_config": {"ffn_act_fn": {"name": "gelu"}}})).validate_moe_config()
    with pytest.raises(AttributeError):
        MoeConfigValidator(types.SimpleNamespace(model_source="hf", model=types.SimpleNamespace())).validate_moe_config()
    (tmp_path / "bad.json").write_text("{not json")
    from neuronx_distributed_b200.utils.utils import HloMetadataLevel, get_dict_from_json, hardware
    with pytest.raises(ValueError):
        get_dict_from_json(tmp_path / "bad.json")
    assert hardware("trn2") is hardware.B200 and HloMetadataLevel(False) is HloMetadataLevel.INFO
    assert to_torch_dtype("bfloat16") == torch.bfloat16 and to_torch_dtype(torch.float16) == torch.float16


def _legacy_expert_mlps(rank, world):
    """``ExpertMLPs`` (first-generation flat-keyword constructor, reference ``modules/moe/expert_mlps.py``) builds the same module
    as ``ExpertMLPsV2`` with the two config objects."""
    import pytest

    from neuronx_distributed_b200.modules.moe import ExpertMLPs, ExpertMLPsV2
    from neuronx_distributed_b200.modules.moe.model_utils import GLUType
    from neuronx_distributed_b200.modules.moe.moe_configs import BlockwiseMatmulConfig, RoutedExpertsMLPOpsConfig
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    old = ExpertMLPs(4, 2, 16, 32, "silu", True, None, block_size=None, normalize_top_k_affinities=True, glu_type=GLUType.GLU,
                     init_method=torch.nn.init.kaiming_uniform_, use_torch_block_wise=True, blockwise_nki_autograd_cls=None,
                     early_expert_affinity_modulation=False, use_shard_on_block_dynamic_while=False, dtype=torch.float32)
    new = ExpertMLPsV2(RoutedExpertsMLPOpsConfig(num_experts=4, top_k=2, hidden_size=16, intermediate_size=32, hidden_act="silu",
                                                 glu_mlp=True, capacity_factor=None, normalize_top_k_affinities=True),
                       BlockwiseMatmulConfig(use_torch_block_wise=True), dtype=torch.float32)
    assert isinstance(old, ExpertMLPsV2) and old.cfg.input_layer_init_method is torch.nn.init.kaiming_uniform_
    assert old.bw == new.bw and old.cfg.top_k == 2 and old.bw.block_size == 512
    new.load_state_dict(old.state_dict())
    x = torch.randn(6, 16)
    aff = torch.softmax(torch.randn(6, 4), -1)
    idx = aff.topk(2, -1).indices
    torch.testing.assert_close(old(x, aff, idx, seq_len=6), new(x, aff, idx, seq_len=6))
    with pytest.raises(TypeError):
        ExpertMLPs(4, 2, 16, 32, "silu", True, None, not_an_option=1)


def test_legacy_expert_mlps_constructor():
    run_distributed(_legacy_expert_mlps, 2, timeout=120)
