import os

import torch

from dist_utils import run_distributed


def test_checkpoint_converter_roundtrip(tmp_path):
    from neuronx_distributed_b200.scripts.checkpoint_converter import CheckpointConverterBase, main

    torch.manual_seed(0)
    full = {"model.embed_tokens.weight": torch.randn(16, 8), "lm_head.weight": torch.randn(16, 8), "model.norm.weight": torch.ones(8)}
    for l in range(4):
        p = f"model.layers.{l}."
        full.update({p + "self_attn.q_proj.weight": torch.randn(8, 8), p + "self_attn.k_proj.weight": torch.randn(8, 8),
                     p + "self_attn.v_proj.weight": torch.randn(8, 8), p + "self_attn.o_proj.weight": torch.randn(8, 8),
                     p + "mlp.gate_proj.weight": torch.randn(12, 8), p + "mlp.up_proj.weight": torch.randn(12, 8),
                     p + "mlp.down_proj.weight": torch.randn(8, 12), p + "input_layernorm.weight": torch.ones(8)})
    src = tmp_path / "full.pt"
    torch.save(full, src)
    out = tmp_path / "sharded"
    main(["--input_dir", str(src), "--output_dir", str(out), "--convert_from_full_state", "--tp_size", "2", "--pp_size", "2",
          "--n_layers", "4"])
    assert os.path.isfile(out / "converted" / "model" / "dp_rank_00_tp_rank_01_pp_rank_01.pt")
    sh = torch.load(out / "converted" / "model" / "dp_rank_00_tp_rank_00_pp_rank_00.pt")
    assert "model.layers.0.self_attn.q_proj.weight" in sh and "model.layers.3.mlp.down_proj.weight" not in sh
    assert sh["model.layers.0.self_attn.q_proj.weight"].shape == (4, 8) and sh["model.layers.0.mlp.down_proj.weight"].shape == (8, 6)
    back = tmp_path / "back"
    main(["--input_dir", str(out), "--output_dir", str(back), "--convert_to_full_state", "--tp_size", "2", "--pp_size", "2"])
    rec = torch.load(back / "pytorch_model.bin")
    for k, v in full.items():
        torch.testing.assert_close(rec[k], v)
    # fused gate_up with stride-2 interleave shards consistently
    c = CheckpointConverterBase()
    class A: fuse_gate_up = True; qkv_linear = False; fuse_qkv = False
    shards = c.convert_full_state_to_tp(dict(full), 2, A)
    w = shards[1]["model.layers.0.mlp.gate_up_proj.weight"]
    torch.testing.assert_close(w[:6], full["model.layers.0.mlp.gate_proj.weight"][6:])
    torch.testing.assert_close(w[6:], full["model.layers.0.mlp.up_proj.weight"][6:])


def _zero_ckpt(rank, world, root):
    import neuronx_distributed_b200 as nxd
    from neuronx_distributed_b200.models.llama import LlamaConfig, LlamaForCausalLM
    from neuronx_distributed_b200.utils.adamw_fp32_optim_params import AdamW_FP32OptimParams

    cfg = nxd.neuronx_distributed_config(tensor_parallel_size=1, optimizer_config={"zero_one_enabled": True, "grad_clipping": True, "max_grad_norm": 1.0})
    mcfg = LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=4,
                       dtype=torch.float32, max_position_embeddings=16)
    torch.manual_seed(0)
    model = nxd.initialize_parallel_model(cfg, lambda: LlamaForCausalLM(mcfg))
    opt = nxd.initialize_parallel_optimizer(cfg, AdamW_FP32OptimParams, model.parameters(), lr=1e-2)
    ids = torch.randint(0, 64, (2, 16), generator=torch.Generator().manual_seed(rank))
    opt.zero_grad(); model.run_train(input_ids=ids, labels=ids); opt.step()
    nxd.save_checkpoint(root, "s1", model=model, optimizer=opt)
    nxd.finalize_checkpoint()


def test_convert_zero_checkpoints(tmp_path):
    from neuronx_distributed_b200.optimizer.convert_zero_checkpoints import main

    run_distributed(_zero_ckpt, 2, str(tmp_path), timeout=120)
    optim = tmp_path / "s1" / "optim"
    assert os.path.isfile(optim / "dp_rank_01_tp_rank_00_pp_rank_00.pt")
    main(["--input_dir", str(optim), "--output_dir", str(tmp_path / "full"), "--convert_to_full"])
    full = torch.load(tmp_path / "full" / "full_tp_rank_00_pp_rank_00.pt", weights_only=False)
    assert full["named"]["master"] and full["named"]["exp_avg"]
    main(["--input_dir", str(optim), "--output_dir", str(tmp_path / "re"), "--convert_to_sharded", "--dp_size", "4"])
    parts = [torch.load(tmp_path / "re" / f"dp_rank_{r:02d}_tp_rank_00_pp_rank_00.pt", weights_only=False) for r in range(4)]
    cat = torch.cat([p["sharded_master_weights"][0] for p in parts])
    orig = full["full"]["master"][0]
    torch.testing.assert_close(cat[: orig.numel()], orig)
    # reference CLI conventions: tag directory as input (optim/ inside), sizes inferred from file names, full → sharded
    from types import SimpleNamespace

    from neuronx_distributed_b200.optimizer.convert_zero_checkpoints import (get_parallel_info, is_full, is_xser,
                                                                             merge_optim_dp_checkpoints)
    tag = SimpleNamespace(input_dir=str(tmp_path / "s1"))
    assert get_parallel_info(tag) == (2, 1, 1) and not is_full(tag) and not is_xser(tag)
    merged = merge_optim_dp_checkpoints(tag, 0, 0)
    torch.testing.assert_close(merged["master"][0], orig)
    main(["--input_dir", str(tmp_path / "s1"), "--output_dir", str(tmp_path / "full2"), "--convert_to_full", "--num_workers", "2"])
    assert is_full(SimpleNamespace(input_dir=str(tmp_path / "full2")))
    assert get_parallel_info(SimpleNamespace(input_dir=str(tmp_path / "full2"))) == (0, 1, 1)
    main(["--input_dir", str(tmp_path / "full2"), "--output_dir", str(tmp_path / "re2"), "--convert_to_sharded", "--dp_size", "2"])
    back = [torch.load(tmp_path / "re2" / "optim" / f"dp_rank_{r:02d}_tp_rank_00_pp_rank_00.pt", weights_only=False) for r in range(2)]
    torch.testing.assert_close(torch.cat([p["sharded_master_weights"][0] for p in back])[: orig.numel()], orig)
    import pytest
    with pytest.raises(ValueError):
        main(["--input_dir", str(tmp_path / "full2"), "--output_dir", str(tmp_path / "x"), "--convert_to_full"])


def test_checkpoint_converter_gqa_ep_xser_megatron(tmp_path):
    import json

    import pytest

    from neuronx_distributed_b200.scripts.checkpoint_converter import CheckpointConverterBase, gqa_q_head_permutation, main

    # (1) the Q-head permutation pairs every TP rank's Q heads with the KV heads it holds (both replication layouts)
    assert gqa_q_head_permutation(8, 2, 2, "tile") == [0, 1, 4, 5, 2, 3, 6, 7] and gqa_q_head_permutation(8, 2, 2, "adjacent") == list(range(8))
    QH, KVH, HD, H, L, TP, MULT = 8, 2, 4, 32, 2, 4, 2
    full = {"model.embed_tokens.weight": torch.randn(16, H), "lm_head.weight": torch.randn(16, H), "model.norm.weight": torch.ones(H)}
    for l in range(L):
        p = f"model.layers.{l}."
        full[p + "self_attn.q_proj.weight"] = torch.arange(QH).repeat_interleave(HD).float()[:, None].expand(QH * HD, H).clone()
        full[p + "self_attn.k_proj.weight"] = torch.arange(KVH).repeat_interleave(HD).float()[:, None].expand(KVH * HD, H).clone()
        full[p + "self_attn.v_proj.weight"] = torch.randn(KVH * HD, H)
        full[p + "self_attn.o_proj.weight"] = torch.arange(QH).repeat_interleave(HD).float()[None, :].expand(H, QH * HD).clone()
        full[p + "mlp.gate_proj.weight"], full[p + "mlp.up_proj.weight"] = torch.randn(24, H), torch.randn(24, H)
        full[p + "mlp.down_proj.weight"] = torch.randn(H, 24)
        full[p + "mlp.expert_mlps.w"] = torch.arange(4.0)[:, None, None].expand(4, 2, 2).clone()
    src = tmp_path / "full.pt"
    torch.save(full, src)
    cfg = tmp_path / "config.json"
    cfg.write_text(json.dumps({"num_attention_heads": QH, "num_key_value_heads": KVH, "hidden_size": QH * HD, "num_hidden_layers": L}))
    for layout in ("tile", "adjacent"):
        out = tmp_path / f"sh_{layout}"
        main(["--input_dir", str(src), "--output_dir", str(out), "--convert_from_full_state", "--tp_size", str(TP), "--qkv_linear", "true",
              "--kv_size_multiplier", str(MULT), "--kv_replication_layout", layout, "--config", str(cfg), "--fuse_gate_up", "--ep_size", "2"])
        for r in range(TP):
            sh = torch.load(out / "converted" / "model" / f"dp_rank_00_ep_rank_00_tp_rank_{r:02d}_pp_rank_00.pt")
            q = sh["model.layers.0.self_attn.qkv_proj.weight_q"][::HD, 0].long().tolist()          # original q head ids on this rank
            kk = sh["model.layers.0.self_attn.qkv_proj.weight_k"][::HD, 0].long().tolist()         # original kv head ids on this rank
            o = sh["model.layers.0.self_attn.o_proj.weight"][0, ::HD].long().tolist()
            assert len(q) == QH // TP and len(kk) == KVH * MULT // TP and o == q
            per = len(q) // len(kk)
            assert all(h // (QH // KVH) == kk[i // per] for i, h in enumerate(q)), (layout, r, q, kk)
            assert sh["model.layers.0.mlp.gate_up_proj.weight"].shape == (48 // TP, H)
            assert sh["model.layers.0.mlp.expert_mlps.w"][:, 0, 0].tolist() == [0.0, 1.0]           # ep rank 0: experts 0-1
            ep1 = torch.load(out / "converted" / "model" / f"dp_rank_00_ep_rank_01_tp_rank_{r:02d}_pp_rank_00.pt")
            assert set(ep1) == {f"model.layers.{l}.mlp.expert_mlps.w" for l in range(L)} and ep1["model.layers.1.mlp.expert_mlps.w"][:, 0, 0].tolist() == [2.0, 3.0]
        back = tmp_path / f"back_{layout}"
        main(["--input_dir", str(out), "--output_dir", str(back), "--convert_to_full_state", "--tp_size", str(TP), "--qkv_linear", "true",
              "--kv_size_multiplier", str(MULT), "--kv_replication_layout", layout, "--config", str(cfg), "--ep_size", "2"])
        rec = torch.load(back / "pytorch_model.bin")
        for k, v in full.items():
            if "gate_proj" in k or "up_proj" in k:
                continue
            torch.testing.assert_close(rec[k], v, msg=lambda m: f"{layout} {k}: {m}")
        gu = rec["model.layers.0.mlp.gate_up_proj.weight"]
        torch.testing.assert_close(gu[:24], full["model.layers.0.mlp.gate_proj.weight"])

    # (2) fused QKV parameter, virtual pipeline stages, xser round trip, legacy layout
    out = tmp_path / "fused"
    main(["--input_dir", str(src), "--output_dir", str(out), "--convert_from_full_state", "--tp_size", "2", "--pp_size", "2",
          "--qkv_linear", "true", "--fuse_qkv", "true", "--config", str(cfg), "--save_xser", "true"])
    assert os.path.isdir(out / "converted" / "model" / "dp_rank_00_tp_rank_00_pp_rank_01.pt.tensors")
    plain = tmp_path / "plain"
    main(["--input_dir", str(out), "--output_dir", str(plain), "--convert_from_xser", "--tp_size", "2", "--pp_size", "2"])
    s0 = torch.load(plain / "converted" / "model" / "dp_rank_00_tp_rank_01_pp_rank_00.pt")
    assert "model.layers.0.self_attn.qkv_proj.weight_qkv" in s0 and "model.layers.1.self_attn.qkv_proj.weight_qkv" not in s0
    assert s0["model.layers.0.self_attn.qkv_proj.weight_qkv"].shape == ((QH + 2 * KVH) * HD // 2, H) and "lm_head.weight" not in s0
    again = tmp_path / "xser2"
    main(["--input_dir", str(plain), "--output_dir", str(again), "--convert_to_xser", "--tp_size", "2", "--pp_size", "2"])
    back = tmp_path / "back_fused"
    main(["--input_dir", str(again), "--output_dir", str(back), "--convert_to_full_state", "--tp_size", "2", "--pp_size", "2",
          "--qkv_linear", "true", "--config", str(cfg), "--load_xser", "true"])
    rec = torch.load(back / "pytorch_model.bin")
    torch.testing.assert_close(rec["model.layers.1.self_attn.k_proj.weight"], full["model.layers.1.self_attn.k_proj.weight"])
    leg = tmp_path / "legacy"
    main(["--input_dir", str(src), "--output_dir", str(leg), "--convert_from_full_state", "--tp_size", "2", "--legacy_format"])
    assert "model" in torch.load(leg / "tp_rank_01_pp_rank_00" / "checkpoint.pt")

    # (3) overridable predicates / name maps
    c = CheckpointConverterBase()
    assert c.get_partition_dim("model.layers.0.mlp.down_proj.weight") == 1 and c.get_partition_dim("lm_head.weight") == 0
    with pytest.raises(AssertionError):
        c.get_partition_dim("model.norm.weight")
    assert c.is_qkv_weight("x.query_key_value.weight") and c.get_fused_qkv_key() == "qkv_proj.weight_qkv"
    h2n, n2h = c.get_hf_to_nxd_model_keys(True, True)
    assert c.get_weight_key(h2n, n2h, "model.layers.0.self_attn.k_proj.weight", True) == "model.layers.0.self_attn.qkv_proj.weight_k"
    assert c.get_weight_key(h2n, n2h, "model.layers.0.self_attn.qkv_proj.weight_v", False) == "model.layers.0.self_attn.v_proj.weight"
    mg = c.rename_keys_for_megatron("model.layers.3.self_attn.o_proj.weight", "megatron", True)
    assert mg == "language_model.encoder.layers.3.self_attention.dense.weight"
    assert c.rename_keys_for_megatron(mg, "megatron", False) == "model.layers.3.self_attn.o_proj.weight"
    assert c.rename_keys_for_megatron("a.b", "hf") == "a.b"

    class Dbrx(CheckpointConverterBase):
        attribute_map = {"num_key_value_heads": "attn_config.kv_n_heads"}

    assert Dbrx()._get_config_value({"attn_config": {"kv_n_heads": 8}}, "num_key_value_heads") == 8
    with pytest.raises(KeyError):
        Dbrx()._get_config_value({}, "num_key_value_heads")
    assert c.find_size({"a": torch.zeros(4), "n": {"b": torch.zeros(2, dtype=torch.float16)}}) == 20
    mha = {f"model.layers.0.self_attn.{n}_proj.weight": torch.full((4, 2), float(i)) for i, n in enumerate("qkv")}
    co = c.coalesce_qkv(mha, {"num_hidden_layers": 1}, 2)["model.layers.0.self_attn.qkv_proj.weight"]
    assert co[:, 0].tolist() == [0, 0, 1, 1, 2, 2] * 2                                           # rank-major [q_r; k_r; v_r]
    with pytest.raises(AssertionError):
        c.run(c.get_arg_parser().parse_args(["--output_dir", "x"]))
