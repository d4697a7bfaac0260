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
