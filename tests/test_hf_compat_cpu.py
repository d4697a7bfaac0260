"""HF interoperability: a randomly initialised ``transformers`` model is loaded into the TP-sharded built-in model straight from
its HF state dict / checkpoint directory and must reproduce the HF logits; gathering the sharded model back yields the HF
tensors bit-exactly (reference flow: ``examples/training/*/convert_checkpoints.py`` + ``scripts/checkpoint_converter.py``,
``examples/inference/modules/checkpoint.py``)."""
import os

import pytest
import torch

from dist_utils import run_distributed

transformers = pytest.importorskip("transformers")


def _hf(fam: str, kvh: int):
    torch.manual_seed(0)
    common = dict(vocab_size=128, hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=8,
                  num_key_value_heads=kvh, max_position_embeddings=64, attn_implementation="eager")
    if fam == "llama":
        hc = transformers.LlamaConfig(**common)
        return hc, transformers.LlamaForCausalLM(hc).eval()
    hc = transformers.MixtralConfig(num_local_experts=4, num_experts_per_tok=2, sliding_window=None, **common)
    return hc, transformers.MixtralForCausalLM(hc).eval()


def _parity(rank, world, fam, kvh, tmp):
    from neuronx_distributed_b200.models import hf_compat
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.parallel_layers.mappings import gather_from_tensor_model_parallel_region

    ps.initialize_model_parallel(world)
    hc, hf = _hf(fam, kvh)
    sd = hf.state_dict()
    cfg = hf_compat.config_from_hf(hc, dtype=torch.float32)
    assert cfg.rope_theta == (1e6 if fam == "mixtral" else 1e4) and cfg.num_key_value_heads == kvh
    if fam == "llama":
        from neuronx_distributed_b200.models.llama import LlamaForCausalLM as Model
    else:
        from neuronx_distributed_b200.models.mixtral import MixtralForCausalLM as Model
    model = Model(cfg).eval()
    res = hf_compat.load_hf_checkpoint(model, sd)
    assert not res.missing_keys and not res.unexpected_keys
    ids = torch.randint(0, 128, (2, 16), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = hf(ids).logits
        _, lg = model(ids)
    lg = gather_from_tensor_model_parallel_region(lg).transpose(0, 1)
    assert float((lg - ref).abs().max()) < 2e-5

    # sharded model → HF names; bit-exact round trip (KV replicas dropped, Q heads un-permuted, experts un-stacked)
    style = None if fam == "llama" else "fused_experts"           # transformers>=5 keeps Mixtral experts as 3-D parameters
    back = hf_compat.nxd_to_hf_state_dict(hf_compat.gather_full_state_dict(model), cfg, style=style, **hf_compat._kv_args(model))
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)

    # through files: save_hf_checkpoint (sharded safetensors + index) → a fresh model loaded from the directory
    out = os.path.join(tmp, "hf_out")
    hf_compat.save_hf_checkpoint(model, out, hf_config=hc, max_shard_bytes=64 << 10)
    assert os.path.isfile(os.path.join(out, "model.safetensors.index.json")) and os.path.isfile(os.path.join(out, "config.json"))
    cfg2 = hf_compat.config_from_hf(out, dtype=torch.float32)
    model2 = Model(cfg2).eval()
    hf_compat.load_hf_checkpoint(model2, out)
    with torch.no_grad():
        _, lg2 = model2(ids)
    assert torch.equal(gather_from_tensor_model_parallel_region(lg2).transpose(0, 1), lg)
    if fam == "mixtral":                                            # the on-disk spelling is per expert (block_sparse_moe.experts.E.w1…)
        disk = hf_compat.read_hf_state_dict(out)
        assert "model.layers.0.block_sparse_moe.experts.3.w2.weight" in disk and disk["model.layers.0.block_sparse_moe.experts.3.w2.weight"].shape == (64, 96)


@pytest.mark.parametrize("fam,world,kvh", [("llama", 2, 2), ("llama", 4, 2), ("mixtral", 2, 2)])
def test_hf_checkpoint_parity(tmp_path, fam, world, kvh):
    """``llama, 4, 2``: two KV heads on four ranks → KV replicated ×2 (tile layout) with the Q-head / o_proj permutation."""
    run_distributed(_parity, world, fam, kvh, str(tmp_path), timeout=300)


def _serving_and_dbrx(rank, world):
    from neuronx_distributed_b200.models import hf_compat
    from neuronx_distributed_b200.models.llama_inference import LlamaForInference
    from neuronx_distributed_b200.models.mixtral import DbrxConfig, MixtralForCausalLM
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(world)
    # serving wrapper: greedy continuation equals HF generate
    hc, hf = _hf("llama", 2)
    cfg = hf_compat.config_from_hf(hc, dtype=torch.float32)
    srv = LlamaForInference(cfg, batch_size=1, max_seq_len=32).eval()
    hf_compat.load_hf_checkpoint(srv, hf.state_dict())
    prompt = torch.randint(0, 128, (1, 8), generator=torch.Generator().manual_seed(2))
    want = hf.generate(prompt, max_new_tokens=6, do_sample=False, pad_token_id=0)[0, 8:]
    tok = srv.context_encoding(prompt, torch.tensor([7]))
    got = [int(tok[0])]
    for i in range(5):
        tok = srv.token_generation(tok.view(1, 1), torch.tensor([8 + i]))
        got.append(int(tok[0]))
    assert got == want.tolist(), (got, want.tolist())

    # DBRX spelling ↔ ours: concatenated experts, fused Wqkv, bit-exact both ways and equal to the Mixtral-spelled conversion
    dcfg = DbrxConfig(vocab_size=128, hidden_size=64, intermediate_size=96, num_hidden_layers=1, num_attention_heads=8, num_key_value_heads=2,
                      num_local_experts=4, num_experts_per_tok=2, dtype=torch.float32, max_position_embeddings=64)
    g = torch.Generator().manual_seed(5)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    b = "transformer.blocks.0."
    dbrx = {"transformer.wte.weight": r(128, 64), "transformer.norm_f.weight": r(64), "lm_head.weight": r(128, 64),
            b + "norm_attn_norm.norm_1.weight": r(64), b + "norm_attn_norm.norm_2.weight": r(64),
            b + "norm_attn_norm.attn.Wqkv.weight": r(64 + 16 + 16, 64), b + "norm_attn_norm.attn.out_proj.weight": r(64, 64),
            b + "ffn.router.layer.weight": r(4, 64), b + "ffn.experts.mlp.w1": r(4 * 96, 64), b + "ffn.experts.mlp.v1": r(4 * 96, 64),
            b + "ffn.experts.mlp.w2": r(4 * 96, 64)}
    ours = hf_compat.hf_to_nxd_state_dict(dbrx, dcfg)
    e2 = dbrx[b + "ffn.experts.mlp.w1"].view(4, 96, 64)[2]                      # expert 2: gate = x · w1ᵀ, down = h · w2
    assert torch.equal(ours["layers.0.mlp.expert_mlps.mlp_op.gate_up_proj.weight"][2, :, :96], e2.t())
    assert torch.equal(ours["layers.0.mlp.expert_mlps.mlp_op.down_proj.weight"][2], dbrx[b + "ffn.experts.mlp.w2"].view(4, 96, 64)[2])
    model = MixtralForCausalLM(dcfg).eval()
    res = hf_compat.load_hf_checkpoint(model, dbrx)
    assert not res.missing_keys and not res.unexpected_keys
    back = hf_compat.nxd_to_hf_state_dict(hf_compat.gather_full_state_dict(model), dcfg, style="dbrx", **hf_compat._kv_args(model))
    assert set(back) == set(dbrx) and all(torch.equal(back[k], dbrx[k]) for k in dbrx)
    d = hf_compat.config_from_hf({"model_type": "dbrx", "d_model": 64, "n_heads": 8, "n_layers": 1, "max_seq_len": 64, "vocab_size": 128,
                                  "attn_config": {"kv_n_heads": 2, "clip_qkv": 8.0, "rope_theta": 5e5},
                                  "ffn_config": {"ffn_hidden_size": 96, "moe_num_experts": 4, "moe_top_k": 2}})
    assert isinstance(d, DbrxConfig) and d.rope_theta == 5e5 and d.num_local_experts == 4 and d.intermediate_size == 96


def test_hf_serving_wrapper_and_dbrx_spelling():
    run_distributed(_serving_and_dbrx, 2, timeout=300)


def _cli_shards(rank, world, out_dir, hf_dir):
    """Every rank loads ITS file written by the converter CLI and compares it with the in-process HF load."""
    from neuronx_distributed_b200.models import hf_compat
    from neuronx_distributed_b200.models.mixtral import MixtralForCausalLM
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(world)
    cfg = hf_compat.config_from_hf(hf_dir, dtype=torch.float32)
    model = MixtralForCausalLM(cfg)
    hf_compat.load_hf_checkpoint(model, os.path.join(hf_dir, "pytorch_model.bin"))
    want = model.state_dict()
    got = torch.load(os.path.join(out_dir, "converted", "model", f"dp_rank_00_tp_rank_{rank:02d}_pp_rank_00.pt"), weights_only=True)
    # the converter writes un-fused q/k/v shards; the live module keeps them fused per rank as [q_r; k_r; v_r]
    for k in [k for k in got if k.endswith("qkv_proj.weight_q")]:
        b = k[: -len("weight_q")]
        got[b + "weight_qkv"] = torch.cat([got.pop(b + "weight_q"), got.pop(b + "weight_k"), got.pop(b + "weight_v")], 0)
    assert set(got) == set(want), set(got) ^ set(want)
    for k in want:
        assert torch.equal(got[k], want[k]), k


def test_moe_convert_checkpoints_cli(tmp_path):
    """``examples/training/mixtral/convert_checkpoints.py``: HF per-expert files → TP=4 shards (KV ×2) → back, bit-exact."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hc, hf = _hf("mixtral", 2)
    from neuronx_distributed_b200.models import hf_compat

    cfg = hf_compat.config_from_hf(hc)
    # on-disk HF spelling (per expert), derived from the in-memory fused one
    disk = hf_compat.nxd_to_hf_state_dict(hf_compat.hf_to_nxd_state_dict(hf.state_dict(), cfg), cfg)
    hf_dir, out, merged = str(tmp_path / "hf"), str(tmp_path / "sharded"), str(tmp_path / "merged")
    os.makedirs(hf_dir)
    torch.save(disk, os.path.join(hf_dir, "pytorch_model.bin"))
    with open(os.path.join(hf_dir, "config.json"), "w") as f:
        json.dump(hc.to_dict(), f, default=str)
    cli = [sys.executable, os.path.join(root, "examples", "training", "mixtral", "convert_checkpoints.py"), "--config",
           os.path.join(hf_dir, "config.json"), "--tp_size", "4", "--kv_size_multiplier", "2"]
    subprocess.run(cli + ["--convert_from_full_state", "--input_dir", hf_dir, "--output_dir", out], check=True, timeout=300)
    run_distributed(_cli_shards, 4, out, hf_dir, timeout=300)
    subprocess.run(cli + ["--convert_to_full_state", "--input_dir", out, "--output_dir", merged], check=True, timeout=300)
    back = torch.load(os.path.join(merged, "pytorch_model.bin"), weights_only=True)
    assert set(back) == set(disk) and all(torch.equal(back[k], disk[k]) for k in disk)


def _encoders(rank, world):
    """BERT, ViT and GPT-NeoX: HF state dict → TP-sharded built-in model → same outputs as ``transformers``; and back."""
    from neuronx_distributed_b200.models import bert, gpt_neox, hf_compat, vit
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.parallel_layers.mappings import gather_from_tensor_model_parallel_region as gather

    ps.initialize_model_parallel(world)
    torch.manual_seed(0)
    ids = torch.randint(0, 64, (2, 16), generator=torch.Generator().manual_seed(1))

    # ---- BERT (MLM + NSP heads)
    hc = transformers.BertConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                                 max_position_embeddings=32, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, attn_implementation="eager")
    hf = transformers.BertForPreTraining(hc).eval()
    with torch.no_grad():
        hf.cls.predictions.bias.normal_()                                     # tied to decoder.bias in HF
    m = bert.BertForPreTraining(bert.BertConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                                                max_position_embeddings=32, hidden_dropout_prob=0.0, dtype=torch.float32)).eval()
    res = hf_compat.load_hf_checkpoint(m, hf.state_dict())
    assert not res.missing_keys and not res.unexpected_keys, res
    tt = torch.zeros_like(ids)
    with torch.no_grad():
        ref = hf(input_ids=ids, token_type_ids=tt)
        _, (mlm, nsp) = m(ids, tt)
    assert float((gather(mlm) - ref.prediction_logits).abs().max()) < 2e-5
    assert float((nsp - ref.seq_relationship_logits).abs().max()) < 2e-5

    # ---- ViT
    hc = transformers.ViTConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4, image_size=32, patch_size=8,
                                num_labels=10, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, attn_implementation="eager")
    hf = transformers.ViTForImageClassification(hc).eval()
    v = vit.ViTForImageClassification(vit.ViTConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4, image_size=32,
                                                    patch_size=8, num_labels=10, dtype=torch.float32)).eval()
    res = hf_compat.load_hf_checkpoint(v, hf.state_dict())
    assert not res.missing_keys and not res.unexpected_keys, res
    px = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        assert float((v(px) - hf(pixel_values=px).logits).abs().max()) < 2e-5
    sd = hf.state_dict()
    back = hf_compat.nxd_to_hf_vit_state_dict(hf_compat.hf_to_nxd_vit_state_dict(sd))
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)

    # ---- GPT-NeoX (HF names and per-head interleaved QKV are used as they are)
    hc = transformers.GPTNeoXConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                                    max_position_embeddings=32, rotary_pct=0.25, attn_implementation="eager")
    hf = transformers.GPTNeoXForCausalLM(hc).eval()
    n = gpt_neox.GPTNeoXForCausalLM(gpt_neox.GPTNeoXConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                                                           num_attention_heads=4, max_position_embeddings=32, dtype=torch.float32)).eval()
    res = hf_compat.load_hf_checkpoint(n, hf.state_dict())
    assert not res.missing_keys and not res.unexpected_keys, res
    with torch.no_grad():
        out = n(ids)
    lg = gather(out[1] if isinstance(out, tuple) else out)
    lg = lg.transpose(0, 1) if lg.shape[0] == 16 else lg
    assert float((lg - hf(ids).logits).abs().max()) < 2e-5

    # BERT names round trip
    hb = transformers.BertForPreTraining(transformers.BertConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=1,
                                                                 num_attention_heads=4, max_position_embeddings=32)).state_dict()
    hb = {k: v for k, v in hb.items() if not k.endswith("position_ids")}
    back = hf_compat.nxd_to_hf_bert_state_dict(hf_compat.hf_to_nxd_bert_state_dict(hb))
    assert set(back) == set(hb) and all(torch.equal(back[k], hb[k]) for k in hb)


def test_hf_bert_vit_neox_parity():
    run_distributed(_encoders, 2, timeout=300)


def _adapter(rank, world):
    """``HuggingFaceGenerationAdapter.generate``: left-padded batch, greedy tokens equal ``transformers``' generate; EOS fills
    with pad; sampling draws identical tokens on every TP rank; logits processors see the full vocabulary."""
    import torch.distributed as dist

    from neuronx_distributed_b200.inference.hf_adapter import HuggingFaceGenerationAdapter
    from neuronx_distributed_b200.models import hf_compat
    from neuronx_distributed_b200.models.llama_inference import LlamaForInference
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(world)
    hc, hf = _hf("llama", 2)
    cfg = hf_compat.config_from_hf(hc, dtype=torch.float32)
    srv = LlamaForInference(cfg, batch_size=2, max_seq_len=32).eval()
    hf_compat.load_hf_checkpoint(srv, hf.state_dict())
    g = torch.Generator().manual_seed(3)
    a, b = torch.randint(1, 128, (9,), generator=g), torch.randint(1, 128, (5,), generator=g)
    ids = torch.zeros(2, 9, dtype=torch.long)
    mask = torch.zeros(2, 9, dtype=torch.long)
    ids[0], mask[0] = a, 1
    ids[1, 4:], mask[1, 4:] = b, 1                                               # left padded, as HF tokenizers do for generation
    want = hf.generate(ids, attention_mask=mask, max_new_tokens=6, do_sample=False, pad_token_id=0)
    gen = HuggingFaceGenerationAdapter(srv, pad_token_id=0)
    got = gen.generate(ids, attention_mask=mask, max_new_tokens=6)
    assert torch.equal(got, want), (got, want)
    # EOS: stop row 0 at its 3rd generated token; the rest of that row is pad, the other row continues
    eos = int(want[0, 9 + 2])
    got = gen.generate(ids, attention_mask=mask, max_new_tokens=6, eos_token_id=eos)
    want_eos = hf.generate(ids, attention_mask=mask, max_new_tokens=6, do_sample=False, pad_token_id=0, eos_token_id=eos)
    assert torch.equal(got[:, : want_eos.shape[1]], want_eos) and bool((got[:, want_eos.shape[1]:] == 0).all())
    # sampling: every TP rank must continue with the same token
    torch.manual_seed(100 + rank)                                                # deliberately different RNG streams
    s = gen.generate(ids, attention_mask=mask, max_new_tokens=5, do_sample=True, top_k=8, top_p=0.9, temperature=0.7)
    ref = s.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(s, ref) and s.shape == (2, 14)
    # a logits processor that forbids everything except token 7
    only7 = lambda hist, logits: torch.full_like(logits, float("-inf")).index_fill_(-1, torch.tensor([7]), 0.0)  # noqa: E731
    assert bool((gen.generate(ids, attention_mask=mask, max_new_tokens=3, logits_processor=[only7])[:, 9:] == 7).all())
    assert srv.on_device_sampling                                               # restored


def test_hf_generation_adapter():
    run_distributed(_adapter, 2, timeout=300)


def _pp_load(rank, world):
    """HF checkpoint → TP=2 × PP=2 partitioned model (each rank keeps its stage's shard only) → the HF loss."""
    import neuronx_distributed_b200 as nxd
    from neuronx_distributed_b200.models import hf_compat
    from neuronx_distributed_b200.models.llama import LlamaDecoderLayer, LlamaForCausalLM

    torch.manual_seed(0)
    hc = transformers.LlamaConfig(vocab_size=128, hidden_size=64, intermediate_size=96, num_hidden_layers=4, num_attention_heads=8,
                                  num_key_value_heads=2, max_position_embeddings=64, attn_implementation="eager")
    hf = transformers.LlamaForCausalLM(hc).eval()
    cfg = nxd.neuronx_distributed_config(tensor_parallel_size=2, pipeline_parallel_size=2, pipeline_config={
        "num_microbatches": 2, "input_names": ["input_ids", "labels"], "output_loss_value_spec": True, "auto_partition": True,
        "transformer_layer_cls": LlamaDecoderLayer})
    mcfg = hf_compat.config_from_hf(hc, dtype=torch.float32)
    model = nxd.initialize_parallel_model(cfg, lambda: LlamaForCausalLM(mcfg))
    res = hf_compat.load_hf_checkpoint(model, hf.state_dict())
    assert not res.missing_keys
    assert all(("layers.0." in k or "layers.1." in k or "embed" in k) == (rank < 2) for k in model.state_dict() if "norm.weight" != k[-11:] or "layers" in k)
    ids = torch.randint(0, 128, (4, 16), generator=torch.Generator().manual_seed(1))
    loss = model.run_eval(input_ids=ids, labels=ids)
    with torch.no_grad():
        ref = hf(ids, labels=ids).loss
    if rank >= 2:                                                  # last stage owns the loss
        assert abs(float(loss) - float(ref)) < 1e-5, (float(loss), float(ref))


def test_hf_load_into_tp_pp_model():
    run_distributed(_pp_load, 4, timeout=300)


def _mixtral_serving(rank, world):
    """MoE family through the serving wrapper (``lm_cls=MixtralForCausalLM``): HF Mixtral weights, greedy == ``generate``."""
    from neuronx_distributed_b200.inference.hf_adapter import HuggingFaceGenerationAdapter
    from neuronx_distributed_b200.models import hf_compat
    from neuronx_distributed_b200.models.llama_inference import LlamaForInference
    from neuronx_distributed_b200.models.mixtral import MixtralForCausalLM
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(world)
    hc, hf = _hf("mixtral", 2)
    cfg = hf_compat.config_from_hf(hc, dtype=torch.float32)
    srv = LlamaForInference(cfg, batch_size=1, max_seq_len=32, lm_cls=MixtralForCausalLM).eval()
    res = hf_compat.load_hf_checkpoint(srv, hf.state_dict())
    assert not res.missing_keys and not res.unexpected_keys
    prompt = torch.randint(1, 128, (1, 7), generator=torch.Generator().manual_seed(4))
    want = hf.generate(prompt, max_new_tokens=6, do_sample=False, pad_token_id=0)
    got = HuggingFaceGenerationAdapter(srv, pad_token_id=0).generate(prompt, max_new_tokens=6)
    assert torch.equal(got, want), (got, want)


def test_hf_mixtral_serving_generation():
    run_distributed(_mixtral_serving, 2, timeout=300)
