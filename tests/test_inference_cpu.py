import os

import torch

from dist_utils import run_distributed


def _infer(rank, world, tmp):
    from neuronx_distributed_b200.trace.model_builder import ModelBuilder, shard_checkpoint
    from neuronx_distributed_b200.models.llama import LlamaConfig
    from neuronx_distributed_b200.models.llama_inference import LlamaForInference
    from neuronx_distributed_b200.operators import argmax
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.parallel_layers.utils import gather_full_weight
    from neuronx_distributed_b200.utils.safetensors_utils import load_state_dict_safetensors

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    cfg = LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, dtype=torch.float32, max_position_embeddings=32)
    torch.manual_seed(0)
    m = LlamaForInference(cfg, batch_size=2, max_seq_len=32).eval()
    ids = torch.randint(0, 64, (2, 8), generator=torch.Generator().manual_seed(1))
    gen = m.generate(ids, 5)
    assert gen.shape == (2, 5)
    # greedy generation with the KV cache == greedy re-running the full model on the growing sequence
    seq, ref = ids.clone(), []
    with torch.no_grad():
        for _ in range(5):
            _, logits = m.lm(seq)                       # [S, B, V/tp]
            nxt = argmax(logits[-1].float(), dim=-1)
            ref.append(nxt)
            seq = torch.cat([seq, nxt.view(2, 1)], 1)
    assert torch.equal(gen, torch.stack(ref, 1)), (gen, torch.stack(ref, 1))
    # bucketed runtime: prefill + decode programs routed by input shape
    mb = ModelBuilder(tp_degree=world, use_cuda_graphs=False)
    mb.add("context_encoding_model", m, [(ids, torch.tensor([7, 7]))], step_fn=lambda mod, i, l: mod.context_encoding(i, l))
    mb.add("token_generation_model", m, [(ids[:, :1], torch.tensor([8, 8]))], step_fn=lambda mod, i, p: mod.token_generation(i, p))
    nxd = mb.trace()
    m.kv.reset()
    t0 = nxd(ids, torch.tensor([7, 7]))
    t1 = nxd(t0.view(2, 1), torch.tensor([8, 8]))
    assert torch.equal(torch.stack([t0, t1], 1), gen[:, :2])
    if world == 2:
        import torch.distributed as dist

        sd = m.lm.state_dict()
        params = dict(m.lm.named_parameters())
        full = {}
        for k, v in sd.items():
            p = params[k]
            if getattr(p, "tensor_model_parallel", False):
                parts = [torch.empty_like(v) for _ in range(world)]
                dist.all_gather(parts, v.contiguous())
                full[k] = torch.cat(parts, 0) if getattr(p, "fused_qkv", False) else \
                    gather_full_weight(parts, p.partition_dim, p.partition_stride)
            else:
                full[k] = v
        shards = shard_checkpoint(full, m.lm, world, serialize_path=tmp if rank == 0 else None)
        for k in sd:
            torch.testing.assert_close(shards[rank][k], sd[k])
        dist.barrier()
        assert os.path.isfile(os.path.join(tmp, "tp0_sharded_checkpoint.safetensors"))
        back = load_state_dict_safetensors(os.path.join(tmp, f"tp{rank}_sharded_checkpoint.safetensors"))
        torch.testing.assert_close(back["lm_head.weight"], sd["lm_head.weight"])


def test_llama_inference_kv_cache_and_buckets_tp1(tmp_path):
    run_distributed(_infer, 1, str(tmp_path), timeout=120)


def test_llama_inference_tp2_and_shard_checkpoint(tmp_path):
    run_distributed(_infer, 2, str(tmp_path), timeout=120)


def test_benchmark_report():
    from neuronx_distributed_b200.inference.benchmark import Benchmark, generate_report

    b = Benchmark(lambda: sum(range(1000)), num_runs=5)
    rep = generate_report(b.run(), 128, 2)
    assert set(rep) >= {"latency_ms_p50", "latency_ms_p99", "latency_ms_avg", "throughput"} and rep["throughput"] > 0


def _speculative(rank, world):
    """Speculative decoding (draft proposes, target verifies a window in one forward) reproduces the target's own greedy
    decoding token for token — with a draft that shares the target's weights (all accepted) and a random draft (mostly rejected)."""
    from neuronx_distributed_b200.models.llama import LlamaConfig
    from neuronx_distributed_b200.models.llama_inference import LlamaForInference
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.utils.speculative import speculative_generate

    ps.initialize_model_parallel(tensor_model_parallel_size=world)

    def make(seed, layers):
        torch.manual_seed(seed)
        cfg = LlamaConfig(vocab_size=96, hidden_size=32, intermediate_size=64, num_hidden_layers=layers, num_attention_heads=4,
                          num_key_value_heads=2, dtype=torch.float32, max_position_embeddings=64)
        return LlamaForInference(cfg, batch_size=1, max_seq_len=64).eval()

    target = make(1, 2)
    prompt = torch.randint(0, 96, (1, 7), generator=torch.Generator().manual_seed(9))
    want = target.generate(prompt, 12)
    for draft in (make(1, 2), make(5, 1)):
        target.kv.reset()
        got, acc = speculative_generate(target, draft, prompt, 12, speculation_length=3)
        assert torch.equal(got, want), (got, want, acc)
    # same-weights draft: every proposal accepted
    target.kv.reset()
    _, acc = speculative_generate(target, make(1, 2), prompt, 12, speculation_length=3)
    assert acc > 2.0


def test_speculative_decoding_matches_greedy():
    run_distributed(_speculative, 1, timeout=120)


def _flash_decode(rank, world):
    """Sequence-sharded KV cache inside a KV-replica group: distributed flash-decoding equals full attention over the whole cache."""
    import math

    import torch.distributed as dist

    from neuronx_distributed_b200.modules.attention.flash_decode import flash_decode_attention, write_decode_sharded

    g = dist.group.WORLD
    B, L, Hl, Hkv, D = 3, 16, 2, 1, 8                       # each rank has 2 query heads; both ranks share 1 kv head
    gen = torch.Generator().manual_seed(4)
    k_full = torch.randn(B, L, Hkv, D, generator=gen); v_full = torch.randn(B, L, Hkv, D, generator=gen)
    q_all = torch.randn(B, 1, world * Hl, D, generator=gen)
    positions = torch.tensor([3, 9, 15])
    l_local = L // world
    k_loc = k_full[:, rank * l_local:(rank + 1) * l_local].clone(); v_loc = v_full[:, rank * l_local:(rank + 1) * l_local].clone()
    # the newest token's K/V arrives through the sharded write
    k_new = torch.randn(B, 1, Hkv, D, generator=gen); v_new = torch.randn(B, 1, Hkv, D, generator=gen)
    write_decode_sharded(k_loc, v_loc, k_new, v_new, positions, rank)
    b = torch.arange(B)
    k_full[b, positions] = k_new[:, 0]; v_full[b, positions] = v_new[:, 0]
    q = q_all[:, :, rank * Hl:(rank + 1) * Hl]
    out = flash_decode_attention(q, k_loc, v_loc, positions, g)
    s = torch.einsum("bhd,blhd->bhl", q[:, 0], k_full.repeat_interleave(Hl, 2)) / math.sqrt(D)
    mask = torch.arange(L)[None, :] <= positions[:, None]
    ref = torch.einsum("bhl,blhd->bhd", s.masked_fill(~mask[:, None], float("-inf")).softmax(-1), v_full.repeat_interleave(Hl, 2))
    torch.testing.assert_close(out[:, 0], ref, rtol=1e-5, atol=1e-5)


def test_distributed_flash_decoding_matches_full_attention():
    run_distributed(_flash_decode, 2, timeout=60)


def _medusa(rank, world):
    """Medusa tree decoding (tree attention mask, path acceptance, KV compaction) equals plain greedy decoding."""
    from neuronx_distributed_b200.models.llama import LlamaConfig
    from neuronx_distributed_b200.models.llama_inference import LlamaForInference
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.utils.medusa import MedusaHeads, medusa_generate

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    torch.manual_seed(1)
    cfg = LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, dtype=torch.float32, max_position_embeddings=96)
    model = LlamaForInference(cfg, batch_size=1, max_seq_len=96).eval()
    prompt = torch.randint(0, 64, (1, 6), generator=torch.Generator().manual_seed(2))
    want = model.generate(prompt, 14)
    choices = [[0], [1], [0, 0], [0, 1], [1, 0], [0, 0, 0]]
    # (a) random heads: mostly rejected, output must still be exact
    torch.manual_seed(3)
    heads = MedusaHeads(32, 64, 3).eval()
    model.kv.reset()
    got, acc = medusa_generate(model, heads, prompt, 14, choices, topk=4)
    assert torch.equal(got, want), (got, want)
    # (b) "oracle" heads are impossible without training; instead check that acceptance > 0 happens on a degenerate model
    #     whose next token does not depend on context much: tie all heads to the lm_head so head k guesses the same token
    for p_, b_ in zip(heads.proj, heads.blocks):
        p_.weight.data.copy_(model.lm.lm_head.weight.data)
    model.kv.reset()
    got2, acc2 = medusa_generate(model, heads, prompt, 14, choices, topk=4)
    assert torch.equal(got2, want), (got2, want, acc2)


def test_medusa_tree_decoding_matches_greedy():
    run_distributed(_medusa, 1, timeout=120)


def _flash_decode_model(rank, world):
    """``LlamaForInference(flash_decoding=True)``: TP=4 over 2 KV heads → each KV head lives on 2 ranks, which shard its cache
    along the sequence; generation equals the replicated-cache model token for token."""
    from neuronx_distributed_b200.models.llama import LlamaConfig
    from neuronx_distributed_b200.models.llama_inference import LlamaForInference
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    cfg = dict(vocab_size=64, hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=8,
               num_key_value_heads=2, dtype=torch.float32, max_position_embeddings=32)
    torch.manual_seed(0)
    plain = LlamaForInference(LlamaConfig(**cfg), batch_size=2, max_seq_len=32).eval()
    torch.manual_seed(0)
    fd = LlamaForInference(LlamaConfig(**cfg), batch_size=2, max_seq_len=32, flash_decoding=True).eval()
    fd.load_state_dict(plain.state_dict())
    assert fd.kv.k[0].shape[1] == 16 and plain.kv.k[0].shape[1] == 32          # half the cache per rank
    ids = torch.randint(0, 64, (2, 20), generator=torch.Generator().manual_seed(1))   # prompt crosses the shard boundary (16)
    lens = torch.tensor([20, 13])
    want = plain.generate(ids, 8, prompt_lens=lens)
    got = fd.generate(ids, 8, prompt_lens=lens)
    assert torch.equal(got, want), (got, want)


def test_llama_inference_flash_decoding_tp4():
    run_distributed(_flash_decode_model, 4, timeout=240)
