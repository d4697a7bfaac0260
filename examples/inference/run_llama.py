#!/usr/bin/env python
"""Llama inference sample + latency benchmark (BASELINE config 5: Llama-2-13B, TP=8, batch 1, seq 2048) — counterpart of the
reference's ``examples/inference/run_llama.py`` / ``runner.py`` (trace → load → generate → benchmark → report).

  torchrun --nproc-per-node 8 examples/inference/run_llama.py --model 13b --tp_degree 8 --batch_size 1 \
      --max_prompt_length 1024 --sequence_length 2048 --benchmark
  # GQA with fewer KV heads than ranks: replicas of a KV head shard its cache along the sequence (flash decoding)
  torchrun --nproc-per-node 8 examples/inference/run_llama.py --model 7b --num_kv_heads 4 --tp_degree 8 --flash_decoding
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples", "training"))

from neuronx_distributed_b200.inference.autobucketing import generate_buckets  # noqa: E402
from neuronx_distributed_b200.inference.benchmark import Benchmark, generate_report  # noqa: E402
from neuronx_distributed_b200.trace.model_builder import ModelBuilder  # noqa: E402
from neuronx_distributed_b200.models.llama import LlamaConfig, llama2_7b_config, llama2_13b_config  # noqa: E402
from neuronx_distributed_b200.models.llama_inference import LlamaForInference  # noqa: E402
from neuronx_distributed_b200.parallel_layers import parallel_state as ps  # noqa: E402
from training_utils import init_distributed  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--model", default="tiny", choices=["tiny", "7b", "13b"])
    p.add_argument("--tp_degree", type=int, default=1)
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--max_prompt_length", type=int, default=128)
    p.add_argument("--sequence_length", type=int, default=256)
    p.add_argument("--num_layers", type=int, default=-1)
    p.add_argument("--benchmark", action="store_true")
    p.add_argument("--num_runs", type=int, default=20)
    p.add_argument("--no_cuda_graphs", action="store_true")
    p.add_argument("--flash_decoding", action="store_true",
                   help="shard the KV cache along the sequence inside each KV-replica group (needs tp_degree > number of KV heads)")
    p.add_argument("--num_kv_heads", type=int, default=0, help="override the number of KV heads (GQA experiments)")
    a = p.parse_args()
    dev = init_distributed()
    ps.initialize_model_parallel(tensor_model_parallel_size=a.tp_degree)
    dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
    kw = dict(dtype=dtype, device=dev, max_position_embeddings=a.sequence_length)
    cfg = {"7b": llama2_7b_config, "13b": llama2_13b_config}.get(a.model, lambda **k: LlamaConfig(
        vocab_size=4096, hidden_size=512, intermediate_size=1408, num_hidden_layers=4, num_attention_heads=8, **k))(**kw)
    if a.num_layers > 0:
        cfg.num_hidden_layers = a.num_layers
    if a.num_kv_heads > 0:
        cfg.num_key_value_heads = a.num_kv_heads
    torch.manual_seed(0)
    model = LlamaForInference(cfg, batch_size=a.batch_size, max_seq_len=a.sequence_length,
                              flash_decoding=a.flash_decoding).eval()
    B, P = a.batch_size, a.max_prompt_length
    mb = ModelBuilder(tp_degree=a.tp_degree, use_cuda_graphs=(dev.type == "cuda" and not a.no_cuda_graphs))
    for bucket in generate_buckets(min(128, P), P):
        mb.add("context_encoding_model", model, [(torch.zeros(B, bucket, dtype=torch.long, device=dev),
                                                  torch.full((B,), bucket - 1, dtype=torch.long, device=dev))],
               step_fn=lambda m, i, l: m.context_encoding(i, l))
    mb.add("token_generation_model", model, [(torch.zeros(B, 1, dtype=torch.long, device=dev),
                                              torch.full((B,), P, dtype=torch.long, device=dev))],
           step_fn=lambda m, i, pos: m.token_generation(i, pos))
    nxd_model = mb.trace()
    prompt = torch.randint(0, cfg.vocab_size, (B, P), device=dev)
    new_tokens = a.sequence_length - P

    def generate():
        tok = nxd_model(prompt, torch.full((B,), P - 1, dtype=torch.long, device=dev)).clone()
        pos = torch.full((B,), P, dtype=torch.long, device=dev)
        out = [tok]
        for _ in range(new_tokens - 1):
            tok = nxd_model(tok.view(B, 1), pos).clone()
            pos = pos + 1
            out.append(tok)
        return torch.stack(out, 1)

    toks = generate()
    if dist.get_rank() == 0:
        print("generated", tuple(toks.shape), toks[0, :8].tolist(), flush=True)
    if a.benchmark:
        lat = Benchmark(generate, num_runs=a.num_runs).run()
        rep = generate_report(lat, a.sequence_length, B)
        if dist.get_rank() == 0:
            print(json.dumps({"e2e_model": rep, "config": vars(a)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
