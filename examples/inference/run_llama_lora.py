#!/usr/bin/env python
"""Llama inference with a LoRA adapter on the attention projections (adapter merged for serving, then unmerged) — counterpart of
the reference's ``examples/inference/run_llama_lora.py``.

  python examples/inference/run_llama_lora.py --lora_rank 8
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples", "training"))

from neuronx_distributed_b200.models.llama import LlamaConfig  # noqa: E402
from neuronx_distributed_b200.models.llama_inference import LlamaForInference  # noqa: E402
from neuronx_distributed_b200.modules.lora import LoraConfig, get_lora_model  # noqa: E402
from neuronx_distributed_b200.parallel_layers import parallel_state as ps  # noqa: E402
from training_utils import init_distributed  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--tp_degree", type=int, default=1)
    p.add_argument("--lora_rank", type=int, default=8)
    p.add_argument("--prompt_length", type=int, default=16)
    p.add_argument("--max_new_tokens", type=int, default=8)
    a = p.parse_args()
    dev = init_distributed()
    ps.initialize_model_parallel(tensor_model_parallel_size=a.tp_degree)
    dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
    L = a.prompt_length + a.max_new_tokens
    cfg = LlamaConfig(vocab_size=4096, hidden_size=256, intermediate_size=704, num_hidden_layers=2, num_attention_heads=8,
                      dtype=dtype, device=dev, max_position_embeddings=L)
    torch.manual_seed(0)
    model = LlamaForInference(cfg, batch_size=1, max_seq_len=L).eval()
    prompt = torch.randint(0, cfg.vocab_size, (1, a.prompt_length), device=dev)
    base = model.generate(prompt, a.max_new_tokens)
    lcfg = LoraConfig(lora_rank=a.lora_rank, lora_alpha=2 * a.lora_rank, target_modules=["o_proj", "down_proj"])
    lora = get_lora_model(model.lm, lcfg)
    n_adapters = 0
    for mod in lora.modules():                      # stand-in for a trained adapter: random B matrices
        if hasattr(mod, "lora_B") and hasattr(mod.lora_B, "weight"):
            nn.init.normal_(mod.lora_B.weight, std=0.05)
            n_adapters += 1
    model.kv.reset()
    with_adapter = model.generate(prompt, a.max_new_tokens)
    lora.merge_lora()                               # fold A·B into the base weights for serving
    model.kv.reset()
    merged = model.generate(prompt, a.max_new_tokens)
    lora.unmerge_lora()
    if dist.get_rank() == 0:
        print(f"{n_adapters} LoRA adapters (rank {a.lora_rank}); merged == unmerged adapter output: {bool(torch.equal(with_adapter, merged))}; "
              f"differs from base model: {not bool(torch.equal(with_adapter, base))}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
