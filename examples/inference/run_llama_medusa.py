#!/usr/bin/env python
"""Medusa decoding: extra heads propose a candidate tree, the base model verifies it in one forward with a tree attention mask —
counterpart of the Medusa path of the reference's inference examples (``utils/medusa_utils.py`` buffers).

  python examples/inference/run_llama_medusa.py --num_medusa_heads 3
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples", "training"))

from neuronx_distributed_b200.models.llama import LlamaConfig, llama2_7b_config  # noqa: E402
from neuronx_distributed_b200.models.llama_inference import LlamaForInference  # noqa: E402
from neuronx_distributed_b200.parallel_layers import parallel_state as ps  # noqa: E402
from neuronx_distributed_b200.utils.medusa import MedusaHeads, medusa_generate  # noqa: E402
from training_utils import init_distributed  # noqa: E402

# a small version of the mc_sim_7b_63 tree of the Medusa paper: (choice of head 0, choice of head 1, …) per path
MEDUSA_CHOICES = [[0], [1], [2], [0, 0], [0, 1], [1, 0], [0, 0, 0], [0, 0, 1], [0, 1, 0]]


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--model", default="tiny", choices=["tiny", "7b"])
    p.add_argument("--tp_degree", type=int, default=1)
    p.add_argument("--num_medusa_heads", type=int, default=3)
    p.add_argument("--topk", type=int, default=10)
    p.add_argument("--prompt_length", type=int, default=32)
    p.add_argument("--max_new_tokens", type=int, default=32)
    a = p.parse_args()
    dev = init_distributed()
    ps.initialize_model_parallel(tensor_model_parallel_size=a.tp_degree)
    dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
    L = a.prompt_length + a.max_new_tokens + len(MEDUSA_CHOICES) + 2
    kw = dict(dtype=dtype, device=dev, max_position_embeddings=L)
    cfg = llama2_7b_config(**kw) if a.model == "7b" else LlamaConfig(vocab_size=4096, hidden_size=256, intermediate_size=704,
                                                                    num_hidden_layers=4, num_attention_heads=8, **kw)
    torch.manual_seed(0)
    model = LlamaForInference(cfg, batch_size=1, max_seq_len=L).eval()
    heads = MedusaHeads(cfg.hidden_size, cfg.vocab_size, a.num_medusa_heads, dtype=dtype, device=dev).eval()
    for proj in heads.proj:        # untrained stand-in: start every head from the LM head (head k then guesses "same token again")
        proj.weight.data.copy_(model.lm.lm_head.weight.data)
    prompt = torch.randint(0, cfg.vocab_size, (1, a.prompt_length), device=dev)
    ref = model.generate(prompt, a.max_new_tokens)
    model.kv.reset()
    out, acc = medusa_generate(model, heads, prompt, a.max_new_tokens, MEDUSA_CHOICES, topk=a.topk)
    if dist.get_rank() == 0:
        print(f"identical to greedy: {bool(torch.equal(out, ref))}; accepted tree tokens per verification: {acc:.2f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
