#!/usr/bin/env python
"""Mixtral / DBRX (sparse MoE) inference: prefill + KV-cache decode with tensor parallelism — counterpart of the reference's
``examples/inference/run_mixtral.py`` and ``run_dbrx.py``.

  torchrun --nproc-per-node 2 examples/inference/run_mixtral.py --family mixtral --tp_degree 2
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples", "training"))

from neuronx_distributed_b200.models.llama_inference import LlamaForInference  # noqa: E402
from neuronx_distributed_b200.models.mixtral import DbrxConfig, MixtralConfig, MixtralForCausalLM  # noqa: E402
from neuronx_distributed_b200.parallel_layers import parallel_state as ps  # noqa: E402
from neuronx_distributed_b200.utils.profiling import device_timer  # noqa: E402
from training_utils import init_distributed  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--family", default="mixtral", choices=["mixtral", "dbrx"])
    p.add_argument("--size", default="tiny", choices=["tiny", "full"])
    p.add_argument("--tp_degree", type=int, default=1)
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--prompt_length", type=int, default=32)
    p.add_argument("--max_new_tokens", type=int, default=16)
    a = p.parse_args()
    dev = init_distributed()
    ps.initialize_model_parallel(tensor_model_parallel_size=a.tp_degree)
    dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
    L = a.prompt_length + a.max_new_tokens
    kw = dict(dtype=dtype, device=dev, max_position_embeddings=L)
    tiny = dict(vocab_size=4096, hidden_size=256, intermediate_size=384, num_hidden_layers=2, num_attention_heads=8,
                num_key_value_heads=4, num_local_experts=4, num_experts_per_tok=2)
    if a.family == "dbrx":
        cfg = DbrxConfig(**kw) if a.size == "full" else DbrxConfig(**{**tiny, "num_local_experts": 8, "num_experts_per_tok": 4}, **kw)
    else:
        cfg = (MixtralConfig(vocab_size=32000, hidden_size=4096, num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=8, **kw)
               if a.size == "full" else MixtralConfig(**tiny, **kw))
    torch.manual_seed(0)
    model = LlamaForInference(cfg, batch_size=a.batch_size, max_seq_len=L, lm_cls=MixtralForCausalLM).eval()
    prompt = torch.randint(0, cfg.vocab_size, (a.batch_size, a.prompt_length), device=dev)
    with device_timer() as t:
        out = model.generate(prompt, a.max_new_tokens)
    if dist.get_rank() == 0:
        print(f"{a.family}: generated {tuple(out.shape)} in {t.ms:.1f} ms; first tokens {out[0, :8].tolist()}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
