#!/usr/bin/env python
"""Llama inference with LoRA — counterpart of the reference's ``examples/inference/run_llama_lora.py``.

* default: ONE adapter on the attention / MLP output projections, served merged and un-merged;
* ``--multi_lora N``: N adapters resident at once (``modules.lora.serving``), a batch whose requests use different adapters
  (and one request the base model) is decoded in one pass.

  python examples/inference/run_llama_lora.py --lora_rank 8
  python examples/inference/run_llama_lora.py --multi_lora 3
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
    p.add_argument("--multi_lora", type=int, default=0, help="number of adapters served concurrently (0 = single merged adapter)")
    a = p.parse_args()
    dev = init_distributed()
    ps.initialize_model_parallel(tensor_model_parallel_size=a.tp_degree)
    dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
    L = a.prompt_length + a.max_new_tokens
    cfg = LlamaConfig(vocab_size=4096, hidden_size=256, intermediate_size=704, num_hidden_layers=2, num_attention_heads=8,
                      dtype=dtype, device=dev, max_position_embeddings=L)
    torch.manual_seed(0)
    if a.multi_lora > 0:
        return multi_lora(a, cfg, dev, L)
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


def multi_lora(a, cfg, dev, L):
    from neuronx_distributed_b200.modules.lora import LoraServingConfig, LoraServingModel

    n, B = a.multi_lora, a.multi_lora + 1
    model = LlamaForInference(cfg, batch_size=B, max_seq_len=L).eval()
    scfg = LoraServingConfig(max_loras=n, max_lora_rank=a.lora_rank, target_modules=["o_proj", "down_proj"])
    model.lm = LoraServingModel(model.lm, scfg)
    g = torch.Generator().manual_seed(1)
    for slot in range(n):                           # stand-ins for trained adapters: full (un-sharded) HF-PEFT style tensors
        sd = {}
        for name, layer in model.lm.lora_layers().items():
            base = layer.base_layer
            sd[f"base_model.model.{name}.lora_A.weight"] = torch.randn(a.lora_rank, base.input_size, generator=g) * 0.05
            sd[f"base_model.model.{name}.lora_B.weight"] = torch.randn(base.output_size, a.lora_rank, generator=g) * 0.05
        model.lm.load_adapter(slot, {"state_dict": sd, "lora_config": {"lora_alpha": 2 * a.lora_rank, "lora_rank": a.lora_rank}})
    prompt = torch.randint(0, cfg.vocab_size, (1, a.prompt_length), device=dev).expand(B, -1).contiguous()
    ids = torch.tensor(list(range(n)) + [-1], device=dev)           # request i uses adapter i, the last one the base model
    model.lm.set_adapter_ids(ids)
    mixed = model.generate(prompt, a.max_new_tokens)
    model.lm.set_adapter_ids(None)
    model.kv.reset()
    base = model.generate(prompt, a.max_new_tokens)
    if dist.get_rank() == 0:
        same_as_base = [bool(torch.equal(mixed[i], base[i])) for i in range(B)]
        print(f"{n} adapters resident, batch of {B} requests decoded together; request equals base-model output: {same_as_base} "
              f"(expected {[False] * n + [True]})", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
