#!/usr/bin/env python
"""Capture intermediate tensors of a served model — counterpart of the reference's
``examples/inference/tensor_capture/tensor_capture_example.py``.

Module outputs (and inputs) of the listed layers plus tensors registered manually inside model code are collected for every
bucket program; the same program runs with capture on or off.  Here: compare layer outputs of the prefill between the fp32
and the bf16 model to find where precision is lost.

  python examples/inference/tensor_capture_example.py --modules model.layers.0.mlp model.norm
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples", "training"))

from neuronx_distributed_b200.models.llama import LlamaConfig  # noqa: E402
from neuronx_distributed_b200.models.llama_inference import LlamaForInference  # noqa: E402
from neuronx_distributed_b200.parallel_layers import parallel_state as ps  # noqa: E402
from neuronx_distributed_b200.utils.tensor_capture import (disable_tensor_capture, enable_tensor_capture,  # noqa: E402
                                                           get_available_modules, get_captured_tensors_dict)
from training_utils import init_distributed  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--modules", nargs="*", default=None, help="module names under the LM (default: every attention output projection, MLP and the final norm)")
    p.add_argument("--prompt_length", type=int, default=32)
    p.add_argument("--capture_inputs", action="store_true")
    a = p.parse_args()
    dev = init_distributed()
    ps.initialize_model_parallel(tensor_model_parallel_size=dist.get_world_size())
    results = {}
    for dtype in (torch.float32, torch.bfloat16):
        cfg = LlamaConfig(vocab_size=4096, hidden_size=256, intermediate_size=704, num_hidden_layers=4, num_attention_heads=8,
                          dtype=dtype, device=dev, max_position_embeddings=a.prompt_length + 8)
        torch.manual_seed(0)
        model = LlamaForInference(cfg, batch_size=1, max_seq_len=a.prompt_length + 8).eval()
        # the serving body calls the blocks' sub-modules directly (attention reads the KV cache in between), so hook those
        names = a.modules or [n for n in get_available_modules(model.lm) if n.endswith((".mlp", ".self_attn.o_proj")) or n == "model.norm"]
        enable_tensor_capture(model.lm, names, max_tensors=4, capture_inputs=a.capture_inputs)
        prompt = torch.randint(0, cfg.vocab_size, (1, a.prompt_length), device=dev, generator=torch.Generator(device=dev).manual_seed(1))
        model.context_encoding(prompt, torch.tensor([a.prompt_length - 1], device=dev))
        results[dtype] = {k: v.float().cpu() for k, v in get_captured_tensors_dict().items()}
        disable_tensor_capture(model.lm)
    if dist.get_rank() == 0:
        print(f"{'tensor':40s} {'shape':>18s}  max|fp32 - bf16| / max|fp32|")
        for k, ref in results[torch.float32].items():
            got = results[torch.bfloat16].get(k)
            if got is not None and got.shape == ref.shape:
                print(f"{k:40s} {str(tuple(ref.shape)):>18s}  {float((ref - got).abs().max() / ref.abs().max().clamp(min=1e-9)):.3e}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
