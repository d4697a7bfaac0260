#!/usr/bin/env python
"""GPT-NeoX pre-training, TP × PP 1F1B (BASELINE config 4: NeoX-20B, TP=4 × PP=2 on 8 GPUs) — counterpart of the reference's
``examples/training/tp_dp_gpt_neox_hf_pretrain``.

  torchrun --nproc-per-node 8 examples/training/gpt_neox/tp_pp_gpt_neox_pretrain.py --model 20b --tensor_parallel_size 4 \
      --pipeline_parallel_size 2 --num_microbatches 8
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples", "training"))

import neuronx_distributed_b200 as nxd  # noqa: E402
from neuronx_distributed_b200.models.gpt_neox import GPTNeoXConfig, GPTNeoXForCausalLM, GPTNeoXLayer  # noqa: E402
from neuronx_distributed_b200.parallel_layers import parallel_state as ps  # noqa: E402
from neuronx_distributed_b200.utils.adamw_fp32_optim_params import AdamW_FP32OptimParams  # noqa: E402
from training_utils import (Throughput, add_checkpoint_args, init_distributed, maybe_resume, maybe_save,  # noqa: E402
                            synthetic_batches)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--model", default="tiny", choices=["tiny", "6.9b", "20b"])
    p.add_argument("--tensor_parallel_size", type=int, default=2)
    p.add_argument("--pipeline_parallel_size", type=int, default=2)
    p.add_argument("--num_microbatches", type=int, default=4)
    p.add_argument("--seq_len", type=int, default=2048)
    p.add_argument("--max_steps", type=int, default=10)
    p.add_argument("--num_layers", type=int, default=-1)
    p.add_argument("--lr", type=float, default=1e-4)
    add_checkpoint_args(p)
    a = p.parse_args()
    dev = init_distributed()
    sp = a.tensor_parallel_size > 1
    pcfg = {"transformer_layer_cls": GPTNeoXLayer, "num_microbatches": a.num_microbatches, "output_loss_value_spec": (True, False),
            "input_names": ["input_ids", "labels"], "broadcast_and_average_loss": True} if a.pipeline_parallel_size > 1 else None
    cfg = nxd.neuronx_distributed_config(tensor_parallel_size=a.tensor_parallel_size, pipeline_parallel_size=a.pipeline_parallel_size,
                                         sequence_parallel=sp, pipeline_config=pcfg,
                                         optimizer_config={"zero_one_enabled": True, "grad_clipping": True, "max_grad_norm": 1.0})
    dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
    shapes = {"20b": dict(hidden_size=6144, num_hidden_layers=44, num_attention_heads=64, intermediate_size=24576, vocab_size=50432),
              "6.9b": dict(hidden_size=4096, num_hidden_layers=32, num_attention_heads=32, intermediate_size=16384, vocab_size=50432),
              "tiny": dict(hidden_size=256, num_hidden_layers=4, num_attention_heads=8, intermediate_size=1024, vocab_size=4096)}[a.model]
    mcfg = GPTNeoXConfig(sequence_parallel_enabled=sp, dtype=dtype, max_position_embeddings=a.seq_len,
                         device=dev if a.pipeline_parallel_size == 1 else None, **shapes)
    if a.num_layers > 0:
        mcfg.num_hidden_layers = a.num_layers

    def model_fn():
        torch.manual_seed(1234)
        return GPTNeoXForCausalLM(mcfg)

    model = nxd.initialize_parallel_model(cfg, model_fn)
    opt = nxd.initialize_parallel_optimizer(cfg, AdamW_FP32OptimParams, model.parameters(), lr=a.lr)
    start = maybe_resume(a, nxd, model, opt)
    data = synthetic_batches(mcfg.vocab_size, a.num_microbatches, a.seq_len, 1 + ps.get_data_parallel_rank(), dev)
    thr = Throughput(a.num_microbatches, ps.get_data_parallel_size(), 1)
    for step in range(start, a.max_steps):
        if a.steps_this_run >= 0 and step - start >= a.steps_this_run:
            break
        opt.zero_grad()
        loss = model.run_train(**next(data))
        opt.step()
        tp = thr.get_throughput()
        if dist.get_rank() == 0:
            print(f"step {step + 1} loss {float(loss):.4f} throughput {tp:.2f} seq/s ({tp * a.seq_len:.0f} tok/s)", flush=True)
        maybe_save(a, nxd, model, opt, step + 1)
    nxd.finalize_checkpoint()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
