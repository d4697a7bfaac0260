#!/usr/bin/env python
"""Llama pre-training with TP × PP (1F1B / interleaved) — counterpart of the reference's
``examples/training/llama/tp_pp_llama_hf_pretrain/run_llama_nxd.py``.

  torchrun --nproc-per-node 8 examples/training/llama/tp_pp_llama_pretrain.py --model 13b --tensor_parallel_size 4 \
      --pipeline_parallel_size 2 --num_microbatches 8 --seq_len 4096
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
from neuronx_distributed_b200.models.llama import (LlamaConfig, LlamaDecoderLayer, LlamaForCausalLM, llama2_7b_config,  # noqa: E402
                                                   llama2_13b_config, llama2_70b_config)
from neuronx_distributed_b200.parallel_layers import parallel_state as ps  # noqa: E402
from neuronx_distributed_b200.utils.adamw_fp32_optim_params import AdamW_FP32OptimParams  # noqa: E402
from training_utils import Throughput, init_distributed, synthetic_batches  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--model", default="tiny", choices=["tiny", "7b", "13b", "70b"])
    p.add_argument("--pretrained_hf", default=None, help="HF Llama directory (config.json + safetensors): continue pre-training / fine-tune from it")
    p.add_argument("--tensor_parallel_size", type=int, default=2)
    p.add_argument("--pipeline_parallel_size", type=int, default=2)
    p.add_argument("--virtual_pipeline_size", type=int, default=1)
    p.add_argument("--num_microbatches", type=int, default=4)
    p.add_argument("--use_sequence_parallel", type=int, default=1)
    p.add_argument("--use_zero_1", type=int, default=1)
    p.add_argument("--seq_len", type=int, default=512)
    p.add_argument("--max_steps", type=int, default=10)
    p.add_argument("--lr", type=float, default=3e-4)
    a = p.parse_args()
    dev = init_distributed()
    sp = bool(a.use_sequence_parallel) and a.tensor_parallel_size > 1
    cfg = nxd.neuronx_distributed_config(
        tensor_parallel_size=a.tensor_parallel_size, pipeline_parallel_size=a.pipeline_parallel_size, sequence_parallel=sp,
        pipeline_config={"transformer_layer_cls": LlamaDecoderLayer, "num_microbatches": a.num_microbatches,
                         "virtual_pipeline_size": a.virtual_pipeline_size, "output_loss_value_spec": (True, False),
                         "input_names": ["input_ids", "labels"], "broadcast_and_average_loss": True},
        optimizer_config={"zero_one_enabled": bool(a.use_zero_1), "grad_clipping": True, "max_grad_norm": 1.0},
    )
    dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
    kw = dict(sequence_parallel_enabled=sp, dtype=dtype, max_position_embeddings=a.seq_len)
    mcfg = {"7b": llama2_7b_config, "13b": llama2_13b_config, "70b": llama2_70b_config}.get(a.model, lambda **k: LlamaConfig(
        vocab_size=4096, hidden_size=512, intermediate_size=1408, num_hidden_layers=8, num_attention_heads=8, **k))(**kw)
    if a.pretrained_hf:                                          # architecture from the HF config, weights loaded after sharding
        from neuronx_distributed_b200.models import hf_compat

        mcfg = hf_compat.config_from_hf(a.pretrained_hf, **kw)

    def model_fn():
        torch.manual_seed(1234)
        return LlamaForCausalLM(mcfg)

    model = nxd.initialize_parallel_model(cfg, model_fn)
    if a.pretrained_hf:
        hf_compat.load_hf_checkpoint(model, a.pretrained_hf)     # every rank keeps its (tp, pp) shard only
    opt = nxd.initialize_parallel_optimizer(cfg, AdamW_FP32OptimParams, model.parameters(), lr=a.lr)
    dp = ps.get_data_parallel_size()
    data = synthetic_batches(mcfg.vocab_size, a.num_microbatches, a.seq_len, 1 + ps.get_data_parallel_rank(), dev)
    thr = Throughput(a.num_microbatches, dp, 1)
    for step in range(a.max_steps):
        opt.zero_grad()
        loss = model.run_train(**next(data))
        opt.step()
        tp = thr.get_throughput()
        if dist.get_rank() == 0:
            print(f"step {step + 1} loss {float(loss):.4f} throughput {tp:.2f} seq/s", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
