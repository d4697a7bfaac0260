#!/usr/bin/env python
"""DBRX pre-training (fine-grained MoE: 16 experts, top-4, clip_qkv, LayerNorm), TP × EP with ZeRO-1 — counterpart of the
reference's ``examples/training/dbrx``.

  torchrun --nproc-per-node 8 examples/training/dbrx/tp_ep_dbrx_pretrain.py --model dbrx --tensor_parallel_size 4 --expert_parallel_size 2
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
from neuronx_distributed_b200.models.mixtral import DbrxConfig, MixtralForCausalLM  # noqa: E402
from neuronx_distributed_b200.parallel_layers import parallel_state as ps  # noqa: E402
from neuronx_distributed_b200.utils.adamw_fp32_optim_params import AdamW_FP32OptimParams  # noqa: E402
from training_utils import Throughput, init_distributed, synthetic_batches  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--tensor_parallel_size", type=int, default=1)
    p.add_argument("--expert_parallel_size", type=int, default=1)
    p.add_argument("--model", default="tiny", choices=["tiny", "dbrx"])
    p.add_argument("--capacity_factor", type=float, default=None, help="None = dropless (blockwise / all-experts)")
    p.add_argument("--seq_len", type=int, default=512)
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--max_steps", type=int, default=10)
    a = p.parse_args()
    dev = init_distributed()
    sp = a.tensor_parallel_size > 1
    cfg = nxd.neuronx_distributed_config(tensor_parallel_size=a.tensor_parallel_size, expert_parallel_size=a.expert_parallel_size,
                                         sequence_parallel=sp,
                                         optimizer_config={"zero_one_enabled": True, "grad_clipping": True, "max_grad_norm": 1.0})
    dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
    shapes = {} if a.model == "dbrx" else dict(vocab_size=4096, hidden_size=256, intermediate_size=384, num_hidden_layers=2,
                                               num_attention_heads=8, num_key_value_heads=4, num_local_experts=8, num_experts_per_tok=4)
    mcfg = DbrxConfig(sequence_parallel_enabled=sp, dtype=dtype, device=dev, max_position_embeddings=a.seq_len,
                      capacity_factor=a.capacity_factor, **shapes)

    def model_fn():
        torch.manual_seed(1234)
        return MixtralForCausalLM(mcfg)

    model = nxd.initialize_parallel_model(cfg, model_fn)
    opt = nxd.initialize_parallel_optimizer(cfg, AdamW_FP32OptimParams, model.parameters(), lr=1e-4)
    data = synthetic_batches(mcfg.vocab_size, a.batch_size, a.seq_len, 1 + ps.get_data_parallel_rank(), dev)
    thr = Throughput(a.batch_size, ps.get_data_parallel_size(), 1)
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
