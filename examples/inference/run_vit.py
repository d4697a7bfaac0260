#!/usr/bin/env python
"""ViT image-classification inference with tensor parallelism — counterpart of the reference's ``examples/inference/run_vit.py``.

  torchrun --nproc-per-node 2 examples/inference/run_vit.py --tp_degree 2 --batch_size 8 --benchmark
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples", "training"))

from neuronx_distributed_b200.models.vit import ViTConfig, ViTForImageClassification  # noqa: E402
from neuronx_distributed_b200.parallel_layers import parallel_state as ps  # noqa: E402
from neuronx_distributed_b200.utils.profiling import device_timer  # noqa: E402
from training_utils import init_distributed  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--model", default="tiny", choices=["tiny", "base", "large"])
    p.add_argument("--tp_degree", type=int, default=1)
    p.add_argument("--batch_size", type=int, default=4)
    p.add_argument("--benchmark", action="store_true")
    p.add_argument("--num_runs", type=int, default=20)
    a = p.parse_args()
    dev = init_distributed()
    ps.initialize_model_parallel(tensor_model_parallel_size=a.tp_degree)
    dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
    shapes = {"large": dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096),
              "base": dict(), "tiny": dict(image_size=32, patch_size=8, hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                                          intermediate_size=128, num_labels=10)}[a.model]
    cfg = ViTConfig(dtype=dtype, device=dev, **shapes)
    torch.manual_seed(0)
    model = ViTForImageClassification(cfg).eval()
    x = torch.randn(a.batch_size, cfg.num_channels, cfg.image_size, cfg.image_size, device=dev, dtype=dtype)
    with torch.no_grad():
        logits = model(x)
        if dist.get_rank() == 0:
            print("top-1 classes:", logits.argmax(-1).tolist())
        if a.benchmark:
            times = []
            for _ in range(a.num_runs):
                with device_timer() as t:
                    model(x)
                times.append(t.ms)
            times.sort()
            if dist.get_rank() == 0:
                print(f"latency p50 {times[len(times) // 2]:.3f} ms  p99 {times[-1]:.3f} ms  ({a.batch_size / times[len(times) // 2] * 1e3:.0f} img/s)")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
