#!/usr/bin/env python
"""``InferenceRunner`` — the harness shared by the inference examples (role of the reference's ``examples/inference/runner.py``
``InferenceRunner``: trace / load / generate / check_accuracy / benchmark, with per-submodule latency collectors).

Sub-class it and provide ``build_model()``; everything else (bucketed "compile" = CUDA-graph capture through ``ModelBuilder``,
generation loop, golden comparison against an eager fp32 copy, latency report in the reference's JSON format) is generic.

  python examples/inference/runner.py --model tiny --check_accuracy --benchmark
"""
import argparse
import json
import os
import sys
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples", "training"))

from neuronx_distributed_b200.inference.autobucketing import generate_buckets  # noqa: E402
from neuronx_distributed_b200.inference.benchmark import Benchmark, LatencyCollector, generate_report  # noqa: E402
from neuronx_distributed_b200.trace.model_builder import ModelBuilder  # noqa: E402
from neuronx_distributed_b200.models.llama import LlamaConfig, llama2_13b_config, llama2_7b_config  # noqa: E402
from neuronx_distributed_b200.models.llama_inference import LlamaForInference  # noqa: E402
from neuronx_distributed_b200.parallel_layers import parallel_state as ps  # noqa: E402
from training_utils import init_distributed  # noqa: E402

CONTEXT_ENCODING_MODEL = "context_encoding_model"
TOKEN_GENERATION_MODEL = "token_generation_model"


class InferenceRunner:
    def __init__(self, tp_degree: int = 1, batch_size: int = 1, max_prompt_length: int = 128, sequence_length: int = 256,
                 use_cuda_graphs: Optional[bool] = None):
        self.dev = init_distributed()
        ps.initialize_model_parallel(tensor_model_parallel_size=tp_degree)
        self.tp_degree, self.batch_size = tp_degree, batch_size
        self.max_prompt_length, self.sequence_length = max_prompt_length, sequence_length
        self.use_cuda_graphs = (self.dev.type == "cuda") if use_cuda_graphs is None else use_cuda_graphs
        self.model = None
        self.nxd_model = None

    # ---- to be provided by the concrete runner ---------------------------------------------------------------------
    def build_model(self):
        raise NotImplementedError

    # ---- generic ---------------------------------------------------------------------------------------------------
    def trace(self):
        """"Compile": capture one program per (sub-model, bucket) with persistent KV cache and return the routed NxDModel."""
        self.model = self.build_model().eval()
        B, P, dev = self.batch_size, self.max_prompt_length, self.dev
        mb = ModelBuilder(tp_degree=self.tp_degree, use_cuda_graphs=self.use_cuda_graphs)
        for bucket in generate_buckets(min(128, P), P):
            mb.add(CONTEXT_ENCODING_MODEL, self.model,
                   [(torch.zeros(B, bucket, dtype=torch.long, device=dev), torch.full((B,), bucket - 1, dtype=torch.long, device=dev))],
                   step_fn=lambda m, i, l: m.context_encoding(i, l))
        mb.add(TOKEN_GENERATION_MODEL, self.model,
               [(torch.zeros(B, 1, dtype=torch.long, device=dev), torch.full((B,), P, dtype=torch.long, device=dev))],
               step_fn=lambda m, i, pos: m.token_generation(i, pos))
        self.nxd_model = mb.trace()
        return self.nxd_model

    def generate(self, prompt: torch.Tensor, max_new_tokens: int) -> torch.Tensor:
        B, P = prompt.shape
        tok = self.nxd_model(prompt, torch.full((B,), P - 1, dtype=torch.long, device=prompt.device)).clone()
        pos = torch.full((B,), P, dtype=torch.long, device=prompt.device)
        out = [tok]
        for _ in range(max_new_tokens - 1):
            tok = self.nxd_model(tok.view(B, 1), pos).clone()
            pos = pos + 1
            out.append(tok)
        return torch.stack(out, 1)

    def check_accuracy(self, prompt: torch.Tensor, max_new_tokens: int = 8) -> bool:
        """Greedy tokens of the bucketed / graph-captured programs vs the same module run eagerly (the reference compares
        Neuron output with a CPU golden, runner.py check_accuracy)."""
        got = self.generate(prompt, max_new_tokens)
        self.model.kv.reset()
        want = self.model.generate(prompt, max_new_tokens)
        ok = bool(torch.equal(got, want))
        if dist.get_rank() == 0:
            print(f"check_accuracy: {'PASS' if ok else 'FAIL'} ({got[0].tolist()} vs {want[0].tolist()})", flush=True)
        return ok

    def benchmark(self, prompt: torch.Tensor, num_runs: int = 20) -> Dict[str, Dict[str, float]]:
        new_tokens = self.sequence_length - prompt.shape[1]
        collectors: Dict[str, LatencyCollector] = {}
        handles: List = []
        for name, mod in (("mlp", self.model.lm.model.layers[0].mlp), ("attention_qkv", self.model.lm.model.layers[0].self_attn.qkv_proj)) \
                if hasattr(self.model.lm, "model") else ():
            c = collectors[name] = LatencyCollector()
            handles += [mod.register_forward_pre_hook(c.pre_hook), mod.register_forward_hook(c.hook)]
        lat = Benchmark(lambda: self.generate(prompt, new_tokens), num_runs=num_runs).run()
        for h in handles:
            h.remove()
        report = {"e2e_model": generate_report(lat, self.sequence_length, self.batch_size)}
        for name, c in collectors.items():
            if c.latencies_s():
                report[name] = {"latency_ms_p50": c.percentile(50) * 1e3, "latency_ms_p99": c.percentile(99) * 1e3}
        if dist.get_rank() == 0:
            print(json.dumps(report), flush=True)
        return report


class LlamaRunner(InferenceRunner):
    def __init__(self, model: str = "tiny", num_layers: int = -1, **kw):
        super().__init__(**kw)
        self.model_name, self.num_layers = model, num_layers

    def build_model(self):
        dtype = torch.bfloat16 if self.dev.type == "cuda" else torch.float32
        kw = dict(dtype=dtype, device=self.dev, max_position_embeddings=self.sequence_length)
        cfg = {"7b": llama2_7b_config, "13b": llama2_13b_config}.get(self.model_name, lambda **k: LlamaConfig(
            vocab_size=4096, hidden_size=256, intermediate_size=704, num_hidden_layers=2, num_attention_heads=8, **k))(**kw)
        if self.num_layers > 0:
            cfg.num_hidden_layers = self.num_layers
        torch.manual_seed(0)
        return LlamaForInference(cfg, batch_size=self.batch_size, max_seq_len=self.sequence_length)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--model", default="tiny", choices=["tiny", "7b", "13b"])
    p.add_argument("--tp_degree", type=int, default=1)
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--max_prompt_length", type=int, default=32)
    p.add_argument("--sequence_length", type=int, default=64)
    p.add_argument("--num_layers", type=int, default=-1)
    p.add_argument("--check_accuracy", action="store_true")
    p.add_argument("--benchmark", action="store_true")
    p.add_argument("--num_runs", type=int, default=5)
    a = p.parse_args()
    r = LlamaRunner(model=a.model, num_layers=a.num_layers, tp_degree=a.tp_degree, batch_size=a.batch_size,
                    max_prompt_length=a.max_prompt_length, sequence_length=a.sequence_length)
    r.trace()
    prompt = torch.randint(0, r.model.cfg.vocab_size, (a.batch_size, a.max_prompt_length), device=r.dev)
    if a.check_accuracy:
        assert r.check_accuracy(prompt)
    if a.benchmark:
        r.benchmark(prompt, a.num_runs)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
