"""BASELINE config 5 through ``bench.py --mode decode``: Llama-2-13B inference latency at TP = N, batch 1, sequence 2048
(prompt 1024 + 1024 generated tokens), random-init weights, synthetic prompt.

Metric (reference ``examples/inference/modules/benchmark.py:9-73``, ``runner.py:521-602``): latency percentiles of the full
``generate`` call over ``num_runs`` runs after one warm-up, plus the per-submodule collectors (context encoding, token
generation).  Here every number is device time (CUDA events on the launching stream), max over ranks:

  value               p50 of the token-generation step (ms per token) under CUDA graphs, cache positions 1024..2047
  generate            p50 / p90 / p100 of the whole generate (prefill + 1023 decode steps), and tokens/s as the reference
                      computes it (runs x max_length x batch / total time)
  e2e                 generate timed end to end: prompt from pinned host memory each run, generated ids read back to the host

The decode step runs the framework's own path: gemv kernels, RoPE, split-KV decode attention, and — for TP > 1 — the fused
GEMV + in-switch all-reduce kernel (csrc/nvls_coll.cu) for o_proj / down_proj; all captured in one CUDA graph per bucket.
"""
from __future__ import annotations

import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _pct(xs, p):
    xs = sorted(xs)
    if not xs:
        return 0.0
    k = (len(xs) - 1) * p / 100.0
    lo, hi = int(k), min(int(k) + 1, len(xs) - 1)
    return xs[lo] + (xs[hi] - xs[lo]) * (k - lo)


def main(args) -> int:
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29535")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from neuronx_distributed_b200 import ops
    from neuronx_distributed_b200.trace.model_builder import ModelBuilder
    from neuronx_distributed_b200.models.llama import llama2_13b_config
    from neuronx_distributed_b200.models.llama_inference import LlamaForInference
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    B, S, P = 1, args.seq if args.seq != 4096 else 2048, 1024
    cfg = llama2_13b_config(dtype=torch.bfloat16, device=dev, max_position_embeddings=S)
    if args.layers != 32:
        cfg.num_hidden_layers = args.layers                      # debug only; the JSON line says so
    torch.manual_seed(0); torch.cuda.manual_seed(0)
    model = LlamaForInference(cfg, batch_size=B, max_seq_len=S).eval()
    mb = ModelBuilder(tp_degree=world, use_cuda_graphs=True)
    mb.add("context_encoding_model", model,
           [(torch.zeros(B, P, dtype=torch.long, device=dev), torch.full((B,), P - 1, dtype=torch.long, device=dev))],
           step_fn=lambda m, i, l: m.context_encoding(i, l))
    mb.add("token_generation_model", model,
           [(torch.zeros(B, 1, dtype=torch.long, device=dev), torch.full((B,), P, dtype=torch.long, device=dev))],
           step_fn=lambda m, i, pos: m.token_generation(i, pos))
    nxd_model = mb.trace()
    gen = torch.Generator().manual_seed(3)
    host_prompt = torch.randint(0, cfg.vocab_size, (B, P), generator=gen).pin_memory()
    prompt = host_prompt.to(dev)
    last = torch.full((B,), P - 1, dtype=torch.long, device=dev)
    new_tokens = S - P

    def generate(prompt_dev, collect=None):
        ev = []
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
        tok = nxd_model(prompt_dev, last).clone()
        e1 = torch.cuda.Event(enable_timing=True); e1.record()
        pos = torch.full((B,), P, dtype=torch.long, device=dev)
        out = [tok]
        for _ in range(new_tokens - 1):
            if collect is not None:
                a = torch.cuda.Event(enable_timing=True); a.record()
            tok = nxd_model(tok.view(B, 1), pos).clone()
            if collect is not None:
                b = torch.cuda.Event(enable_timing=True); b.record()
                ev.append((a, b))
            pos = pos + 1
            out.append(tok)
        e2 = torch.cuda.Event(enable_timing=True); e2.record()
        if collect is not None:
            collect.append((e0, e1, e2, ev))
        return torch.stack(out, 1)

    def sync():
        dist.barrier(); torch.cuda.synchronize()

    for _ in range(max(1, args.warmup // 3 or 1)):
        toks = generate(prompt)
    sync()
    ops._ext.reset_launches()
    runs = []
    n_runs = max(1, args.steps)
    for _ in range(n_runs):
        sync()
        generate(prompt, collect=runs)
    sync()
    launches = ops._ext.launches()
    gen_ms, pre_ms, tok_ms = [], [], []
    for e0, e1, e2, ev in runs:
        gen_ms.append(e0.elapsed_time(e2)); pre_ms.append(e0.elapsed_time(e1))
        tok_ms.extend(a.elapsed_time(b) for a, b in ev)

    def allmax(vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    gen_ms, pre_ms, tok_ms = allmax(gen_ms), allmax(pre_ms), allmax(tok_ms)
    # end to end: pinned host prompt → device every run, generated ids → host
    e2e_ms, h2d, d2h = [], 0, 0
    for _ in range(min(3, n_runs)):
        sync()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        pd = host_prompt.to(dev, non_blocking=True); h2d += host_prompt.numel() * host_prompt.element_size()
        out = generate(pd).cpu(); d2h += out.numel() * out.element_size()
        b.record(); torch.cuda.synchronize()
        e2e_ms.append(a.elapsed_time(b))
    e2e_ms = allmax(e2e_ms)
    if rank == 0:
        p50 = _pct(tok_ms, 50)
        total_s = sum(gen_ms) / 1e3
        print(json.dumps({
            "metric": "Llama-2-13B inference latency, batch 1, seq 2048 (device-timed, max over ranks): token-generation step p50",
            "value": p50, "unit": "ms/token", "n_gpus": world, "steps": n_runs, "warmup": args.warmup, "ms_per_step": p50,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "impl": "ours",
            "config": {"model": "llama2-13b" if args.layers == 32 else f"llama2-13b-{args.layers}L(debug)", "global_batch": B,
                       "seq_len": S, "prompt_len": P, "new_tokens": new_tokens, "parallelism": f"tp{world}",
                       "cuda_graphs": True, "l2": "weights (26 GB / tp) >> L2: every step streams them from HBM"},
            "token_generation_ms": {"p50": p50, "p90": _pct(tok_ms, 90), "p99": _pct(tok_ms, 99), "avg": statistics.fmean(tok_ms)},
            "context_encoding_ms": {"p50": _pct(pre_ms, 50), "p100": max(pre_ms)},
            "generate_ms": {"p50": _pct(gen_ms, 50), "p90": _pct(gen_ms, 90), "p100": max(gen_ms), "avg": statistics.fmean(gen_ms),
                            "throughput_tokens_per_s_reference_formula": n_runs * S * B / total_s},
            "e2e": {"value": _pct(e2e_ms, 50), "unit": "ms per generate (prompt H2D from pinned memory + ids D2H)",
                    "h2d_bytes_per_step": h2d // max(1, len(e2e_ms)), "d2h_bytes_per_step": d2h // max(1, len(e2e_ms))},
            "gpu_launches": launches, "sample_tokens": toks[0, :6].tolist()}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0
