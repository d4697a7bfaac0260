"""Reference arm of bench.py: the UNMODIFIED reference (``baseline/_ref``, installed with
``pip install --no-deps --target baseline/_ref``) running Llama-2-7B training through ITS OWN public API —
``neuronx_distributed_config`` → ``initialize_parallel_model`` → ``initialize_parallel_optimizer`` (ZeRO-1, fp32 master
weights + fp32 grad accumulation) with its ``ColumnParallelLinear`` / ``RowParallelLinear`` / ``ParallelEmbedding`` /
``GQAQKVColumnParallelLinear`` / ``RMSNorm`` / ``parallel_cross_entropy`` and its ``AdamW_FP32OptimParams``.

None of this repository's models, kernels or engine is imported here.  The Neuron-only dependencies the reference
imports are satisfied by ``baseline/xla_stubs.py`` (XLA collectives → NCCL, device → CUDA, torch_xla ZeRO → NCCL
re-implementation); GEMMs are cuBLAS via ``torch.matmul``/``einsum`` exactly as the reference issues them, attention is
``F.scaled_dot_product_attention`` standing in for its NKI flash kernel.  The model glue below follows the structure of
the reference example ``examples/training/llama/modeling_llama_nxd.py`` (which itself cannot be imported: it targets
transformers 4.x; this image ships 5.x).
"""
from __future__ import annotations

import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def main(args) -> int:
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(HERE, "_ref"))
    if not os.path.isdir(os.path.join(HERE, "_ref", "neuronx_distributed")):
        raise RuntimeError("baseline/_ref is not installed (pip install --no-deps --target baseline/_ref /root/reference)")
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from torch import nn

    import xla_stubs

    xla_stubs.install()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus
    cpu_debug = not torch.cuda.is_available()          # plumbing check on a CPU box (gloo); never a benchmark
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29534")
    if cpu_debug:
        # NOTE: NXD_CPU_MODE stays unset so the reference takes the same xm.* code path it takes on the GPU
        dev = torch.device("cpu")
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        torch.cuda.set_device(local_rank)
        dev = torch.device("cuda", local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    import neuronx_distributed as nxd
    from neuronx_distributed.modules.qkv_linear import GQAQKVColumnParallelLinear
    from neuronx_distributed.modules.rms_norm import RMSNorm
    from neuronx_distributed.parallel_layers import (ColumnParallelLinear, ParallelEmbedding, RowParallelLinear,
                                                     parallel_cross_entropy)
    from neuronx_distributed.parallel_layers import parallel_state as ps
    from neuronx_distributed.utils.adamw_fp32_optim_params import AdamW_FP32OptimParams

    tp = world
    sp = tp > 1
    H, I, L, NH, V, S = 4096, 11008, args.layers, 32, 32000, args.seq
    if cpu_debug:
        H, I, NH, V = 64, 128, 4, 256
    D = H // NH
    init = lambda w: nn.init.normal_(w, mean=0.0, std=0.02)
    dt = torch.float32 if cpu_debug else torch.bfloat16

    # `device="xla"` literals are rewritten only while the reference sets itself up (parallel_state.py:655 is the one on this
    # path); the timed step runs WITHOUT a TorchFunctionMode.  If a step still trips over an "xla" device the mode is
    # entered for good and the JSON line says so (`stub_device_rewrite`).
    rewrite_scope = "setup-only"
    with xla_stubs.device_rewrite():
        cfg = nxd.neuronx_distributed_config(
            tensor_parallel_size=tp, sequence_parallel=sp,
            optimizer_config={"zero_one_enabled": True, "grad_clipping": True, "max_grad_norm": 1.0},
            mixed_precision_config={"use_master_weights": True, "use_fp32_grad_acc": True, "use_master_weights_in_ckpt": False},
        )

    def rope_tables(seq, dim, device):
        inv = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float32, device=device) / dim))
        fr = torch.outer(torch.arange(seq, dtype=torch.float32, device=device), inv)
        emb = torch.cat([fr, fr], dim=-1)
        return emb.cos()[None, None], emb.sin()[None, None]           # [1,1,S,D]

    def rotate_half(x):
        return torch.cat([-x[..., x.shape[-1] // 2:], x[..., : x.shape[-1] // 2]], dim=-1)

    class Attn(nn.Module):
        def __init__(self):
            super().__init__()
            self.qkv_proj = GQAQKVColumnParallelLinear(H, [NH * D, NH * D], bias=False, gather_output=False, init_method=init,
                                                       sequence_parallel_enabled=sp, kv_size_multiplier=1, fuse_qkv=True, dtype=dt)
            self.o_proj = RowParallelLinear(NH * D, H, bias=False, input_is_parallel=True, init_method=init,
                                            sequence_parallel_enabled=sp, dtype=dt)
            self.nh = NH // tp

        def forward(self, x, cos, sin):                      # x [S(/tp), B, H]
            q, k, v = self.qkv_proj(x)
            Sq, B = q.shape[0], q.shape[1]
            q, k, v = (t.view(Sq, B, self.nh, D).permute(1, 2, 0, 3) for t in (q, k, v))   # [B,h,S,D]
            q = (q * cos + rotate_half(q) * sin).to(q.dtype)
            k = (k * cos + rotate_half(k) * sin).to(k.dtype)
            o = F.scaled_dot_product_attention(q, k, v, is_causal=True)
            return self.o_proj(o.permute(2, 0, 1, 3).reshape(Sq, B, self.nh * D))

    class MLP(nn.Module):
        def __init__(self):
            super().__init__()
            self.gate_up_proj = ColumnParallelLinear(H, 2 * I, stride=2, bias=False, gather_output=False, init_method=init,
                                                     sequence_parallel_enabled=sp, dtype=dt)
            self.down_proj = RowParallelLinear(I, H, bias=False, input_is_parallel=True, init_method=init,
                                               sequence_parallel_enabled=sp, dtype=dt)

        def forward(self, x):
            g, u = self.gate_up_proj(x).chunk(2, dim=-1)
            return self.down_proj(F.silu(g) * u)

    class Layer(nn.Module):
        def __init__(self):
            super().__init__()
            self.input_layernorm = RMSNorm(H, eps=1e-5, sequence_parallel_enabled=sp) if _rms_takes_sp(RMSNorm) else RMSNorm(H, eps=1e-5)
            self.post_attention_layernorm = RMSNorm(H, eps=1e-5, sequence_parallel_enabled=sp) if _rms_takes_sp(RMSNorm) else RMSNorm(H, eps=1e-5)
            self.self_attn, self.mlp = Attn(), MLP()

        def forward(self, x, cos, sin):
            x = x + self.self_attn(self.input_layernorm(x), cos, sin)
            return x + self.mlp(self.post_attention_layernorm(x))

    class Llama(nn.Module):
        def __init__(self):
            super().__init__()
            self.embed_tokens = ParallelEmbedding(V, H, init_method=init, dtype=dt, sequence_parallel_enabled=sp)
            self.layers = nn.ModuleList([Layer() for _ in range(L)])
            self.norm = RMSNorm(H, eps=1e-5, sequence_parallel_enabled=sp) if _rms_takes_sp(RMSNorm) else RMSNorm(H, eps=1e-5)
            self.lm_head = ColumnParallelLinear(H, V, bias=False, gather_output=False, init_method=init,
                                                sequence_parallel_enabled=sp, dtype=dt)

        def forward(self, input_ids, labels):
            x = self.embed_tokens(input_ids)                         # SP: [S/tp,B,H]; else [B,S,H]
            if not sp:
                x = x.transpose(0, 1).contiguous()
            cos, sin = rope_tables(input_ids.shape[1], D, input_ids.device)
            for l in self.layers:
                x = l(x, cos, sin)
            logits = self.lm_head(self.norm(x)).float()              # [S,B,V/tp]
            tgt = labels.transpose(0, 1)
            tgt = torch.cat([tgt[1:], torch.zeros_like(tgt[:1])], 0)
            loss = parallel_cross_entropy(logits, tgt)
            return loss[:-1].mean()

    def model_fn():
        torch.manual_seed(1234)
        return Llama()

    with xla_stubs.device_rewrite():
        model = nxd.initialize_parallel_model(cfg, model_fn)
        opt = nxd.initialize_parallel_optimizer(cfg, AdamW_FP32OptimParams, model.parameters(), lr=3e-4, betas=(0.9, 0.95),
                                                weight_decay=0.1)
    gbs = args.global_batch
    mbs = getattr(args, "micro_batch", 0) or {1: 2, 2: 2, 4: 4, 8: 4}.get(args.gpus, 1)   # same rule as the other arm
    mbs = max(1, min(mbs, gbs))
    while gbs % mbs:
        mbs -= 1
    gen = torch.Generator().manual_seed(7)
    n_host = args.steps + args.warmup + 2
    host_ids = [torch.randint(0, V, (gbs, S), generator=gen) for _ in range(n_host)]
    if not cpu_debug:
        host_ids = [h.pin_memory() for h in host_ids]
    dev_ids = [h.to(dev) for h in host_ids]

    def train_step(ids_dev):
        opt.zero_grad()
        total = None
        for mb in range(0, gbs, mbs):
            ids = ids_dev[mb:mb + mbs]
            loss = model.run_train(ids, ids)
            total = loss if total is None else total + loss
        opt.step()
        return total / (gbs // mbs)

    def sync():
        dist.barrier()
        if not cpu_debug:
            torch.cuda.synchronize()

    for i in range(args.warmup):
        try:
            try:
                train_step(dev_ids[i])
            except RuntimeError as e:
                if "xla" not in str(e).lower() or isinstance(e, torch.cuda.OutOfMemoryError):
                    raise
                xla_stubs.device_rewrite().__enter__()          # an "xla" literal inside the step: rewrite process-wide
                rewrite_scope = "global (an xla device literal is used inside the step)"
                opt.zero_grad()
                train_step(dev_ids[i])
        except torch.cuda.OutOfMemoryError:
            # the un-fused reference keeps more activation memory than the other arm; on ONE GPU (no peers to desynchronise)
            # fall back to micro-batch 1 instead of failing — the JSON line reports the micro-batch actually used
            if world != 1 or mbs == 1:
                raise
            opt.zero_grad()
            import gc
            gc.collect(); torch.cuda.empty_cache()
            mbs = 1
            train_step(dev_ids[i])
    sync()
    if cpu_debug:
        sync(); t0 = time.perf_counter()
        for i in range(args.steps):
            loss = train_step(dev_ids[args.warmup + i])
        sync(); elapsed_ms = (time.perf_counter() - t0) * 1e3
    else:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync(); ev0.record()
        for i in range(args.steps):
            loss = train_step(dev_ids[args.warmup + i])
        ev1.record(); sync()
        elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t)
    value = gbs * S * args.steps / (ms_total / 1e3)
    e2e = None
    if not args.no_e2e:
        sync(); h2d = d2h = 0
        if cpu_debug:
            t0 = time.perf_counter()
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
        for i in range(args.steps):
            h = host_ids[i % len(host_ids)]
            ids = h.to(dev, non_blocking=True); h2d += h.numel() * h.element_size()
            lv = train_step(ids).float().cpu(); d2h += lv.numel() * lv.element_size()
        if cpu_debug:
            sync(); dt = time.perf_counter() - t0
        else:
            e1.record(); sync(); dt = e0.elapsed_time(e1) / 1e3          # CUDA events, as in the other arm
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": gbs * S * args.steps / float(tt), "unit": "tokens/s", "h2d_bytes_per_step": h2d // args.steps,
               "d2h_bytes_per_step": d2h // args.steps}
    if rank == 0:
        print(json.dumps({
            "impl": "reference", "metric": "Llama-2-7B training tokens/sec (whole job, device-timed, max over ranks)",
            "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "dtype": "bf16",
            "data": "synthetic", "vs_baseline": value / (6.90 * 8192),
            "config": {"model": "llama2-7b" if L == 32 else f"llama2-7b-{L}L(debug)", "global_batch": gbs, "micro_batch": mbs,
                       "seq_len": S, "parallelism": f"tp{tp}" + ("+sp" if sp else ""),
                       "stack": "unmodified reference (baseline/_ref) + xla_stubs: NCCL collectives, cuBLAS GEMM, SDPA attention",
                       "stub_device_rewrite": rewrite_scope,
                       "final_loss": float(loss)},
            "e2e": e2e, "gpu_launches": None}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0


def _rms_takes_sp(cls) -> bool:
    import inspect

    return "sequence_parallel_enabled" in inspect.signature(cls.__init__).parameters
