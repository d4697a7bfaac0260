"""BASELINE config 4 through ``bench.py --mode pp``: GPT-NeoX-20B pre-training step with tensor parallel x pipeline parallel
(NxDPPModel, 1F1B schedule, NCCL p2p between stages, fused TP kernels inside each stage), bf16, ZeRO-1 (fp32 master + fp32
gradient accumulation), synthetic tokens, random-init weights built directly on the device (meta-device init).

    torchrun --nproc-per-node 8 bench.py --mode pp --gpus 8            # TP=4 x PP=2 (the BASELINE shape)
    torchrun --nproc-per-node 4 bench.py --mode pp --gpus 4            # TP=2 x PP=2

value = tokens/s of the whole job (device-timed with CUDA events, max over ranks).  The pipeline bubble is measured, not
assumed: the step is timed with m and 2m micro-batches; t_mb = (T(2m) - T(m)) / m is the steady-state cost of one micro-batch
and bubble = 1 - m * t_mb / T(m), reported next to the 1F1B ideal (pp - 1) / (m + pp - 1).
Reference: examples/training/tp_dp_gpt_neox_hf_pretrain (model), pipeline/model.py:968-1009 (run_train), scheduler.py:157-253.
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(args) -> int:
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus and world % 2 == 0, "--mode pp needs an even number of GPUs (PP=2)"
    cpu_debug = not torch.cuda.is_available()            # plumbing check on a CPU box (gloo, tiny shapes); never a benchmark
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29536")
    if cpu_debug:
        os.environ["NXD_CPU_MODE"] = "1"
        dev = torch.device("cpu")
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        torch.cuda.set_device(local_rank)
        dev = torch.device("cuda", local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    class _Ev:                                            # CUDA event on the GPU, wall clock in the CPU plumbing check
        def __init__(self):
            self.e = torch.cuda.Event(enable_timing=True) if not cpu_debug else None
            self.t = 0.0

        def record(self):
            if self.e is not None:
                self.e.record()
            else:
                import time
                self.t = time.perf_counter()

        def elapsed_time(self, other):
            return self.e.elapsed_time(other.e) if self.e is not None else (other.t - self.t) * 1e3
    import neuronx_distributed_b200 as nxd
    from neuronx_distributed_b200 import ops
    from neuronx_distributed_b200.models.gpt_neox import GPTNeoXConfig, GPTNeoXForCausalLM, GPTNeoXLayer
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.utils.adamw_fp32_optim_params import AdamW_FP32OptimParams

    pp = 2
    tp = args.tp if args.tp > 0 else world // pp
    assert tp * pp == world, "data parallel is not part of this mode"
    S = 2048 if args.seq == 4096 else args.seq
    m = args.global_batch if args.global_batch != 4 else 16          # micro-batches of one sequence per step
    sp = tp > 1
    layers = 44 if args.layers == 32 else args.layers
    dims = dict(hidden_size=6144, num_attention_heads=64, intermediate_size=24576, vocab_size=50432)
    if cpu_debug:
        dims, layers, S, m = dict(hidden_size=64, num_attention_heads=4, intermediate_size=128, vocab_size=256), 4, 32, 4

    def init_fn(module, device):
        for name, p in module.named_parameters(recurse=False):
            with torch.no_grad():
                if p.dim() > 1:
                    p.normal_(0.0, 0.02)
                elif "bias" in name:
                    p.zero_()
                else:
                    p.fill_(1.0)
        for b in module.buffers(recurse=False):
            pass

    def build(num_mb):
        pcfg = {"transformer_layer_cls": GPTNeoXLayer, "num_microbatches": num_mb, "output_loss_value_spec": (True, False),
                "input_names": ["input_ids", "labels"], "broadcast_and_average_loss": True}
        cfg = nxd.neuronx_distributed_config(
            tensor_parallel_size=tp, pipeline_parallel_size=pp, sequence_parallel=sp, pipeline_config=pcfg,
            optimizer_config={"zero_one_enabled": True, "grad_clipping": True, "max_grad_norm": 1.0},
            mixed_precision_config={"use_master_weights": True, "use_fp32_grad_acc": True, "use_master_weights_in_ckpt": False},
            model_init_config={"meta_device_init": True, "param_init_fn": init_fn, "sequential_move_factor": 11})
        return cfg

    cfg = build(m)
    mcfg = GPTNeoXConfig(sequence_parallel_enabled=sp, dtype=torch.float32 if cpu_debug else torch.bfloat16, max_position_embeddings=S,
                         device=None, num_hidden_layers=layers, **dims)

    def model_fn():
        torch.manual_seed(1234)
        if not cpu_debug:
            torch.cuda.manual_seed(1234)
        return GPTNeoXForCausalLM(mcfg)

    model = nxd.initialize_parallel_model(cfg, model_fn)
    opt = nxd.initialize_parallel_optimizer(cfg, AdamW_FP32OptimParams, model.parameters(), lr=1e-4, betas=(0.9, 0.95), weight_decay=0.1)
    gen = torch.Generator().manual_seed(11)
    n_host = args.steps + args.warmup + 2
    host = [torch.randint(0, mcfg.vocab_size, (2 * m, S), generator=gen) for _ in range(n_host)]
    if not cpu_debug:
        host = [h.pin_memory() for h in host]
    devb = [h.to(dev) for h in host]
    pipe = model.original_module() if hasattr(model, "original_module") else model

    def set_microbatches(n):
        eng = getattr(model, "module", model)
        for obj in (eng, getattr(eng, "module", None)):
            if obj is not None and hasattr(obj, "num_microbatches"):
                obj.num_microbatches = n

    def step(ids, n_mb):
        opt.zero_grad()
        loss = model.run_train(input_ids=ids[:n_mb], labels=ids[:n_mb])
        opt.step()
        return loss

    def sync():
        dist.barrier()
        if not cpu_debug:
            torch.cuda.synchronize()

    def timed(n_mb, n_steps, offset):
        set_microbatches(n_mb)
        step(devb[0], n_mb); sync()                                # shape-metadata exchange / allocator warm-up for this count
        e0, e1 = _Ev(), _Ev()
        sync(); e0.record()
        for i in range(n_steps):
            loss = step(devb[offset + i], n_mb)
        e1.record(); sync()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t) / n_steps, loss

    for i in range(args.warmup):
        step(devb[i], m)
    sync()
    ops._ext.reset_launches()
    ms_m, loss = timed(m, args.steps, args.warmup)
    launches = ops._ext.launches()
    ms_2m, _ = timed(2 * m, max(2, args.steps // 2), args.warmup)
    t_mb = (ms_2m - ms_m) / m
    bubble = max(0.0, 1.0 - m * t_mb / ms_m)
    # end to end: tokens from pinned host memory every step, loss read back
    set_microbatches(m)
    e2e = None
    if not args.no_e2e:
        sync()
        a, b = _Ev(), _Ev()
        a.record(); h2d = d2h = 0
        for i in range(args.steps):
            ids = host[i % len(host)][:m].to(dev, non_blocking=True); h2d += m * S * 8
            lv = step(ids, m)
            lv = lv.float().cpu() if torch.is_tensor(lv) else torch.tensor(float(lv)); d2h += 4
        b.record(); sync()
        tt = torch.tensor([a.elapsed_time(b) / 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": m * S * args.steps / float(tt), "unit": "tokens/s", "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps}
    if rank == 0:
        value = m * S / (ms_m / 1e3)
        print(json.dumps({
            "metric": "GPT-NeoX-20B training tokens/sec, TP x PP=2 1F1B (whole job, device-timed, max over ranks)",
            "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_m,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "impl": "ours",
            "config": {"model": "gpt-neox-20b" if layers == 44 else f"gpt-neox-20b-{layers}L(debug)", "global_batch": m, "micro_batch": 1,
                       "seq_len": S, "parallelism": f"tp{tp}" + ("+sp" if sp else "") + f" x pp{pp} (1F1B)",
                       "optimizer": "AdamW fp32 master + fp32 grad-acc (ZeRO-1, dp=1)", "final_loss": float(loss),
                       "l2": "inputs(weights+activations)>>L2, no flush needed"},
            "pipeline": {"num_microbatches": m, "ms_per_step_2m": ms_2m, "ms_per_microbatch_steady": t_mb,
                         "bubble_fraction_measured": bubble, "bubble_fraction_1f1b_ideal": (pp - 1) / (m + pp - 1)},
            "e2e": e2e, "gpu_launches": launches}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0
