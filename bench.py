#!/usr/bin/env python
"""Headline benchmark: Llama-2-7B training tokens/s on N GPUs of one node with TP=N (BASELINE.json).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one optimizer step over `--global-batch` sequences of `--seq` tokens (micro-batch 1,
gradient accumulation), bf16 compute, fp32 master weights + fp32 gradient accumulation + AdamW through the
ZeRO-1 optimizer (DP=1), sequence parallel on for TP>1 — the reference's tp_zero1_llama2_7B recipe
(examples/training/llama/tp_zero1_llama_hf_pretrain/tp_zero1_llama2_7B_hf_pretrain.sh:21-41,155-172) with
synthetic tokens and random-init weights.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# published number closest to the metric: Llama-2-7B gate 6.90 seq/s at seq 8192 on 32 NeuronCores
# (test/integration/llama2_7B/test_long_seqlen.py:95-97) = 56,525 tokens/s
PUBLISHED_TOKENS_PER_S = 6.90 * 8192


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "decode", "pp"],
                    help="train (default, the headline: Llama-2-7B training tokens/s) | decode (BASELINE config 5: Llama-2-13B "
                         "inference latency, batch 1, seq 2048) | pp (BASELINE config 4: GPT-NeoX-20B TP x PP=2 1F1B training)")
    ap.add_argument("--seq", type=int, default=4096)
    ap.add_argument("--global-batch", type=int, default=4)
    ap.add_argument("--layers", type=int, default=32, help="32 = Llama-2-7B (anything else is a debug run)")
    ap.add_argument("--backend", default=os.environ.get("NXD_TP_BACKEND", "fused"), choices=["fused", "nccl"])
    ap.add_argument("--act-ckpt", default="none", choices=["none", "full"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-comm-report", action="store_true",
                    help="skip the extra pass after the timed regions (same step with the TP collectives replaced by local copies) "
                         "that reports the exposed TP-collective time per step")
    ap.add_argument("--tp", type=int, default=0, help="tensor-parallel degree (default: gpus / dp)")
    ap.add_argument("--dp", type=int, default=1, help="data-parallel degree: ZeRO-1 shards the optimizer over it and its gradient "
                                                      "reduce-scatter kernel runs in the step (BASELINE config 3: --gpus 8 --tp 4 --dp 2)")
    ap.add_argument("--micro-batch", type=int, default=0,
                    help="sequences per forward/backward (0 = auto: 2 up to 2 GPUs, 4 beyond — fewer fp32 wgrad read-modify-write "
                         "passes and TP collectives off their latency floor; 4 does not fit in 180 GB at TP=1); same value in both arms")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region (rank 0)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.proc, self.path, self.gpu = None, f"/tmp/nxd_clocks_{os.getpid()}.csv", gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def run_reference(args):
    """Reference arm: the unmodified reference from baseline/_ref through its own API (see baseline/README.md)."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import reference_arm  # noqa: F401

        return reference_arm.main(args)
    except Exception as e:  # noqa: BLE001
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"{type(e).__name__}: {str(e)[:160]}"}))
        return 0


def _tp_comm_info(tp: int, backend: str) -> dict:
    """What crosses NVLink on the TP hot path, stated in the JSON line (wire precision must be in `config`)."""
    if tp == 1:
        return {"tp_collectives": "none (tp=1)", "wire_dtype": "n/a"}
    if backend != "fused":
        return {"tp_collectives": "NCCL all-gather / reduce-scatter + own GEMM", "wire_dtype": "fp32 reduce-scatter (reduce_dtype), bf16 all-gather"}
    try:
        from neuronx_distributed_b200.ops import _fused_impl
        from neuronx_distributed_b200.parallel_layers import parallel_state as ps

        ws = _fused_impl.workspace(ps.get_tensor_model_parallel_group())
        if ws.nv is not None and ws.nvls_enabled():
            mode = "NVLS multimem.st all-gather / multimem.ld_reduce reduce-scatter fused in the GEMM kernels" if ws.nv.has_multicast \
                else "NVLS kernels on unicast peer pointers (no multicast mapping)"
            wd = _fused_impl.wire_dtype()
            return {"tp_collectives": mode,
                    "wire_dtype": "bf16 partials, fp32 accumulation in the switch (one rounding)" if wd == "bf16" else "fp32 partials (reference reduce_dtype)"}
        return {"tp_collectives": "cudaIpc peer pushes fused in the GEMM kernels", "wire_dtype": "bf16 partials, fp32 accumulation at the owner"}
    except Exception as e:  # noqa: BLE001
        return {"tp_collectives": f"unknown ({type(e).__name__})", "wire_dtype": "unknown"}


def _micro_batch(args) -> int:
    mbs = args.micro_batch if args.micro_batch > 0 else {1: 2, 2: 2, 4: 4, 8: 4}.get(args.gpus, 1)
    mbs = max(1, min(mbs, args.global_batch))
    while args.global_batch % mbs:
        mbs -= 1
    return mbs


def main():
    args = parse()
    if args.mode != "train":
        if args.impl == "reference":
            # the reference's inference path needs neuronx-cc (HLO → NEFF) and its PP engine needs XLA send/recv emulation over
            # 2-rank all-gathers of torch_xla tensors: neither exists on a CUDA box
            if int(os.environ.get("RANK", "0")) == 0:
                print(json.dumps({"impl": "reference", "mode": args.mode,
                                  "unavailable": "reference inference needs neuronx-cc/NEFF and its pipeline engine needs torch_xla lazy tensors; only the training arm runs on CUDA"}))
            return 0
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        if args.mode == "decode":
            import bench_decode

            return bench_decode.main(args)
        import bench_pp

        return bench_pp.main(args)
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    import neuronx_distributed_b200 as nxd
    from neuronx_distributed_b200 import ops
    from neuronx_distributed_b200.models.llama import LlamaForCausalLM, llama2_7b_config
    from neuronx_distributed_b200.utils.adamw_fp32_optim_params import AdamW_FP32OptimParams

    ops.tp_fused.set_backend(args.backend)
    dp = max(1, args.dp)
    tp = args.tp if args.tp > 0 else world // dp
    assert tp * dp == world, f"--tp {tp} x --dp {dp} != {world} GPUs"
    # dp > 1: ZeRO-1's reduce-scatter / all-gather kernels run in optimizer.step().  The bucketed variant that launches them
    # under the backward (NXD_ZERO1_OVERLAP=1) measured SLOWER on 2 GPUs (610 vs 393 ms/step: its 16 resident CTAs keep the
    # persistent 148-CTA GEMMs from being fully resident, so every GEMM runs a second partial wave) and stays opt-in.
    sp = tp > 1
    cfg = nxd.neuronx_distributed_config(
        tensor_parallel_size=tp, sequence_parallel=sp,
        optimizer_config={"zero_one_enabled": True, "grad_clipping": True, "max_grad_norm": 1.0},
        mixed_precision_config={"use_master_weights": True, "use_fp32_grad_acc": True, "use_master_weights_in_ckpt": False},
    )
    mcfg = llama2_7b_config(sequence_parallel_enabled=sp, dtype=torch.bfloat16, device=dev,
                            max_position_embeddings=args.seq, activation_checkpointing=args.act_ckpt)
    mcfg.num_hidden_layers = args.layers

    def model_fn():
        torch.manual_seed(1234)
        torch.cuda.manual_seed(1234)
        return LlamaForCausalLM(mcfg)

    model = nxd.initialize_parallel_model(cfg, model_fn)
    opt = nxd.initialize_parallel_optimizer(cfg, AdamW_FP32OptimParams, model.parameters(), lr=3e-4,
                                            betas=(0.9, 0.95), weight_decay=0.1)
    gbs, S = args.global_batch, args.seq
    mbs = _micro_batch(args)
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    dp_rank = ps.get_data_parallel_rank()
    assert gbs % dp == 0, "global batch must divide by dp"
    per_dp = gbs // dp
    mbs = max(1, min(mbs, per_dp))
    while per_dp % mbs:
        mbs -= 1
    V = mcfg.vocab_size
    gen = torch.Generator().manual_seed(7)
    n_host = args.steps + args.warmup + 2
    host_ids = [torch.randint(0, V, (gbs, S), generator=gen).pin_memory() for _ in range(n_host)]
    dev_ids = [h.to(dev) for h in host_ids]      # every step sees a fresh synthetic batch

    phases = [] if os.environ.get("NXD_BENCH_PHASES", "0") == "1" else None     # debug: device time of optimizer.step()

    def train_step(ids_dev):
        opt.zero_grad()
        total = None
        n_mb = per_dp // mbs
        for i, mb in enumerate(range(dp_rank * per_dp, (dp_rank + 1) * per_dp, mbs)):
            ids = ids_dev[mb:mb + mbs]
            # gradient accumulation: only the last micro-batch may release ZeRO-1 buckets to the overlapped reduce-scatter
            ctx = opt.no_sync() if (dp > 1 and i < n_mb - 1 and hasattr(opt, "no_sync")) else contextlib.nullcontext()
            with ctx:
                loss = model.run_train(input_ids=ids, labels=ids)
            total = loss if total is None else total + loss
        if phases is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); opt.step(); e1.record()
            phases.append((e0, e1))
        else:
            opt.step()
        return total / n_mb

    def sync():
        dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ---------------------------------------------------------------------
    trace = []
    for i in range(args.warmup):
        trace.append(train_step(dev_ids[i]))
    sync()
    # ---- device-timed region: inputs resident on device, CUDA events, max over ranks ---
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ops._ext.reset_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    ev0.record()
    for i in range(args.steps):
        loss = train_step(dev_ids[args.warmup + i])
        trace.append(loss)
    ev1.record()
    sync()
    launches = ops._ext.launches()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    clocks = sampler.stop() if sampler else None
    final_loss = float(loss.item())
    tokens_per_step = gbs * S
    value = tokens_per_step * args.steps / (ms_total / 1e3)

    # ---- end-to-end region through the public API: pinned host → device every step, loss → host ----
    e2e = None
    if not args.no_e2e:
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h2d = d2h = 0
        for i in range(args.steps):
            h = host_ids[i % len(host_ids)]
            ids = h.to(dev, non_blocking=True)
            h2d += h.numel() * h.element_size()
            l = train_step(ids)
            lv = l.float().cpu()       # device→host read of the step's loss
            d2h += lv.numel() * lv.element_size()
        e1.record()
        sync()
        dt = e0.elapsed_time(e1) / 1e3          # device-timed (CUDA events on the launching stream), like the headline value
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": tokens_per_step * args.steps / float(tt.item()), "unit": "tokens/s",
               "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps}

    # ---- exposed tensor-parallel communication (BASELINE metric): same step with the TP collectives replaced by local copies of
    #      the same shape ("nocomm" backend, numerically meaningless) → exposed = t(step) − t(nocomm step).  Runs last because it
    #      trashes the weights; any failure only drops this key.
    comm_report = None
    if tp > 1 and args.backend == "fused" and not args.no_comm_report:
        try:
            ops.tp_fused.set_backend("nocomm")
            train_step(dev_ids[0])
            sync()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for i in range(args.steps):
                train_step(dev_ids[args.warmup + i])
            c1.record()
            sync()
            tc = torch.tensor([c0.elapsed_time(c1)], device=dev, dtype=torch.float64)
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            nocomm_ms = float(tc.item()) / args.steps
            comm_report = {"compute_only_ms_per_step": nocomm_ms,
                           "exposed_tp_collective_ms_per_step": ms_total / args.steps - nocomm_ms,
                           "method": "same step with Column/Row TP collectives replaced by local copies (tp backend 'nocomm')"}
        except Exception as e:  # noqa: BLE001
            comm_report = {"error": f"{type(e).__name__}: {str(e)[:120]}"}
        finally:
            ops.tp_fused.set_backend(args.backend)

    if rank == 0:
        out = {
            "metric": "Llama-2-7B training tokens/sec (whole job, device-timed, max over ranks)",
            "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": value / PUBLISHED_TOKENS_PER_S, "dtype": "bf16", "data": "synthetic",
            "impl": "ours",
            "config": {"model": "llama2-7b" if args.layers == 32 else f"llama2-7b-{args.layers}L(debug)",
                       "global_batch": gbs, "micro_batch": mbs, "seq_len": S,
                       "parallelism": f"tp{tp}" + ("+sp" if sp else "") + (f" x dp{dp}" if dp > 1 else ""),
                       "optimizer": f"AdamW fp32 master + fp32 grad-acc (ZeRO-1, dp={dp}"
                                    + (", peer-memory reduce-scatter / all-gather kernels" if dp > 1 else "")
                                    + (", bucketed under backward" if dp > 1 and os.environ.get("NXD_ZERO1_OVERLAP", "0") == "1" else "") + ")",
                       "tp_backend": args.backend, **_tp_comm_info(tp, args.backend),
                       "act_ckpt": args.act_ckpt, "l2": "inputs(weights+activations)>>L2, no flush needed",
                       "baseline_note": "vs_baseline divides by the only published number: Trn1 32-core gate 6.90 seq/s @ seq 8192",
                       "final_loss": final_loss, "loss_trace": [round(float(x), 4) for x in trace],
                       "grad_norm": float(opt.grad_norm) if opt.grad_norm is not None else None},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "comm": comm_report,
        }
        if phases:
            torch.cuda.synchronize()
            ts = [a.elapsed_time(b) for a, b in phases[args.warmup:args.warmup + args.steps]]
            out["optimizer_step_ms"] = sum(ts) / max(1, len(ts))
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
