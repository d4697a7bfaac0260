"""Shared helpers for the training examples (role of reference ``examples/training/llama/training_utils.py:268-376``):
moving-average throughput, a JSON metrics file, synthetic / memory-mapped token datasets, distributed bring-up."""
from __future__ import annotations

import json
import os
import time
from typing import Dict, Iterator

import numpy as np
import torch
import torch.distributed as dist


def init_distributed():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if torch.cuda.is_available() and os.environ.get("NXD_CPU_MODE", "0") != "1":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        return torch.device("cuda", local)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    return torch.device("cpu")


class Throughput:
    """seq/s = window·(batch·dp·grad_accum·log_interval)/window_time, moving average over ``moving_avg_window`` steps."""

    def __init__(self, batch_size: int, world_size_dp: int, grad_accum_usteps: int, moving_avg_window_size: int = 10,
                 logging_interval: int = 1):
        self.seqs_per_iteration = batch_size * world_size_dp * grad_accum_usteps * logging_interval
        self.moving_avg_window_size = moving_avg_window_size
        self.window = []
        self.window_time = 0.0
        self.start = time.time()
        self.peak = 0.0

    def get_throughput(self) -> float:
        now = time.time()
        dt, self.start = now - self.start, now
        self.window.append(dt)
        self.window_time += dt
        if len(self.window) > self.moving_avg_window_size:
            self.window_time -= self.window.pop(0)
        tp = len(self.window) * self.seqs_per_iteration / max(self.window_time, 1e-9)
        self.peak = max(self.peak, tp)
        return tp


class TrainingMetrics:
    def __init__(self, path: str):
        self.path = path
        self.data: Dict[str, Dict] = {}

    def store_parameters(self, params: Dict) -> None:
        self.data.setdefault("parameters", {}).update(params)
        self._flush()

    def store_metrics(self, metrics: Dict) -> None:
        self.data.setdefault("metrics", {}).update(metrics)
        self._flush()

    def _flush(self) -> None:
        if (dist.get_rank() if dist.is_initialized() else 0) == 0:
            with open(self.path, "w") as f:
                json.dump(self.data, f, indent=1, default=str)


def synthetic_batches(vocab: int, batch: int, seq: int, seed: int, device) -> Iterator[Dict[str, torch.Tensor]]:
    g = torch.Generator().manual_seed(seed)
    pin = torch.cuda.is_available()
    while True:
        ids = torch.randint(0, vocab, (batch, seq), generator=g)
        if pin:
            ids = ids.pin_memory()
        ids = ids.to(device, non_blocking=True)
        yield {"input_ids": ids, "labels": ids}


def memmap_batches(path: str, batch: int, seq: int, dp_rank: int, dp_size: int, device, dtype=None) -> Iterator[Dict[str, torch.Tensor]]:
    """Flat token file → [batch, seq] windows, strided over data-parallel ranks.  The element type comes from the ``<path>.json``
    side file written by ``llama/get_dataset.py`` (``uint16`` without one)."""
    if dtype is None:
        dtype = np.uint16
        if os.path.exists(path + ".json"):
            with open(path + ".json") as f:
                dtype = np.dtype(json.load(f).get("dtype", "uint16"))
    data = np.memmap(path, dtype=dtype, mode="r")
    assert len(data) > seq, f"{path} holds {len(data)} tokens, fewer than one sequence of {seq}"
    n = (len(data) - 1) // seq
    i = dp_rank
    while True:
        rows = []
        for _ in range(batch):
            rows.append(torch.from_numpy(data[(i % n) * seq:(i % n) * seq + seq].astype(np.int64)))
            i += dp_size
        ids = torch.stack(rows).to(device, non_blocking=True)
        yield {"input_ids": ids, "labels": ids}


def linear_warmup_cosine(optimizer, warmup: int, total: int, min_ratio: float = 0.1, constant: int = 0):
    """Linear warm-up, `constant` steps at the peak, cosine decay to `min_ratio` x peak at `total`."""
    import math

    def f(step):
        if step < warmup:
            return (step + 1) / max(1, warmup)
        if step < warmup + constant:
            return 1.0
        p = min(1.0, (step - warmup - constant) / max(1, total - warmup - constant))
        return min_ratio + (1 - min_ratio) * 0.5 * (1 + math.cos(math.pi * p))

    return torch.optim.lr_scheduler.LambdaLR(optimizer, f)


# ---- checkpoint / resume flags shared by the smaller example scripts (BERT, GPT-NeoX, MoE) ---------------------------------
def add_checkpoint_args(p) -> None:
    """Flags of the reference's BERT / GPT-NeoX scripts (``--resume_ckpt``, ``--resume_step``, ``--resume_ckpt_path``,
    ``--minimal_ckpt``, ``--shards_per_ckpt``, ``--test_checkpointing``) over ``nxd.save_checkpoint`` tags ``step_<n>``."""
    p.add_argument("--output_dir", default="./output")
    p.add_argument("--checkpoint_freq", "--shards_per_ckpt", dest="checkpoint_freq", type=int, default=0,
                   help="save every N steps (the reference counts data shards; a step is the unit here); 0 = never")
    p.add_argument("--resume_ckpt", action="store_true", help="resume from --resume_step, or from the newest complete checkpoint")
    p.add_argument("--resume_step", type=int, default=-1)
    p.add_argument("--resume_ckpt_path", default=None, help="checkpoint directory to resume from instead of <output_dir>/checkpoints")
    p.add_argument("--minimal_ckpt", action="store_true", help="model weights only (no optimizer / scheduler state)")
    p.add_argument("--test_checkpointing", action="store_true",
                   help="after the first save, load it back into the live model and check that nothing changed")
    p.add_argument("--steps_this_run", type=int, default=-1)
    p.add_argument("--debug", action="store_true", help="print the batch shapes and the checkpoint actions")
    p.add_argument("--enable_pt_autocast", action="store_true", help="accepted: bf16 parameters, fp32 statistics is the only mode")


def maybe_resume(a, nxd, model, opt, sched=None) -> int:
    """Returns the step to continue from (0 without ``--resume_ckpt`` or without a checkpoint)."""
    d = a.resume_ckpt_path or os.path.join(a.output_dir, "checkpoints")
    if not a.resume_ckpt or not nxd.has_checkpoint(d):
        return 0
    tag = f"step_{a.resume_step}" if a.resume_step >= 0 else None
    uc = nxd.load_checkpoint(d, tag=tag, model=model, optimizer=None if a.minimal_ckpt else opt,
                             scheduler=None if a.minimal_ckpt else sched)
    step = int((uc or {}).get("total_steps", max(a.resume_step, 0)))
    if a.debug and dist.get_rank() == 0:
        print(f"resumed from {d} ({tag or 'newest'}) at step {step}", flush=True)
    return step


def maybe_save(a, nxd, model, opt, step: int, sched=None) -> None:
    if a.checkpoint_freq <= 0 or step % a.checkpoint_freq:
        return
    d = os.path.join(a.output_dir, "checkpoints")
    nxd.save_checkpoint(d, f"step_{step}", model=model, optimizer=None if a.minimal_ckpt else opt,
                        scheduler=None if a.minimal_ckpt else sched, user_content={"total_steps": step}, use_xser=False,
                        num_kept_ckpts=2)
    if a.debug and dist.get_rank() == 0:
        print(f"saved {d}/step_{step}", flush=True)
    if a.test_checkpointing and not getattr(a, "_ckpt_tested", False):
        a._ckpt_tested = True
        before = {k: v.detach().clone() for k, v in model.state_dict().items()}
        nxd.load_checkpoint(d, tag=f"step_{step}", model=model, optimizer=None if a.minimal_ckpt else opt)
        for k, v in model.state_dict().items():
            assert torch.equal(v, before[k]), f"checkpoint round trip changed {k}"
        if dist.get_rank() == 0:
            print("checkpoint round trip ok", flush=True)

