"""Run a function on N local processes (gloo on CPU / nccl on GPUs) — the analogue of the
reference's ``NXD_CPU_MODE=1 torchrun --nproc-per-node=N`` integration tier."""
from __future__ import annotations

import os
import socket
import sys
import traceback

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, args, use_cuda, errq):
    try:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank), LOCAL_WORLD_SIZE=str(world))
        import torch.distributed as dist

        if use_cuda == "loopback":
            # every rank on cuda:0 — NCCL refuses two ranks on one device, so the control plane is gloo; the peer-memory
            # kernels (VMM fd exchange / cudaIpc both work intra-device) run their real flag / epoch protocols between the
            # two processes' contexts, which the GPU time-slices
            torch.cuda.set_device(0)
            os.environ["NXD_LOOPBACK"] = "1"
            dist.init_process_group("gloo", rank=rank, world_size=world)
        elif use_cuda:
            torch.cuda.set_device(rank)
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        else:
            os.environ["NXD_CPU_MODE"] = "1"
            torch.set_num_threads(1)
            dist.init_process_group("gloo", rank=rank, world_size=world)
        try:
            fn(rank, world, *args)
        finally:
            from neuronx_distributed_b200.parallel_layers import parallel_state as ps

            if ps.model_parallel_is_initialized():
                ps.destroy_model_parallel()
            dist.destroy_process_group()
    except Exception:
        errq.put((rank, traceback.format_exc()))
        raise


def run_distributed(fn, world: int, *args, use_cuda: bool = False, timeout: float = 300.0):
    if not use_cuda:
        # CPU workers re-import torch in every spawned process (slow on a cold or busy box): callers' tighter limits are floors
        timeout = max(float(timeout), 300.0)
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, args, use_cuda, errq)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout)
    alive = [p for p in procs if p.is_alive()]
    for p in alive:
        p.terminate()
    errs = []
    while not errq.empty():
        errs.append(errq.get())
    if errs:
        raise AssertionError("worker failure:\n" + "\n".join(f"[rank {r}]\n{tb}" for r, tb in errs))
    assert not alive, "distributed test timed out"
    for p in procs:
        assert p.exitcode == 0, f"worker exit code {p.exitcode}"
