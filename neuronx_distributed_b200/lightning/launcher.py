"""Process launching (reference ``lightning/launcher.py:12-91``, which spawns the workers with ``xmp.spawn``).

Two situations:

* the job was started by ``torchrun`` (``RANK`` / ``WORLD_SIZE`` in the environment — the normal case: one process per GPU):
  the ranks already exist, ``launch`` runs the function in this process;
* a plain ``python script.py`` with ``devices = N > 1``: ``launch`` spawns ``N`` workers on this node (``torch.multiprocessing``
  spawn context, rendezvous on 127.0.0.1), each with its rank environment set the way ``torchrun`` would, joins them and returns
  the value worker 0 produced — the contract of the reference launcher.  Everything passed through must pickle and the script's
  entry point must sit under ``if __name__ == "__main__"``."""
from __future__ import annotations

import os
import socket
import traceback
from typing import Any, Callable, Optional


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, function: Callable, args, kwargs, queue) -> None:
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), LOCAL_WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    try:
        out = function(*args, **kwargs)
        if rank == 0:
            queue.put(("ok", out))
    except BaseException:  # noqa: BLE001  (reported to the parent, which re-raises)
        queue.put(("error", f"[rank {rank}]\n{traceback.format_exc()}"))
        raise


class _NeuronXLALauncher:
    is_interactive_compatible = False

    def __init__(self, strategy=None, num_processes: Optional[int] = None) -> None:
        self._strategy = strategy
        self.num_processes = num_processes

    def _world(self) -> int:
        if self.num_processes is not None:
            return int(self.num_processes)
        n = getattr(self._strategy, "num_processes", None)
        return int(n) if n else 1

    def launch(self, function: Callable, *args: Any, trainer=None, **kwargs: Any) -> Any:
        world = self._world()
        if "RANK" in os.environ or world <= 1:                # torchrun (or a single process): the rank is this process
            return function(*args, **kwargs)
        import torch.multiprocessing as mp

        ctx = mp.get_context("spawn")
        queue = ctx.SimpleQueue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, function, args, kwargs, queue)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        result, errors = None, []
        while not queue.empty():
            kind, payload = queue.get()
            if kind == "ok":
                result = payload
            else:
                errors.append(payload)
        if errors or any(p.exitcode != 0 for p in procs):
            raise RuntimeError("launched workers failed:\n" + "\n".join(errors or [f"exit codes {[p.exitcode for p in procs]}"]))
        return result
