"""``NxDModel`` — runtime holding the compiled bucket programs (reference ``trace/nxd_model/nxd_model.py:41-969``).

* a **shape-based router** selects the program whose example input shapes match (per key, first fit by total size);
* each :class:`BucketProgram` owns persistent input buffers and, on CUDA, a captured graph: ``forward`` copies the
  inputs into the static buffers, replays the graph and returns the static outputs;
* state (KV cache) lives in the wrapped module and is shared by all programs;
* ``save``/``load`` persist weights as per-rank safetensors."""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn


class BucketProgram:
    def __init__(self, key: str, module: nn.Module, fn: Callable, example: Tuple[torch.Tensor, ...],
                 use_cuda_graph: bool = True, warmup: int = 2):
        self.key, self.module, self.fn = key, module, fn
        self.shapes = tuple(tuple(t.shape) for t in example)
        self.dtypes = tuple(t.dtype for t in example)
        self.static_in = [t.clone() for t in example]
        self.graph = None
        self.static_out = None
        cuda = all(t.is_cuda for t in example) and torch.cuda.is_available()
        if use_cuda_graph and cuda:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s), torch.no_grad():
                for _ in range(warmup):
                    self.fn(self.module, *self.static_in)
            torch.cuda.current_stream().wait_stream(s)
            self.graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(self.graph):
                self.static_out = self.fn(self.module, *self.static_in)

    def matches(self, inputs: Sequence[torch.Tensor]) -> bool:
        return len(inputs) == len(self.shapes) and all(tuple(t.shape) == s for t, s in zip(inputs, self.shapes))

    def __call__(self, *inputs: torch.Tensor):
        if self.graph is None:
            with torch.no_grad():
                return self.fn(self.module, *inputs)
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out


class BaseNxDModel(nn.Module):
    pass


class NxDModel(BaseNxDModel):
    def __init__(self, world_size: int = 1, router: Any = None, start_rank: Optional[int] = None, local_ranks_size: Optional[int] = None):
        super().__init__()
        self.world_size, self.custom_router = world_size, router
        self.programs: Dict[str, List[BucketProgram]] = {}
        self.loaded_on_device = False

    def add_program(self, prog: BucketProgram) -> None:
        self.programs.setdefault(prog.key, []).append(prog)
        self.programs[prog.key].sort(key=lambda p: sum(int(torch.tensor(s).prod()) for s in p.shapes))

    def get_available_keys(self) -> List[str]:
        return list(self.programs)

    def router(self, inputs: Sequence[torch.Tensor], key: Optional[str] = None) -> BucketProgram:
        if self.custom_router is not None:
            r = self.custom_router(inputs)
            if isinstance(r, BucketProgram):
                return r
            key = r if isinstance(r, str) else key
        keys = [key] if key is not None else list(self.programs)
        for k in keys:
            for p in self.programs.get(k, []):
                if p.matches(inputs):
                    return p
        raise ValueError(f"no compiled bucket for input shapes {[tuple(t.shape) for t in inputs]} (keys {keys})")

    def forward(self, *inputs: torch.Tensor, model_name: Optional[str] = None, forward_mode: str = "default"):
        return self.router(inputs, model_name)(*inputs)

    # ---- weights -----------------------------------------------------------------------
    def set_weights(self, sharded_checkpoint: Sequence[Dict[str, torch.Tensor]]) -> None:
        from ..parallel_layers import parallel_state as ps

        r = ps.get_tensor_model_parallel_rank() if ps.model_parallel_is_initialized() else 0
        sd = sharded_checkpoint[r] if len(sharded_checkpoint) > r else sharded_checkpoint[0]
        seen = set()
        for progs in self.programs.values():
            for p in progs:
                if id(p.module) not in seen:
                    seen.add(id(p.module))
                    p.module.load_state_dict(sd, strict=False)

    def to_neuron(self) -> None:     # reference name; weights already live on the device
        self.loaded_on_device = True

    to_device = to_neuron

    def save(self, path: str, save_weights: bool = False) -> None:
        from ..parallel_layers import parallel_state as ps
        from ..utils.safetensors_utils import save_state_dict_safetensors

        os.makedirs(path, exist_ok=True)
        meta = {k: [{"shapes": p.shapes, "dtypes": [str(d) for d in p.dtypes]} for p in v] for k, v in self.programs.items()}
        torch.save(meta, os.path.join(path, "nxd_model_meta.pt"))
        if save_weights:
            r = ps.get_tensor_model_parallel_rank() if ps.model_parallel_is_initialized() else 0
            seen = {}
            for progs in self.programs.values():
                for p in progs:
                    seen[id(p.module)] = p.module
            for i, m in enumerate(seen.values()):
                save_state_dict_safetensors(m.state_dict(), os.path.join(path, f"weights_{i}_tp{r}.safetensors"))
