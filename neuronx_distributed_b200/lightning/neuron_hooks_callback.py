"""Per-layer activation / gradient dumps to disk for numerical bisecting (reference ``lightning/neuron_hooks_callback.py:9-218``)."""
from __future__ import annotations

import os
from typing import Any, List, Optional

import torch

from ._compat import Callback


class NeuronHooksCallback(Callback):
    """Two ways to configure: keyword arguments (dump the outputs / output-gradients of leaf modules matching
    ``module_filter`` at the listed ``steps``), or the reference's config object (``cfg.hooks``, ``cfg.target_layers`` =
    comma-separated module names, ``cfg.hooks_interval``, ``cfg.enable_activation_dumps`` / ``enable_grad_dumps``,
    ``cfg.dump_only_norms``, ``cfg.dump_only_master_rank``, ``cfg.master_print_model_layers``), which records
    (input, output) pairs per target layer every ``hooks_interval`` global steps and writes them on batch end."""

    def __init__(self, cfg: Any = "hooks_dump", steps: Optional[List[int]] = None, module_filter: Optional[str] = None,
                 dump_grads: bool = True, dump_dir: Optional[str] = None, cfg_or_dump_dir: Any = None):
        """``cfg``: the reference's config object (attributes ``hooks``, ``hooks_dump_base_directory``, ``target_layers``, …;
        neuron_hooks_callback.py:9-40) or simply a dump directory."""
        cfg_or_dump_dir = cfg if cfg_or_dump_dir is None else cfg_or_dump_dir
        self._handles: List[Any] = []
        self._step = 0
        cfg = None if isinstance(cfg_or_dump_dir, str) else cfg_or_dump_dir
        self.dump_dir = dump_dir or (cfg_or_dump_dir if cfg is None else "hooks_dump")
        self.steps, self.module_filter, self.dump_grads = set(steps or [0]), module_filter, dump_grads
        self.hooks = bool(getattr(cfg, "hooks", False)) if cfg is not None else False
        self.activations_map: Optional[dict] = {} if self.hooks else None
        self.gradients_map: Optional[dict] = {} if self.hooks else None
        g = (lambda k, d=None: getattr(cfg, k, d)) if cfg is not None else (lambda k, d=None: d)
        self.hooks_dump_base_directory = g("hooks_dump_base_directory",
                                           f"./hooks_outputs/{os.environ.get('SLURM_JOB_ID', 'local')}/hooks_dumps") if self.hooks else None
        self.master_print_model_layers = bool(g("master_print_model_layers", False))
        tl = g("target_layers", "") or ""
        self.target_layers = [t.strip() for t in (tl.split(",") if isinstance(tl, str) else tl) if str(t).strip()]
        self.dump_only_master_rank, self.dump_only_norms = g("dump_only_master_rank"), g("dump_only_norms")
        self.hooks_interval = g("hooks_interval", 1) or 1
        self.enable_activation_dumps, self.enable_grad_dumps = bool(g("enable_activation_dumps", False)), bool(g("enable_grad_dumps", False))

    # ---- reference-style recording (neuron_hooks_callback.py:42-198) -----------------------------------------------------
    @staticmethod
    def process_input_output(input, output):  # noqa: A002
        """First tensor of each side, detached and copied to host (so dumping never extends device tensor lifetimes)."""
        pick = lambda v: v if isinstance(v, torch.Tensor) else next((t for t in v if isinstance(t, torch.Tensor)), None)  # noqa: E731
        i, o = pick(input), pick(output)
        return (None if i is None else i.detach().float().cpu()), (None if o is None else o.detach().float().cpu())

    def _record(self, store: dict, layer_name: str, pl_module, a, b) -> None:
        if getattr(pl_module, "global_step", self._step) % self.hooks_interval == 0:
            i, o = self.process_input_output(a, b)
            store.setdefault(layer_name, []).append((getattr(pl_module, "global_step", self._step), i, o))

    def create_forward_hook(self, layer_name: str, pl_module):
        return lambda module, inp, out: self._record(self.activations_map, layer_name, pl_module, inp, out)

    def create_backward_hook(self, layer_name: str, pl_module):
        return lambda module, grad_input, grad_output: self._record(self.gradients_map, layer_name, pl_module, grad_input, grad_output)

    def register_forward_hook_wrapper(self, layer_name: str, layer, pl_module) -> None:
        self._handles.append(layer.register_forward_hook(self.create_forward_hook(layer_name, pl_module)))

    def register_backward_hook_wrapper(self, layer_name: str, layer, pl_module) -> None:
        self._handles.append(layer.register_full_backward_hook(self.create_backward_hook(layer_name, pl_module)))

    def _save_map(self, store: Optional[dict], names) -> None:
        from ..parallel_layers import parallel_state as ps

        if not store:
            return
        if self.dump_only_master_rank and torch.distributed.is_initialized() and torch.distributed.get_rank() != 0:
            store.clear()
            return
        tag = ps.get_rank_info_str() if ps.model_parallel_is_initialized() else "rank0"
        for layer, records in store.items():
            for step, a, b in records:
                d = os.path.join(self.hooks_dump_base_directory, layer, f"global_step_{step}")
                os.makedirs(d, exist_ok=True)
                for nm, t in zip(names, (a, b)):
                    if t is not None:
                        torch.save(t.norm() if self.dump_only_norms else t, os.path.join(d, f"{nm}_{tag}.pt"))
        store.clear()

    def save_activations_map(self, pl_module=None) -> None:
        self._save_map(self.activations_map, ("input", "output"))

    def save_gradients_map(self, pl_module=None) -> None:
        self._save_map(self.gradients_map, ("grad_input", "grad_output"))

    def _want(self, name: str) -> bool:
        return self.module_filter is None or self.module_filter in name

    def attach(self, model: torch.nn.Module) -> None:
        os.makedirs(self.dump_dir, exist_ok=True)
        for name, mod in model.named_modules():
            if not name or not self._want(name) or list(mod.children()):
                continue

            def fwd(m, inp, out, _n=name):
                if self._step in self.steps:
                    t = out[0] if isinstance(out, (tuple, list)) else out
                    if isinstance(t, torch.Tensor):
                        torch.save(t.detach().float().cpu(), os.path.join(self.dump_dir, f"step{self._step}_{_n}_fwd.pt"))

            self._handles.append(mod.register_forward_hook(fwd))
            if self.dump_grads:
                def bwd(m, gin, gout, _n=name):
                    if self._step in self.steps and gout and isinstance(gout[0], torch.Tensor):
                        torch.save(gout[0].detach().float().cpu(), os.path.join(self.dump_dir, f"step{self._step}_{_n}_bwd.pt"))

                self._handles.append(mod.register_full_backward_hook(bwd))

    def on_train_start(self, trainer=None, pl_module=None) -> None:
        model = getattr(pl_module, "model", None)
        if model is None:
            return
        if not self.hooks:
            return self.attach(model)
        if self.master_print_model_layers and (not torch.distributed.is_initialized() or torch.distributed.get_rank() == 0):
            print("Printing Model Layers:\n" + "\n".join(n for n, _ in model.named_modules()))
        for name, layer in model.named_modules():
            # a target names a module of the USER's model; the trainer's wrappers (``module.`` / pipeline stage prefixes) may sit
            # in front of it, so a dotted-suffix match counts and the dump directory carries the name the user gave
            target = next((t for t in self.target_layers if name.strip() == t or name.endswith("." + t)), None)
            if target is not None:
                if self.enable_activation_dumps:
                    self.register_forward_hook_wrapper(target, layer, pl_module)
                if self.enable_grad_dumps:
                    self.register_backward_hook_wrapper(target, layer, pl_module)

    def on_train_batch_end(self, trainer=None, pl_module=None, *a, **k) -> None:
        self._step += 1
        if self.hooks:
            self.save_activations_map(pl_module)
            self.save_gradients_map(pl_module)

    def detach(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles.clear()
