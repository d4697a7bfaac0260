"""Import-time stand-ins for the Neuron-only dependencies of the UNMODIFIED reference (``baseline/_ref``).

The reference imports ``torch_xla``, ``torch_neuronx``, ``neuronxcc``, ``libneuronxla`` (and ``tenacity``)
unconditionally (reference parallel_state.py:7-20), none of which exist on a CUDA box.  This module installs a
meta-path finder that fabricates those packages:

* everything the *training path* really calls is implemented on top of ``torch.distributed``/NCCL and plain CUDA —
  ``xm.all_reduce / all_gather / reduce_scatter / all_to_all / mark_step / rendezvous / xla_device``, ``xr.*`` rank
  queries, ``get_platform_target``, and ``torch_xla.distributed.zero_redundancy_optimizer.ZeroRedundancyOptimizer``
  (dim-0 sharded optimizer state + fp32 master weights + reduce-scatter / all-gather);
* everything else resolves to inert placeholders so that ``import neuronx_distributed`` succeeds.

This is the "reference's own NCCL(+cuBLAS) build" of BASELINE.md: the reference's code runs unchanged, its XLA
collectives land on NCCL and its matmuls on cuBLAS.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import sys
import types
from typing import Optional, Sequence

import torch
import torch.distributed as dist

_PREFIXES = ("torch_xla", "torch_neuronx", "neuronxcc", "libneuronxla", "tenacity", "nkilib", "torchdistx", "nki", "neuronx_cc", "torch_neuron")
# modules of installed packages that newer releases removed (transformers 5 dropped the fx tracer the reference imports)
_EXACT = ("transformers.utils.fx", "boto3", "botocore", "awscrt", "s3transfer")


class _DummyMeta(type):
    """Class-level attribute access also yields placeholders (``structure.Packer``, ``hlo_pb2.HloModuleProto`` …)."""

    def __getattr__(cls, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _DummyMeta(name, (_Dummy,), {})

    def __getitem__(cls, item):
        return cls

    def __or__(cls, other):
        return cls

    __ror__ = __or__


class _Dummy(metaclass=_DummyMeta):
    """Callable / subscriptable / inheritable placeholder."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]          # used as a decorator
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __getitem__(self, item):
        return _Dummy()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)

    def __or__(self, other):
        return self

    __ror__ = __or__


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        val = _DummyMeta(name, (_Dummy,), {})
        setattr(self, name, val)
        return val


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _PREFIXES or fullname in _EXACT or fullname.split(".")[0] in _EXACT:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        _populate(module)


# ------------------------------------------------------------------------------------------------
# group resolution: the reference passes ``groups=[[ranks], …]`` (replica groups) to xm.* collectives
# ------------------------------------------------------------------------------------------------
_PG_CACHE = {}
_PG_BY_ID = {}          # id(groups) → (groups, pg): the reference passes the same replica-group lists on every call


def _pg_for(groups: Optional[Sequence[Sequence[int]]]):
    if groups is None:
        return dist.group.WORLD
    hit = _PG_BY_ID.get(id(groups))
    if hit is not None and hit[0] is groups:
        return hit[1]
    pg = _pg_for_slow(groups)
    _PG_BY_ID[id(groups)] = (groups, pg)      # keeps `groups` alive, so the id cannot be recycled
    return pg


def _pg_for_slow(groups):
    me = dist.get_rank()
    mine = None
    for g in groups:
        if me in g:
            mine = tuple(g)
            break
    assert mine is not None, f"rank {me} not in {groups}"
    if len(mine) == dist.get_world_size():
        return dist.group.WORLD
    key = tuple(tuple(g) for g in groups)
    if key not in _PG_CACHE:
        made = None
        for g in groups:                       # all ranks create all groups in the same order
            pg = dist.new_group(list(g))
            if me in g:
                made = pg
        _PG_CACHE[key] = made
    return _PG_CACHE[key]


_OPS = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}


def _xm_all_reduce(reduce_type, inputs, scale=1.0, groups=None, pin_layout=True):
    pg = _pg_for(groups)
    op = _OPS.get(str(reduce_type).lower().replace("reduce_", ""), dist.ReduceOp.SUM)
    single = isinstance(inputs, torch.Tensor)
    ts = [inputs] if single else list(inputs)
    for t in ts:
        if dist.get_world_size(pg) > 1:
            dist.all_reduce(t, op=op, group=pg)
        if scale != 1.0:
            t.mul_(scale)
    return inputs if single else ts


def _xm_all_gather(value, dim=0, groups=None, output=None, pin_layout=True):
    pg = _pg_for(groups)
    n = dist.get_world_size(pg)
    if n == 1:
        return value
    v = value.contiguous()
    if dist.get_backend(pg) == "gloo":             # CPU plumbing check only
        parts = [torch.empty_like(v) for _ in range(n)]
        dist.all_gather(parts, v, group=pg)
        res = torch.cat(parts, dim=dim)
    else:
        dim = dim % v.dim()
        try:
            if dim == 0:
                res = torch.empty((n * v.shape[0],) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
                dist.all_gather_into_tensor(res, v, group=pg)      # one NCCL call straight into the result, no cat copy
            else:
                stacked = torch.empty((n,) + tuple(v.shape), dtype=v.dtype, device=v.device)
                dist.all_gather_into_tensor(stacked, v, group=pg)
                res = torch.cat(list(stacked.unbind(0)), dim=dim)
        except (RuntimeError, ValueError, TypeError):              # e.g. a 0-d tensor: list form handles everything
            parts = [torch.empty_like(v) for _ in range(n)]
            dist.all_gather(parts, v, group=pg)
            res = torch.cat(parts, dim=dim) if v.dim() else torch.stack(parts)
    if output is not None:                      # torch_xla writes into `output` when given
        output.copy_(res)
        return output
    return res


def _xm_reduce_scatter(reduce_type, input, scale, scatter_dim, shard_count, groups=None, output=None, pin_layout=True):
    pg = _pg_for(groups)
    n = dist.get_world_size(pg)
    single = isinstance(input, torch.Tensor)
    ins = [input] if single else list(input)
    outs = []
    for t in ins:
        if n == 1:
            o = t
        else:
            xin = t.movedim(scatter_dim, 0).contiguous() if scatter_dim != 0 else t.contiguous()
            if dist.get_backend(pg) == "gloo":          # CPU plumbing check only
                full = xin.clone()
                dist.all_reduce(full, group=pg)
                o = full.chunk(n, 0)[dist.get_rank(pg)].contiguous()
            else:
                o = torch.empty((xin.shape[0] // n,) + tuple(xin.shape[1:]), dtype=t.dtype, device=t.device)
                dist.reduce_scatter_tensor(o, xin, group=pg)
            if scatter_dim != 0:
                o = o.movedim(0, scatter_dim).contiguous()
        if scale != 1.0:
            o = o * scale
        outs.append(o)
    if output is not None:                      # in-place form used by the reference's mappings
        if single:
            output.copy_(outs[0])
            return output
        for dst, src in zip(output, outs):
            dst.copy_(src)
        return output
    return outs[0] if single else outs


def _xm_all_to_all(value, split_dimension, concat_dimension, split_count, groups=None, pin_layout=True):
    pg = _pg_for(groups)
    n = dist.get_world_size(pg)
    if n == 1:
        return value
    pieces = [p.contiguous() for p in value.chunk(n, dim=split_dimension)]
    outs = [torch.empty_like(pieces[0]) for _ in range(n)]
    dist.all_to_all(outs, pieces, group=pg)
    return torch.cat(outs, dim=concat_dimension)


def _device(*a, **k):
    if not torch.cuda.is_available():       # CPU plumbing check only
        return torch.device("cpu")
    return torch.device("cuda", torch.cuda.current_device())


class ZeroRedundancyOptimizer(torch.optim.Optimizer):
    """NCCL stand-in for ``torch_xla.distributed.zero_redundancy_optimizer.ZeroRedundancyOptimizer`` with the keyword
    surface the reference uses (trainer.py:246-301): every parameter is padded and sharded along dim 0 over the
    sharding group; gradients are reduce-scattered (fp32 when ``higher_cc_precision``), the base optimizer steps fp32
    master shards, updated parameters are all-gathered back."""

    def __init__(self, params, optimizer_class, optimizer_dtype=None, grad_clipping=True, max_norm=None, pin_layout=True,
                 sharding_groups=None, grad_norm_groups=None, lazy_init=False, coalesce_cc=False, use_grad_acc_hook=False,
                 higher_cc_precision=False, save_master_weights=False, bucket_cap_mb_all_gather=0,
                 bucket_cap_mb_reduce_scatter=0, **defaults):
        super().__init__(params, defaults)
        self.optimizer_dtype = optimizer_dtype if optimizer_dtype not in (None, torch.double) else torch.float32
        self.grad_clipping, self.max_norm = grad_clipping, (max_norm if max_norm is not None else 1.0)
        self.sharding_groups, self.grad_norm_groups = sharding_groups, grad_norm_groups
        self._sharding_groups, self._grad_norm_groups = sharding_groups, grad_norm_groups
        self.use_grad_acc_hook, self.higher_cc_precision = use_grad_acc_hook, higher_cc_precision
        self.pg = _pg_for(sharding_groups)
        self.local_world_size = dist.get_world_size(self.pg)
        self.local_rank = dist.get_rank(self.pg)
        self.inited = False
        self.optimizer_class, self.opt_defaults = optimizer_class, defaults
        self._grad_norm = None
        self.init_zero()

    # ---- sharding helpers ---------------------------------------------------------------
    def _pad(self, t):
        n = self.local_world_size
        rows = t.shape[0] if t.dim() > 0 else 1
        pad = (-rows) % n
        t2 = t.reshape(rows, -1) if t.dim() > 0 else t.reshape(1, 1)
        if pad:
            t2 = torch.cat([t2, t2.new_zeros(pad, t2.shape[1])], 0)
        return t2

    def _shard(self, t):
        return self._pad(t).chunk(self.local_world_size, 0)[self.local_rank].clone()

    def _shard_parameters(self):
        return None

    def init_zero(self):
        base_groups = []
        self._pairs = []
        for g in self.param_groups:
            shards = []
            for p in g["params"]:
                if not p.requires_grad:
                    continue
                s = torch.nn.Parameter(self._shard(p.data).to(self.optimizer_dtype))
                shards.append(s)
                self._pairs.append((p, s))
                if self.use_grad_acc_hook:
                    p.main_grad = torch.zeros_like(p.data, dtype=torch.float32)

                    def hook(param):
                        param.main_grad.add_(param.grad.float())
                        param.grad = None
                    p.register_post_accumulate_grad_hook(hook)
            base_groups.append({**{k: v for k, v in g.items() if k != "params"}, "params": shards})
        self.base_optimizer = self.optimizer_class(base_groups, **self.opt_defaults)
        self.inited = True

    def zero_grad(self, set_to_none: bool = True):
        for p, _ in self._pairs:
            p.grad = None
            if self.use_grad_acc_hook:
                p.main_grad.zero_()

    def _get_grad(self, p):
        return p.main_grad if self.use_grad_acc_hook else p.grad

    def _clip_grad_norm(self, max_norm):          # overridden by the reference's NeuronZero1Optimizer
        sq = torch.zeros((), device=self._pairs[0][1].device)
        for _, s in self._pairs:
            if s.grad is not None:
                sq += s.grad.float().pow(2).sum()
        dist.all_reduce(sq)
        self._grad_norm = sq.sqrt()
        coef = torch.clamp(max_norm / (self._grad_norm + 1e-6), max=1.0)
        for _, s in self._pairs:
            if s.grad is not None:
                s.grad.mul_(coef)

    @torch.no_grad()
    def step(self, closure=None, **kwargs):
        n = self.local_world_size
        for p, s in self._pairs:
            g = self._get_grad(p)
            if g is None:
                continue
            g2 = self._pad(g.float() if self.higher_cc_precision else g)
            if n > 1:
                out = torch.empty_like(g2.chunk(n, 0)[0])
                dist.reduce_scatter_tensor(out, g2.contiguous(), group=self.pg)
                out.div_(n)
            else:
                out = g2
            s.grad = out.to(s.dtype)
        if self.grad_clipping:
            self._clip_grad_norm(max_norm=self.max_norm)
        for g, bg in zip(self.param_groups, self.base_optimizer.param_groups):
            for k, v in g.items():
                if k != "params":
                    bg[k] = v
        self.base_optimizer.step()
        if n == 1 and all(s.data.numel() == p.data.numel() for p, s in self._pairs):
            # one multi-tensor cast+copy for all parameters (torch_xla coalesces this too) instead of 2 kernels per parameter
            try:
                torch._foreach_copy_([p.data.view(-1) for p, _ in self._pairs], [s.data.view(-1) for _, s in self._pairs])
                for _, s in self._pairs:
                    s.grad = None
                return
            except Exception:          # noqa: BLE001 - any doubt: the plain per-parameter loop below is always valid
                pass
        fulls = []
        shs = [s.data.to(p.dtype).contiguous() for p, s in self._pairs]
        for (p, s), sh in zip(self._pairs, shs):
            fulls.append(torch.empty((sh.shape[0] * n, sh.shape[1]), dtype=p.dtype, device=p.device) if n > 1 else sh)
        if n > 1:
            try:
                cm = dist._coalescing_manager(group=self.pg, device=shs[0].device, async_ops=False)
            except Exception:  # noqa: BLE001
                import contextlib

                cm = contextlib.nullcontext()
            with cm:                                        # one NCCL group launch for all parameters (bucketed all-gather)
                for full, sh in zip(fulls, shs):
                    dist.all_gather_into_tensor(full, sh, group=self.pg)
        for (p, s), full in zip(self._pairs, fulls):
            rows = p.shape[0] if p.dim() > 0 else 1
            p.data.copy_(full[:rows].reshape(p.shape))
            s.grad = None

    def state_dict(self):
        return {"base_state": self.base_optimizer.state_dict(), "param_groups": []}


def _populate(module: types.ModuleType) -> None:
    name = module.__name__
    if name == "torch_xla.core.xla_model":
        module.REDUCE_SUM, module.REDUCE_MAX, module.REDUCE_MIN = "sum", "max", "min"
        module.all_reduce, module.all_gather = _xm_all_reduce, _xm_all_gather
        module.reduce_scatter, module.all_to_all = _xm_reduce_scatter, _xm_all_to_all
        module.mark_step = lambda *a, **k: None
        module.xla_device = _device
        module.rendezvous = lambda tag, *a, **k: (dist.barrier() if dist.is_initialized() else None)
        module.master_print = lambda *a, **k: print(*a, **k) if (not dist.is_initialized() or dist.get_rank() == 0) else None
        module.get_ordinal = lambda *a, **k: dist.get_rank() if dist.is_initialized() else 0
        module.xrt_world_size = lambda *a, **k: dist.get_world_size() if dist.is_initialized() else 1
        module.is_master_ordinal = lambda *a, **k: (not dist.is_initialized()) or dist.get_rank() == 0
        module.add_step_closure = lambda fn, args=(), **k: fn(*args)
        module.wait_device_ops = lambda *a, **k: torch.cuda.synchronize() if torch.cuda.is_available() else None
        module.get_local_ordinal = lambda *a, **k: int(__import__('os').environ.get('LOCAL_RANK', '0'))
        module.set_rng_state = lambda seed, *a, **k: torch.cuda.manual_seed(seed)
        module.get_rng_state = lambda *a, **k: torch.cuda.initial_seed()
    elif name == "torch_xla.runtime":
        module.world_size = lambda: dist.get_world_size() if dist.is_initialized() else 1
        module.global_ordinal = lambda: dist.get_rank() if dist.is_initialized() else 0
        module.local_ordinal = lambda: int(__import__('os').environ.get('LOCAL_RANK', '0'))
    elif name == "torch_xla.distributed.zero_redundancy_optimizer":
        module.ZeroRedundancyOptimizer = ZeroRedundancyOptimizer
    elif name == "torch_xla.utils.checkpoint":
        from torch.utils.checkpoint import checkpoint as _ckpt

        module.checkpoint = lambda fn, *a, **k: _ckpt(fn, *a, use_reentrant=False, **k)
    elif name in ("torch_neuronx.utils", "torch_neuronx.utils.utils"):
        module.get_platform_target = lambda *a, **k: "trn1"
        module.SUPPORTED_TYPES = ["trn1", "trn2", "inf2"]
    elif name == "torch_xla.core.xla_env_vars":
        import os as _os

        module.HOST_WORLD_SIZE = "XRT_HOST_WORLD_SIZE"
        _os.environ.setdefault("XRT_HOST_WORLD_SIZE", "1")          # single node
        module.WORLD_SIZE, module.ORDINAL, module.LOCAL_ORDINAL = "WORLD_SIZE", "RANK", "LOCAL_RANK"
    elif name == "torch_xla":
        module._XLAC = _StubModule("torch_xla._XLAC")
    elif name == "tenacity":
        module.retry = lambda *a, **k: (lambda f: f)


class _XlaDeviceRewrite(torch.overrides.TorchFunctionMode):
    """`device="xla"` literals in the reference (e.g. parallel_state.py:655 collective warm-up) → the real device."""

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        d = kwargs.get("device", None)
        if d is not None and (d == "xla" or (isinstance(d, torch.device) and d.type == "xla")):
            kwargs = dict(kwargs)
            kwargs["device"] = _device()
        return func(*args, **kwargs)


_INSTALLED = False


def install() -> None:
    global _INSTALLED
    if _INSTALLED:
        return
    sys.meta_path.insert(0, _Finder())
    # the reference builds its device groups with backend-specific pg_options; NCCL takes none
    _orig_new_group = dist.new_group

    _MESHES = {}

    def new_group(ranks=None, timeout=None, backend=None, pg_options=None, **kw):
        if backend in ("xla",):
            backend = None
        pg = _orig_new_group(ranks=ranks, backend=backend) if timeout is None else _orig_new_group(ranks=ranks, timeout=timeout, backend=backend)
        mesh = None
        if isinstance(pg_options, dict):
            mesh = pg_options.get("xla_pg_options", {}).get("mesh")
        if mesh is not None and pg is not None and not isinstance(pg, int):
            _MESHES[pg.group_name] = mesh           # torch_xla's "xla" backend exposes the replica mesh as pg._mesh
        return pg

    dist.new_group = new_group

    def _mesh_of(self):
        m = _MESHES.get(self.group_name)
        if m is None:
            m = [dist.get_process_group_ranks(self)] if self is not dist.group.WORLD else [list(range(dist.get_world_size()))]
        return m

    dist.ProcessGroup._mesh = property(_mesh_of)
    # `torch.classes.neuron.SPMDModel` appears in annotations evaluated at import time (reference trace/spmd.py:19)
    try:
        setattr(torch.classes, "neuron", _Dummy())
    except Exception:
        pass
    _INSTALLED = True


def device_rewrite():
    """Context manager: `device="xla"` literals → the CUDA device.  Entered only around the reference's parallel-state /
    model construction (the one literal on our path is parallel_state.py:655), NOT process-wide: a global
    TorchFunctionMode would tax every torch call of the timed step with a Python `__torch_function__` hop."""
    return _XlaDeviceRewrite()
