"""``save_checkpoint`` / ``load_checkpoint`` with the reference's on-disk layout
(reference ``trainer/checkpoint.py:54-972``):

    <dir>/<tag>/checkpoint                                  (marker: save started)
    <dir>/<tag>/model/dp_rank_00_tp_rank_XX_pp_rank_XX.pt    (+ ``_ep_rank_XX`` when EP>1)
    <dir>/<tag>/optim/dp_rank_XX_tp_rank_XX_pp_rank_XX.pt    (per dp-rank only for ZeRO-1)
    <dir>/<tag>/scheduler.pt, user_content.pt
    <dir>/<tag>/done                                         (marker: save complete)

``use_xser=True`` stores every tensor in its own file ``<file>.tensors/tensor_<i>.pt`` next to a
reference file whose tensors are :class:`TensorReference` stubs plus ``<file>.info.pt`` with
``{tid: {dtype, shape}}`` — tensor files are spread over DP peers by greedy bin-packing so replicas
share the write bandwidth, and on load one rank per replica group reads each tensor and broadcasts it.
Other features: ``num_kept_ckpts`` rotation with clean-up of interrupted saves, ``async_save`` on a
1-thread executor with an atexit flush, ``avoid_saving_lower_precision_weights``, newest-complete-tag
auto-resume.
"""
from __future__ import annotations

import atexit
import os
import threading
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Any, Dict, List, Optional, Tuple

import contextlib

import torch
import torch.distributed as dist

from ..parallel_layers import parallel_state as ps
from ..utils.logger import get_logger
from .checkpoint_storage import BaseCheckpointStorage, create_checkpoint_storage

logger = get_logger()


class TensorReference:
    """Stub left in an xser reference file in place of a tensor.  Pickled under the NAME the reference's files use —
    ``torch_xla.utils.serialization.TensorReference`` (same single attribute ``tid``) — so xser checkpoints written by the
    reference load here and checkpoints written here load there (``NXD_XSER_NATIVE_PICKLE=1`` keeps this module's own name)."""

    def __init__(self, tid: int):
        self.tid = tid

    def __repr__(self) -> str:
        return f"TensorReference({self.tid})"


_XLA_SER = "torch_xla.utils.serialization"


@contextlib.contextmanager
def _torch_xla_pickle_names():
    """While active, ``torch_xla.utils.serialization.TensorReference`` resolves to :class:`TensorReference` (and that is the
    name pickle writes for it).  If the real torch_xla is importable nothing is shimmed except the class identity on load."""
    import sys
    import types

    created = []
    for name in ("torch_xla", "torch_xla.utils", _XLA_SER):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []                                  # behaves as a package for the sub-module lookups
            sys.modules[name] = m
            created.append(name)
    ser = sys.modules[_XLA_SER]
    prev = getattr(ser, "TensorReference", None)
    ser.TensorReference = TensorReference
    old_mod = TensorReference.__module__
    if os.environ.get("NXD_XSER_NATIVE_PICKLE", "0") != "1":
        TensorReference.__module__ = _XLA_SER
    try:
        yield
    finally:
        TensorReference.__module__ = old_mod
        if prev is not None:
            ser.TensorReference = prev
        elif _XLA_SER not in created:
            try:
                del ser.TensorReference
            except AttributeError:
                pass
        for name in created:
            sys.modules.pop(name, None)


class _RawBytes:
    """Already-serialised payload: storages write it verbatim (``save_object`` of anything else goes through torch.save)."""

    def __init__(self, data: bytes):
        self.data = data


def _get_path(prefix: str, tp: bool = True, pp: bool = True, dp: bool = False, ep: bool = False) -> str:
    """``prefix/dp_rank_XX[_ep_rank_XX]_tp_rank_XX_pp_rank_XX`` (reference :54-63)."""
    dp_rank = (ps.get_expert_data_parallel_rank() if ep else ps.get_data_parallel_rank()) if dp else 0
    name = f"dp_rank_{dp_rank:02d}"
    if ep:
        name += f"_ep_rank_{ps.get_expert_model_parallel_rank():02d}"
    name += f"_tp_rank_{ps.get_tensor_model_parallel_rank() if tp else 0:02d}"
    name += f"_pp_rank_{ps.get_pipeline_model_parallel_rank() if pp else 0:02d}"
    return os.path.join(prefix, name)


def _determine_remove_tags(storage: BaseCheckpointStorage, num_kept: Optional[int]) -> List[str]:
    """Tags to delete: everything incomplete that is older than the newest complete tag (interrupted
    saves / deletes) plus complete tags beyond ``num_kept`` (reference :66-98)."""
    if num_kept is None or num_kept < 0:
        return []
    tags = storage.list_checkpoint_tags()
    done = [t for t in tags if storage.is_checkpoint_tag_completed(t)]
    remove = []
    if done:
        newest_done = tags.index(done[-1])
        remove += [t for t in tags[:newest_done] if t not in done]
    if len(done) > num_kept:
        remove += done[: len(done) - num_kept]
    return remove


class CheckpointIOState:
    """Serialises checkpoint IO; optionally runs it on a background thread (reference :110-325)."""

    def __init__(self, async_save: bool = False):
        self.async_save = async_save
        self.executor = ThreadPoolExecutor(max_workers=1) if async_save else None
        self.save_future: Optional[Future] = None
        self.remove_future: Optional[Future] = None
        self.items: List[Tuple[Any, str]] = []
        self.storage: Optional[BaseCheckpointStorage] = None
        self.tag: Optional[str] = None
        self.lock = threading.Lock()

    # -- lifecycle -----------------------------------------------------------------------
    def begin(self, checkpoint_dir: BaseCheckpointStorage, tag: str) -> None:
        storage = checkpoint_dir              # reference argument name; it is the storage object of the checkpoint directory
        self.wait_all()                       # at most one save in flight
        self.storage, self.tag, self.items = storage, tag, []
        if _rank() == 0:
            storage.create_dir(tag)
            storage.save_text("1", os.path.join(tag, "checkpoint"))
        _barrier()

    def add_save_task(self, obj: Any, filename: str) -> None:
        self.items.append((obj, filename))

    def end(self, num_kept: Optional[int]) -> None:
        storage, tag, items = self.storage, self.tag, self.items
        assert storage is not None and tag is not None

        def write_all():
            for obj, filename in items:
                storage.save_object(obj, filename)

        if self.async_save:
            # tensors were already copied to host by the caller; hand the writes to the thread
            def job():
                write_all()
                return True

            self.save_future = self.executor.submit(job)
            self._finish_async(storage, tag, num_kept)
        else:
            write_all()
            _barrier()
            if _rank() == 0:
                storage.save_text("1", os.path.join(tag, "done"))
            _barrier()
            self.submit_remove(num_kept, async_remove=False)
            self.wait_remove()
        self.items = []

    # -- removal of old checkpoints (reference :246-315) ------------------------------------
    def submit_remove(self, num_kept: Optional[int], async_remove: bool = False, remove_tags: Optional[List[str]] = None) -> None:
        """Delete all but the newest ``num_kept`` completed checkpoints (or exactly ``remove_tags``).  Rank 0 first deletes the
        ``done`` markers — a crash in the middle of a deletion then leaves an *incomplete* tag, never a corrupt one that looks
        complete — and then the files, on the IO thread when ``async_remove``; directories go in :meth:`wait_remove`."""
        storage = self.storage
        assert storage is not None, "begin() must be called first"
        tags = list(remove_tags) if remove_tags else _determine_remove_tags(storage, num_kept)
        _barrier()
        if not tags:
            return
        self._remove_tags = tags
        if _rank() == 0:
            storage.remove_files([os.path.join(t, "done") for t in tags])

            def job():
                storage.remove_dirs(tags)
                return True

            if async_remove and self.executor is not None:
                self.remove_future = self.executor.submit(job)
            else:
                job()

    def wait_remove(self) -> None:
        if self.remove_future is not None:
            self.remove_future.result()
            self.remove_future = None
        self._remove_tags = None
        _barrier()                                # nobody lists tags while rank 0 may still be deleting

    def add_dcp_save_task(self, checkpoint_dir: BaseCheckpointStorage, state_dict: dict, optimizer, model, ckpt_path: str) -> None:
        """ZeRO-1 optimizer state through ``torch.distributed.checkpoint`` (re-shardable on load; reference :161-169)."""
        from ..optimizer import zero_dcp_utils as dcp_utils

        path = os.path.join(checkpoint_dir.dirname(), ckpt_path, "optim")
        inner = getattr(optimizer, "optimizer", optimizer)
        dcp_utils.save_optim_state_dict(path, state_dict, inner)

    def _finish_async(self, storage, tag, num_kept) -> None:
        prev = self.save_future

        def finalize():
            prev.result()
            return True

        # the "done" marker needs *all* ranks' writes → it is written at the next synchronisation point
        self._pending_done = (storage, tag, num_kept)

    def wait_save(self, async_remove: bool = False) -> None:
        """``async_remove`` (reference :198): also wait for a pending asynchronous removal of old checkpoints."""
        if async_remove and hasattr(self, "wait_remove"):
            self.wait_remove()
        if self.save_future is not None:
            self.save_future.result()
            self.save_future = None
        pend = getattr(self, "_pending_done", None)
        if pend is not None:
            storage, tag, num_kept = pend
            self._pending_done = None
            _barrier()
            if _rank() == 0:
                storage.save_text("1", os.path.join(tag, "done"))
            _barrier()
            self.submit_remove(num_kept, async_remove=False)
            self.wait_remove()

    def wait_all(self) -> None:
        self.wait_save()

    def finalize_shutdown(self) -> None:
        """atexit flush of an asynchronous save.  The process group may already be gone, so ranks cannot barrier: every
        rank drops a `<tag>/.async_done/rank_N` marker after ITS writes finished and rank 0 publishes `done` only once all
        `world` markers exist (bounded wait); otherwise the tag stays incomplete and auto-resume skips it — a tag is never
        marked complete while another rank was still writing (or had crashed)."""
        try:
            if self.save_future is not None:
                self.save_future.result()
                self.save_future = None
            pend = getattr(self, "_pending_done", None)
            if pend is not None:
                import time

                storage, tag, _ = pend
                self._pending_done = None
                world = int(os.environ.get("WORLD_SIZE", "1"))
                rank = int(os.environ.get("RANK", "0"))
                try:
                    if dist.is_initialized():
                        world, rank = dist.get_world_size(), dist.get_rank()
                except Exception:                     # noqa: BLE001 - group already destroyed
                    pass
                storage.save_text("1", os.path.join(tag, ".async_done", f"rank_{rank}"))
                if rank == 0:
                    deadline = time.time() + float(os.environ.get("NXD_ASYNC_CKPT_SHUTDOWN_WAIT_S", "120"))
                    while time.time() < deadline:
                        if all(storage.file_exists(os.path.join(tag, ".async_done", f"rank_{r}")) for r in range(world)):
                            storage.save_text("1", os.path.join(tag, "done"))
                            break
                        time.sleep(0.2)
                    else:
                        logger.warning("async checkpoint %s left incomplete at shutdown: not every rank finished writing", tag)
        finally:
            if self.executor is not None:
                self.executor.shutdown(wait=True)


g_iostate: Optional[CheckpointIOState] = None


def _rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def _barrier() -> None:
    if dist.is_initialized() and os.environ.get("NXD_SKIP_RENDEZVOUS", "0") != "1":
        dist.barrier()


def _to_cpu(obj: Any) -> Any:
    if isinstance(obj, torch.Tensor):
        return obj.detach().to("cpu", copy=True) if obj.device.type != "cpu" else obj.detach().clone()
    if isinstance(obj, dict):
        return {k: _to_cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cpu(v) for v in obj)
    return obj


# ---------------------------------------------------------------------------- xser format
def _flatten_tensors(obj: Any, out: List[torch.Tensor]) -> Any:
    if isinstance(obj, torch.Tensor):
        out.append(obj)
        return TensorReference(len(out) - 1)
    if isinstance(obj, dict):
        return {k: _flatten_tensors(v, out) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_flatten_tensors(v, out) for v in obj)
    return obj


def _unflatten_tensors(obj: Any, tensors: Dict[int, torch.Tensor]) -> Any:
    if isinstance(obj, TensorReference):
        return tensors[obj.tid]
    if isinstance(obj, dict):
        return {k: _unflatten_tensors(v, tensors) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_unflatten_tensors(v, tensors) for v in obj)
    return obj


def _assign_tensors_to_bins(tensors: List[torch.Tensor], bin_count: int) -> List[List[int]]:
    """Greedy largest-first bin packing of tensor indices by byte size (reference :443-474)."""
    order = sorted(range(len(tensors)), key=lambda i: -tensors[i].numel() * tensors[i].element_size())
    bins: List[List[int]] = [[] for _ in range(bin_count)]
    load = [0] * bin_count
    for i in order:
        b = load.index(min(load))
        bins[b].append(i)
        load[b] += tensors[i].numel() * tensors[i].element_size()
    return bins


def _xser_tasks(state: Any, path: str, writers: int, writer_rank: int, iostate: CheckpointIOState) -> None:
    tensors: List[torch.Tensor] = []
    ref = _flatten_tensors(state, tensors)
    bins = _assign_tensors_to_bins(tensors, max(1, writers))
    for tid in bins[writer_rank % max(1, writers)]:
        iostate.add_save_task(_to_cpu(tensors[tid]), os.path.join(path + ".tensors", f"tensor_{tid}.pt"))
    if writer_rank == 0:
        import io as _io

        buf = _io.BytesIO()
        with _torch_xla_pickle_names():            # serialised NOW (tiny object) under the reference's class name
            torch.save(ref, buf)
        iostate.add_save_task(_RawBytes(buf.getvalue()), path)
        info = {i: {"dtype": t.dtype, "shape": tuple(t.shape),
                    "expert_model_parallel": bool(getattr(t, "expert_model_parallel", False))} for i, t in enumerate(tensors)}
        iostate.add_save_task(info, path + ".info.pt")


def _xser_load(storage: BaseCheckpointStorage, path: str, group, readers: int, reader_rank: int) -> Any:
    """One reader per replica group loads each tensor file and broadcasts it (reference :347-432)."""
    with _torch_xla_pickle_names():                # files written by the reference name torch_xla's TensorReference
        ref = storage.load_object(path, map_location="cpu")
    info = storage.load_object(path + ".info.pt", map_location="cpu")
    tensors: Dict[int, torch.Tensor] = {}
    from ..utils import get_device

    dev = get_device()
    for tid in sorted(info):
        meta = info[tid]
        if readers <= 1:
            tensors[tid] = storage.load_object(os.path.join(path + ".tensors", f"tensor_{tid}.pt"), map_location="cpu")
            continue
        owner = tid % readers
        if reader_rank == owner:
            t = storage.load_object(os.path.join(path + ".tensors", f"tensor_{tid}.pt"), map_location="cpu").to(dev)
        else:
            t = torch.empty(meta["shape"], dtype=meta["dtype"], device=dev)
        dist.broadcast(t, src=dist.get_global_rank(group, owner), group=group)
        tensors[tid] = t
    return _unflatten_tensors(ref, tensors)


# ---------------------------------------------------------------------------- public API
def has_checkpoint(checkpoint_dir_str: str) -> bool:
    return len(create_checkpoint_storage(checkpoint_dir_str).list_completed_checkpoint_tags()) > 0


def _zero1_states_have_master_weights(sd: Any) -> bool:
    return isinstance(sd, dict) and "sharded_master_weights" in sd


def save_checkpoint(
    checkpoint_dir_str: str,
    tag: str,
    model: Any = None,
    optimizer: Any = None,
    scheduler: Any = None,
    user_content: Any = None,
    num_workers: int = 8,
    use_xser: bool = False,
    num_kept_ckpts: Optional[int] = None,
    async_save: bool = False,
    zero1_optimizer: bool = False,
    use_zero1_dcp: bool = False,
    avoid_saving_lower_precision_weights: bool = False,
) -> None:
    global g_iostate
    assert dist.is_initialized(), "Only support distributed training mode."
    storage = create_checkpoint_storage(checkpoint_dir_str)
    if g_iostate is None or g_iostate.async_save != async_save:
        if g_iostate is not None:
            g_iostate.wait_all()
        g_iostate = CheckpointIOState(async_save)
        atexit.register(g_iostate.finalize_shutdown)
    io = g_iostate
    io.begin(storage, tag)
    ep = ps.get_expert_model_parallel_size() > 1
    dp_rank, dp_size = ps.get_data_parallel_rank(), ps.get_data_parallel_size()
    # Under expert parallelism a model / non-ZeRO optimizer file is keyed by ep_rank and replicated only over the EXPERT
    # data-parallel group (reference checkpoint.py:443-474, 545-585: edp_size / edp_rank): that group spreads the xser tensor
    # bins, and exactly one member (edp_rank 0) writes the plain file, the ref file and the info file.
    rep_rank, rep_size = (ps.get_expert_data_parallel_rank(), ps.get_expert_data_parallel_size()) if ep else (dp_rank, dp_size)

    opt_has_master = False
    if optimizer is not None:
        sd = optimizer.state_dict() if hasattr(optimizer, "state_dict") else optimizer
        from ..optimizer.zero_redundancy_optimizer import NeuronEPZero1Optimizer, Zero1Optimizer

        inner = getattr(optimizer, "optimizer", optimizer)
        zero1 = zero1_optimizer or isinstance(inner, (Zero1Optimizer, NeuronEPZero1Optimizer))
        opt_has_master = _zero1_states_have_master_weights(sd)
        # ZeRO-1 state is per dp rank (every rank owns a different shard); a plain optimizer's state is replicated over the
        # (expert-)data-parallel group and differs between EP ranks (reference `_get_path("optim", ep=True)`)
        path = _get_path(os.path.join(tag, "optim"), dp=zero1, ep=(ep and not zero1))
        if use_zero1_dcp and zero1:
            from ..optimizer import zero_dcp_utils

            zero_dcp_utils.save_optim_state_dict(os.path.join(storage.dirname(), tag, "optim"), sd, inner)
        elif use_xser:
            writers, wrank = (1, 0) if zero1 else (rep_size, rep_rank)
            _xser_tasks(sd, path, writers, wrank, io)
        elif zero1 or rep_rank == 0:
            io.add_save_task(_to_cpu(sd), path + ".pt")

    if model is not None:
        skip_weights = avoid_saving_lower_precision_weights and opt_has_master
        if skip_weights:
            logger.info("model weights are not saved: optimizer checkpoint holds the fp32 master weights")
        else:
            sd = model.state_dict() if hasattr(model, "state_dict") else model
            path = _get_path(os.path.join(tag, "model"), dp=False, ep=ep)
            if use_xser:
                _xser_tasks(sd, path, rep_size, rep_rank, io)
            elif rep_rank == 0:
                io.add_save_task(_to_cpu(sd), path + ".pt")

    if _rank() == 0:
        if scheduler is not None:
            io.add_save_task(scheduler.state_dict() if hasattr(scheduler, "state_dict") else scheduler,
                             os.path.join(tag, "scheduler.pt"))
        if user_content is not None:
            io.add_save_task(user_content, os.path.join(tag, "user_content.pt"))
    io.end(num_kept_ckpts)


def load_checkpoint(
    path: str,
    tag: Optional[str] = None,
    model: Optional[torch.nn.Module] = None,
    optimizer: Optional[torch.optim.Optimizer] = None,
    scheduler: Any = None,
    num_workers: int = 8,
    strict: bool = True,
    use_zero1_dcp: bool = False,
) -> Any:
    assert dist.is_initialized(), "Only support distributed training mode."
    global g_iostate
    if g_iostate is not None:
        g_iostate.wait_all()
    storage = create_checkpoint_storage(path)
    if tag is None:
        tags = storage.list_completed_checkpoint_tags()
        if not tags:
            raise RuntimeError(f"no completed checkpoint under {path}")
        tag = tags[-1]
    ep = ps.get_expert_model_parallel_size() > 1
    model_dir, optim_dir = os.path.join(tag, "model"), os.path.join(tag, "optim")
    use_xser = storage.is_checkpoint_xser(model_dir) or storage.is_checkpoint_xser(optim_dir)
    dp_group, dp_size, dp_rank = ps.get_data_parallel_group(), ps.get_data_parallel_size(), ps.get_data_parallel_rank()
    if ep:      # replicas of an ep_rank-keyed file live in the expert-data-parallel group (see save_checkpoint)
        rep_group, rep_size, rep_rank = (ps.get_expert_data_parallel_group(), ps.get_expert_data_parallel_size(),
                                         ps.get_expert_data_parallel_rank())
    else:
        rep_group, rep_size, rep_rank = dp_group, dp_size, dp_rank

    if model is not None and storage.dir_exists(model_dir):
        p = _get_path(model_dir, dp=False, ep=ep)
        sd = _xser_load(storage, p, rep_group, rep_size, rep_rank) if use_xser else storage.load_object(p + ".pt", "cpu")
        model.load_state_dict(sd, strict=strict)
    if optimizer is not None:
        from ..optimizer.zero_redundancy_optimizer import NeuronEPZero1Optimizer, Zero1Optimizer

        inner = getattr(optimizer, "optimizer", optimizer)
        zero1 = isinstance(inner, (Zero1Optimizer, NeuronEPZero1Optimizer))
        p = _get_path(optim_dir, dp=zero1, ep=(ep and not zero1))
        if use_zero1_dcp and zero1:
            from ..optimizer import zero_dcp_utils

            sd = zero_dcp_utils.load_optim_state_dict(os.path.join(storage.dirname(), optim_dir), inner)
        elif use_xser:
            sd = _xser_load(storage, p, rep_group, 1 if zero1 else rep_size, 0 if zero1 else rep_rank)
        else:
            sd = storage.load_object(p + ".pt", "cpu")
        optimizer.load_state_dict(sd)
        if model is not None and not storage.dir_exists(model_dir) and zero1:
            pass  # weights were restored from the fp32 master shards by the optimizer's load_state_dict
    if scheduler is not None and storage.file_exists(os.path.join(tag, "scheduler.pt")):
        scheduler.load_state_dict(storage.load_object(os.path.join(tag, "scheduler.pt"), "cpu"))
    user_content = None
    if storage.file_exists(os.path.join(tag, "user_content.pt")):
        user_content = storage.load_object(os.path.join(tag, "user_content.pt"), "cpu")
    _barrier()
    return user_content


def finalize_checkpoint() -> None:
    if g_iostate is not None:
        g_iostate.wait_all()
