"""Offline conversion of ZeRO-1 optimizer checkpoints (reference ``optimizer/convert_zero_checkpoints.py:179-231``,
console script ``nxd_convert_zero_checkpoints``): merge the per-dp-rank shards of one (tp, pp) coordinate into a
full state, or re-shard them for a different data-parallel degree.

The flat-buffer layout written by :class:`Zero1Optimizer` (``flat_layout`` / ``shape_info`` in the state dict) makes
this pure tensor concatenation + re-slicing: shard r of group g covers flat elements ``[r·n, (r+1)·n)``."""
from __future__ import annotations

import argparse
import os
import re
from typing import Any, Dict, List

import torch

_ALIGN = 128


def _load(path: str) -> Dict[str, Any]:
    return torch.load(path, map_location="cpu", weights_only=False)


def _dp_files(optim_dir: str, tp: int, pp: int) -> List[str]:
    pat = re.compile(rf"dp_rank_(\d+)_tp_rank_{tp:02d}_pp_rank_{pp:02d}\.pt$")
    found = sorted((int(m.group(1)), f) for f in os.listdir(optim_dir) if (m := pat.match(f)))
    return [os.path.join(optim_dir, f) for _, f in found]


def merge_shards(shards: List[Dict[str, Any]]) -> Dict[str, Any]:
    """Concatenate flat master weights and per-group optimizer state of all dp ranks (rank order)."""
    n_groups = len(shards[0]["flat_layout"])
    full = {"param_groups": shards[0]["param_groups"], "shape_info": shards[0]["shape_info"], "master": {}, "state": {}}
    for g in range(n_groups):
        full["master"][g] = torch.cat([s["sharded_master_weights"][g] for s in shards])
        st = {}
        for k, v in shards[0]["base_state"].get(g, {}).items():
            if isinstance(v, torch.Tensor) and v.dim() > 0:
                st[k] = torch.cat([s["base_state"][g][k] for s in shards])
            else:
                st[k] = v
        full["state"][g] = st
    return full


def full_to_named(full: Dict[str, Any]) -> Dict[str, Dict[str, torch.Tensor]]:
    """Flat → per-parameter tensors (index keys follow ``shape_info``)."""
    out: Dict[str, Dict[str, torch.Tensor]] = {"master": {}, "exp_avg": {}, "exp_avg_sq": {}}
    for idx, info in full["shape_info"].items():
        g, off, n, shape = info["group"], info["flat_offset"], info["numel"], info["shape"]
        out["master"][idx] = full["master"][g][off:off + n].view(shape)
        for k in ("exp_avg", "exp_avg_sq"):
            if k in full["state"].get(g, {}):
                out[k][idx] = full["state"][g][k][off:off + n].view(shape)
    return out


def reshard(full: Dict[str, Any], new_dp: int) -> List[Dict[str, Any]]:
    out = []
    n_groups = len(full["master"])
    for r in range(new_dp):
        sd = {"param_groups": full["param_groups"], "shape_info": full["shape_info"], "sharded_master_weights": {},
              "base_state": {}, "state": {}, "flat_layout": []}
        for g in range(n_groups):
            total = full["master"][g].numel()
            per = (total + new_dp - 1) // new_dp
            per = (per + _ALIGN - 1) // _ALIGN * _ALIGN

            def cut(t):
                pad = per * new_dp - t.numel()
                if pad > 0:
                    t = torch.cat([t, t.new_zeros(pad)])
                return t[r * per:(r + 1) * per].clone()

            sd["sharded_master_weights"][g] = cut(full["master"][g])
            sd["base_state"][g] = {k: (cut(v) if isinstance(v, torch.Tensor) and v.dim() > 0 else v)
                                   for k, v in full["state"][g].items()}
            sd["flat_layout"].append({"shard_numel": per, "total": per * new_dp, "world": new_dp, "rank": r})
        sd["state"] = sd["base_state"]
        out.append(sd)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# reference-named helpers (convert_zero_checkpoints.py:15-177) + CLI
# ---------------------------------------------------------------------------------------------------------------------
def _optim_dir(args) -> str:
    """``--input_dir`` may be the checkpoint tag directory (reference) or its ``optim`` sub-directory."""
    d = os.path.join(args.input_dir, "optim")
    return d if os.path.isdir(d) else args.input_dir


def is_full(args) -> bool:
    return any(f.startswith("full") for f in os.listdir(_optim_dir(args)))


def is_xser(args) -> bool:
    """Tensors stored out of line (``<file>.tensors/`` directories next to the ``.pt`` stubs)."""
    d = _optim_dir(args)
    return any(f.endswith(".tensors") and os.path.isdir(os.path.join(d, f)) for f in os.listdir(d))


def get_parallel_info(args):
    """``(dp, tp, pp)`` sizes inferred from the file names (``dp == 0`` for a full checkpoint)."""
    dp = tp = pp = -1
    for f in os.listdir(_optim_dir(args)):
        if not f.endswith(".pt"):
            continue
        nums = [int(n) for n in re.findall(r"\d+", f)]
        if f.startswith("full") and len(nums) >= 2:
            tp, pp = max(tp, nums[0]), max(pp, nums[1])
        elif len(nums) >= 3:
            dp, tp, pp = max(dp, nums[0]), max(tp, nums[1]), max(pp, nums[2])
    return dp + 1, tp + 1, pp + 1


def _load_any(path: str) -> Dict[str, Any]:
    """Plain ``torch.save`` file or an out-of-line (xser) one."""
    if os.path.isdir(path + ".tensors"):
        from ..trainer.checkpoint import _xser_load
        from ..trainer.checkpoint_storage import FilesysCheckpointStorage

        d, f = os.path.split(path)
        return _xser_load(FilesysCheckpointStorage(d), f, None, 1, 0)
    return _load(path)


def merge_optim_dp_checkpoints(args, tp_rank: int, pp_rank: int) -> Dict[str, Any]:
    files = _dp_files(_optim_dir(args), tp_rank, pp_rank)
    if not files:
        raise FileNotFoundError(f"no optimizer shards for tp={tp_rank} pp={pp_rank} under {_optim_dir(args)}")
    return merge_shards([_load_any(f) for f in files])


def split_and_save_ckpts(args, merged_ckpt: Dict[str, Any], tp_rank: int, pp_rank: int) -> None:
    out = os.path.join(args.output_dir, "optim") if getattr(args, "nested_output", False) else args.output_dir
    os.makedirs(out, exist_ok=True)
    for r, sd in enumerate(reshard(merged_ckpt, args.new_dp_size)):
        torch.save(sd, os.path.join(out, f"dp_rank_{r:02d}_tp_rank_{tp_rank:02d}_pp_rank_{pp_rank:02d}.pt"))


def _save_full(args, full: Dict[str, Any], tp: int, pp: int) -> None:
    out = os.path.join(args.output_dir, "optim") if getattr(args, "nested_output", False) else args.output_dir
    os.makedirs(out, exist_ok=True)
    torch.save({"full": full, "named": full_to_named(full)}, os.path.join(out, f"full_tp_rank_{tp:02d}_pp_rank_{pp:02d}.pt"))


def _sharded_to_full_task(args, tp_rank: int, pp_rank: int) -> None:
    _save_full(args, merge_optim_dp_checkpoints(args, tp_rank, pp_rank), tp_rank, pp_rank)


def _full_to_sharded_task(args, tp_rank: int, pp_rank: int) -> None:
    full = _load(os.path.join(_optim_dir(args), f"full_tp_rank_{tp_rank:02d}_pp_rank_{pp_rank:02d}.pt"))["full"]
    split_and_save_ckpts(args, full, tp_rank, pp_rank)


def _sharded_to_sharded_task(args, tp_rank: int, pp_rank: int) -> None:
    split_and_save_ckpts(args, merge_optim_dp_checkpoints(args, tp_rank, pp_rank), tp_rank, pp_rank)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="Convert ZeRO-1 optimizer checkpoints: sharded ↔ full, or re-shard for a new dp size")
    ap.add_argument("--input_dir", required=True, help="checkpoint tag directory (containing optim/) or the optim directory itself")
    ap.add_argument("--output_dir", required=True)
    ap.add_argument("--num_workers", type=int, default=1, help="(tp, pp) coordinates converted concurrently")
    ap.add_argument("--dp_size", type=int, default=None, help="target data-parallel size when converting to sharded")
    ap.add_argument("--tp_size", type=int, default=None, help="override the tp size inferred from the file names")
    ap.add_argument("--pp_size", type=int, default=None, help="override the pp size inferred from the file names")
    ap.add_argument("--convert_to_full", action="store_true")
    ap.add_argument("--convert_to_sharded", action="store_true")
    a, _ = ap.parse_known_args(argv)
    if not (a.convert_to_full or a.convert_to_sharded):
        ap.error("one of --convert_to_full / --convert_to_sharded is required")
    a.nested_output = os.path.isdir(os.path.join(a.input_dir, "optim"))       # mirror the input's directory convention
    a.is_xser, a.new_dp_size = is_xser(a), a.dp_size
    dp, tp, pp = get_parallel_info(a)
    a.dp_size, a.tp_size, a.pp_size = dp, a.tp_size or tp, a.pp_size or pp
    tasks = []
    if a.convert_to_full:
        if is_full(a):
            raise ValueError("Invalid inputs: convert full optim states to full optim states")
        tasks.append(_sharded_to_full_task)
    if a.convert_to_sharded:
        assert a.new_dp_size, "--dp_size is required with --convert_to_sharded"
        tasks.append(_full_to_sharded_task if is_full(a) else _sharded_to_sharded_task)
    coords = [(t, p) for p in range(a.pp_size) for t in range(a.tp_size)]
    for task in tasks:
        if a.num_workers > 1:
            import concurrent.futures

            with concurrent.futures.ThreadPoolExecutor(max_workers=a.num_workers) as ex:
                for f in [ex.submit(task, a, t, p) for t, p in coords]:
                    f.result()
        else:
            for t, p in coords:
                task(a, t, p)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
