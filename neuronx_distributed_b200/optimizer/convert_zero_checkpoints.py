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


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="Convert ZeRO-1 optimizer checkpoints: sharded → full / re-sharded")
    ap.add_argument("--input_dir", required=True, help="<ckpt>/<tag>/optim")
    ap.add_argument("--output_dir", required=True)
    ap.add_argument("--convert_to_full", action="store_true")
    ap.add_argument("--convert_to_sharded", action="store_true")
    ap.add_argument("--dp_size", type=int, default=None, help="target data-parallel size when re-sharding")
    ap.add_argument("--tp_size", type=int, default=1)
    ap.add_argument("--pp_size", type=int, default=1)
    a = ap.parse_args(argv)
    os.makedirs(a.output_dir, exist_ok=True)
    for pp in range(a.pp_size):
        for tp in range(a.tp_size):
            files = _dp_files(a.input_dir, tp, pp)
            if not files:
                raise FileNotFoundError(f"no optimizer shards for tp={tp} pp={pp} under {a.input_dir}")
            full = merge_shards([_load(f) for f in files])
            if a.convert_to_full:
                torch.save({"full": full, "named": full_to_named(full)},
                           os.path.join(a.output_dir, f"full_tp_rank_{tp:02d}_pp_rank_{pp:02d}.pt"))
            if a.convert_to_sharded:
                assert a.dp_size, "--dp_size is required with --convert_to_sharded"
                for r, sd in enumerate(reshard(full, a.dp_size)):
                    torch.save(sd, os.path.join(a.output_dir, f"dp_rank_{r:02d}_tp_rank_{tp:02d}_pp_rank_{pp:02d}.pt"))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
