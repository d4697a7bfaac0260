"""Full (HF-style) ↔ sharded (TP/PP) model checkpoint conversion (reference ``scripts/checkpoint_converter.py:23-944``).

``CheckpointConverterBase`` shards a full state dict for every (tp, pp) coordinate using *name-pattern rules*
(which weights are column-parallel, row-parallel, fused QKV / gate-up, replicated) and writes either on-disk layout:
the trainer format ``<out>/model/dp_rank_00_tp_rank_XX_pp_rank_XX.pt`` or the legacy ``tp_rank_XX_pp_rank_XX/checkpoint.pt``.
The reverse direction gathers shards back into a full state dict.  Subclass and override ``get_partition_dim`` /
``pre_process_full_state_before_tp_conversion`` for other architectures (a Llama rule set is built in)."""
from __future__ import annotations

import argparse
import os
import re
from typing import Dict, List, Optional, Tuple

import torch

from ..parallel_layers.utils import create_local_weight, gather_full_weight


class CheckpointConverterBase:
    # (regex, partition_dim, stride); first match wins; no match → replicated
    rules: List[Tuple[str, int, int]] = [
        (r".*embed_tokens\.weight$", 0, 1), (r".*lm_head\.weight$", 0, 1),
        (r".*(q_proj|k_proj|v_proj)\.weight$", 0, 1), (r".*qkv_proj\.weight_(q|k|v)$", 0, 1),
        (r".*gate_up_proj\.weight$", 0, 2), (r".*(gate_proj|up_proj)\.weight$", 0, 1),
        (r".*(o_proj|down_proj)\.weight$", 1, 1),
    ]

    def get_partition_dim(self, name: str) -> Optional[Tuple[int, int]]:
        for pat, dim, stride in self.rules:
            if re.match(pat, name):
                return dim, stride
        return None

    # ---- hooks ------------------------------------------------------------------------
    def pre_process_full_state_before_tp_conversion(self, state: Dict[str, torch.Tensor], args) -> Dict[str, torch.Tensor]:
        """Coalesce q/k/v → fused qkv (with GQA KV replication) and gate/up → gate_up when requested."""
        out = dict(state)
        if getattr(args, "fuse_gate_up", False):
            for k in [k for k in state if k.endswith("gate_proj.weight")]:
                base = k[: -len("gate_proj.weight")]
                out[base + "gate_up_proj.weight"] = torch.cat([out.pop(k), out.pop(base + "up_proj.weight")], 0)
        if getattr(args, "qkv_linear", False):
            mult = getattr(args, "kv_size_multiplier", 1)
            for k in [k for k in state if k.endswith("q_proj.weight")]:
                base = k[: -len("q_proj.weight")]
                q, kk, v = out.pop(k), out.pop(base + "k_proj.weight"), out.pop(base + "v_proj.weight")
                if mult > 1:
                    kk, v = kk.repeat(mult, 1), v.repeat(mult, 1)
                out[base + "qkv_proj.weight_q"], out[base + "qkv_proj.weight_k"], out[base + "qkv_proj.weight_v"] = q, kk, v
        return out

    def post_process_full_state_after_tp_conversion(self, state: Dict[str, torch.Tensor], args) -> Dict[str, torch.Tensor]:
        return state

    # ---- full → sharded ---------------------------------------------------------------
    def convert_full_state_to_tp(self, full: Dict[str, torch.Tensor], tp_size: int, args=None) -> List[Dict[str, torch.Tensor]]:
        full = self.pre_process_full_state_before_tp_conversion(full, args)
        shards: List[Dict[str, torch.Tensor]] = [dict() for _ in range(tp_size)]
        fuse = bool(getattr(args, "fuse_qkv", False))
        for name, w in full.items():
            rule = self.get_partition_dim(name)
            for r in range(tp_size):
                if rule is None or not isinstance(w, torch.Tensor) or w.dim() == 0:
                    shards[r][name] = w
                else:
                    dim, stride = rule
                    shards[r][name] = create_local_weight(w, dim, w.shape[dim] // tp_size, stride, rank=r, world_size=tp_size).clone()
        if fuse:
            for r in range(tp_size):
                for k in [k for k in shards[r] if k.endswith("qkv_proj.weight_q")]:
                    base = k[: -len("weight_q")]
                    shards[r][base + "weight_qkv"] = torch.cat([shards[r].pop(base + "weight_q"), shards[r].pop(base + "weight_k"),
                                                                shards[r].pop(base + "weight_v")], 0)
        return shards

    def partition_pp(self, shard: Dict[str, torch.Tensor], pp_size: int, pp_rank: int, num_layers: int,
                     layer_regex: str = r".*layers\.(\d+)\..*") -> Dict[str, torch.Tensor]:
        if pp_size == 1:
            return shard
        from ..pipeline.partition import create_partitions

        cuts = create_partitions(num_layers, pp_size)
        bounds = [0] + [c + 1 for c in cuts] + [num_layers]
        lo, hi = bounds[pp_rank], bounds[pp_rank + 1]
        out = {}
        for k, v in shard.items():
            m = re.match(layer_regex, k)
            if m:
                if lo <= int(m.group(1)) < hi:
                    out[k] = v
            elif ("embed" in k and pp_rank == 0) or (("lm_head" in k or k.endswith("norm.weight")) and pp_rank == pp_size - 1):
                out[k] = v
        return out

    def save_sharded(self, shards: List[Dict[str, torch.Tensor]], output_dir: str, pp_size: int = 1, num_layers: int = 0,
                     legacy: bool = False, tag: Optional[str] = None) -> None:
        for tp, sd in enumerate(shards):
            for pp in range(pp_size):
                part = self.partition_pp(sd, pp_size, pp, num_layers)
                if legacy:
                    d = os.path.join(output_dir, f"tp_rank_{tp:02d}_pp_rank_{pp:02d}")
                    os.makedirs(d, exist_ok=True)
                    torch.save({"model": part}, os.path.join(d, "checkpoint.pt"))
                else:
                    d = os.path.join(output_dir, tag or "converted", "model")
                    os.makedirs(d, exist_ok=True)
                    torch.save(part, os.path.join(d, f"dp_rank_00_tp_rank_{tp:02d}_pp_rank_{pp:02d}.pt"))
        if not legacy:
            root = os.path.join(output_dir, tag or "converted")
            for marker in ("checkpoint", "done"):
                open(os.path.join(root, marker), "w").write("1")

    # ---- sharded → full ---------------------------------------------------------------
    def convert_tp_to_full_state(self, shards: List[Dict[str, torch.Tensor]], args=None) -> Dict[str, torch.Tensor]:
        full: Dict[str, torch.Tensor] = {}
        for name in shards[0]:
            rule = self.get_partition_dim(name)
            parts = [s[name] for s in shards]
            if rule is None or not isinstance(parts[0], torch.Tensor) or parts[0].dim() == 0:
                full[name] = parts[0]
            else:
                full[name] = gather_full_weight(parts, rule[0], rule[1])
        return self.post_process_full_state_after_tp_conversion(full, args)

    def load_sharded(self, input_dir: str, tp_size: int, pp_size: int = 1, legacy: bool = False, tag: Optional[str] = None):
        shards = []
        for tp in range(tp_size):
            merged: Dict[str, torch.Tensor] = {}
            for pp in range(pp_size):
                if legacy:
                    ck = torch.load(os.path.join(input_dir, f"tp_rank_{tp:02d}_pp_rank_{pp:02d}", "checkpoint.pt"),
                                    map_location="cpu", weights_only=False)
                    ck = ck.get("model", ck)
                else:
                    ck = torch.load(os.path.join(input_dir, tag or "converted", "model",
                                                 f"dp_rank_00_tp_rank_{tp:02d}_pp_rank_{pp:02d}.pt"), map_location="cpu", weights_only=False)
                merged.update(ck)
            shards.append(merged)
        return shards

    # ---- CLI ----------------------------------------------------------------------------
    def get_arg_parser(self) -> argparse.ArgumentParser:
        ap = argparse.ArgumentParser()
        ap.add_argument("--input_dir", required=True)
        ap.add_argument("--output_dir", required=True)
        ap.add_argument("--convert_from_full_state", action="store_true")
        ap.add_argument("--convert_to_full_state", action="store_true")
        ap.add_argument("--tp_size", type=int, default=1)
        ap.add_argument("--pp_size", type=int, default=1)
        ap.add_argument("--n_layers", type=int, default=0)
        ap.add_argument("--kv_size_multiplier", type=int, default=1)
        ap.add_argument("--qkv_linear", action="store_true")
        ap.add_argument("--fuse_qkv", action="store_true")
        ap.add_argument("--fuse_gate_up", action="store_true")
        ap.add_argument("--load_xser", action="store_true")
        ap.add_argument("--save_xser", action="store_true")
        ap.add_argument("--legacy_format", action="store_true")
        ap.add_argument("--tag", default=None)
        return ap

    def run(self, args) -> None:
        if args.convert_from_full_state:
            full = torch.load(args.input_dir, map_location="cpu", weights_only=False) if os.path.isfile(args.input_dir) else \
                torch.load(os.path.join(args.input_dir, "pytorch_model.bin"), map_location="cpu", weights_only=False)
            shards = self.convert_full_state_to_tp(full, args.tp_size, args)
            self.save_sharded(shards, args.output_dir, args.pp_size, args.n_layers, args.legacy_format, args.tag)
        elif args.convert_to_full_state:
            shards = self.load_sharded(args.input_dir, args.tp_size, args.pp_size, args.legacy_format, args.tag)
            os.makedirs(args.output_dir, exist_ok=True)
            torch.save(self.convert_tp_to_full_state(shards, args), os.path.join(args.output_dir, "pytorch_model.bin"))
        else:
            raise SystemExit("choose --convert_from_full_state or --convert_to_full_state")


def main(argv=None) -> int:
    c = CheckpointConverterBase()
    c.run(c.get_arg_parser().parse_args(argv))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
