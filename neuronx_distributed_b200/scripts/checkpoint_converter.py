"""Full (HF / Megatron style) ↔ sharded (TP × PP × EP) model checkpoint conversion
(reference ``scripts/checkpoint_converter.py:23-944``, console script ``nxd_convert_checkpoint``).

``CheckpointConverterBase`` is meant to be subclassed per architecture: the class attributes (partition dims, layer name
pattern, config attribute map) and the small predicates (``is_qkv_weight``, ``get_partition_dim``, ``get_weight_key``,
``rename_keys_for_megatron`` …) are the override points, the driver methods (``convert_from_full_state``,
``convert_to_full_state``, ``convert_from_xser`` / ``convert_to_xser``) and the CLI stay the same.

Sharding is rule driven (``rules``: regex → partition dim, stride) instead of the reference's if/elif chain, and one pass
produces every (tp, pp, ep) coordinate.  What the rules cannot express is handled explicitly:

* **GQA with replicated KV heads** (``--qkv_linear --kv_size_multiplier m``): K/V are replicated ``m`` times in the layout
  of ``GQAQKVColumnParallelLinear`` (``tile``: ``K0…Kn K0…Kn``, ``adjacent``: ``K0 K0 … K1 K1``) and — for ``tile`` — the Q
  heads (and the matching ``o_proj`` columns) are permuted so that every TP rank's Q heads meet *their* KV head
  (``gqa_q_head_permutation``; the reference does the same reshuffle inline, :560-640);
* fused ``gate_up_proj`` (stride-2 interleave), fused / non-fused QKV parameter names, expert-parallel slicing of
  ``expert_mlps`` weights on dim 0, pipeline stages from layer indices (virtual stages included), Megatron key names.

On-disk layouts read and written: trainer format ``<dir>/model/dp_rank_00[_ep_rank_XX]_tp_rank_XX_pp_rank_XX.pt`` (plain or
out-of-line "xser") and the legacy ``tp_rank_XX_pp_rank_XX/checkpoint.pt``.
"""
from __future__ import annotations

import argparse
import json
import os
import re
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from ..parallel_layers.utils import create_local_weight, gather_full_weight
from ..pipeline.partition import create_partitions, stage_to_pipeline_parallel_rank


def gqa_q_head_permutation(q_heads: int, kv_heads: int, kv_size_multiplier: int, layout: str = "tile") -> List[int]:
    """Order in which the Q heads must be stored so that contiguous TP sharding pairs them with the replicated KV heads.

    Replicated KV slot ``j`` holds KV head ``j % kv_heads`` (``tile``) or ``j // m`` (``adjacent``), copy ``c``; it serves the
    ``c``-th ``1/m`` chunk of that KV head's query group."""
    group = q_heads // kv_heads
    assert q_heads % kv_heads == 0 and group % kv_size_multiplier == 0, \
        f"q_heads/kv_heads ({group}) must be a multiple of kv_size_multiplier ({kv_size_multiplier})"
    chunk = group // kv_size_multiplier
    order: List[int] = []
    for j in range(kv_heads * kv_size_multiplier):
        kv, c = (j % kv_heads, j // kv_heads) if layout == "tile" else (j // kv_size_multiplier, j % kv_size_multiplier)
        order += list(range(kv * group + c * chunk, kv * group + (c + 1) * chunk))
    return order


def _permute_heads(w: torch.Tensor, order: Sequence[int], head_dim: int, dim: int) -> torch.Tensor:
    shape = list(w.shape)
    n = shape[dim] // head_dim
    v = w.reshape(*shape[:dim], n, head_dim, *shape[dim + 1:])
    return v.index_select(dim, torch.as_tensor(list(order))).reshape(shape)


class CheckpointConverterBase:
    # ---- per-architecture knobs (reference :24-38) -------------------------------------------------------------------
    attribute_map: Dict[str, str] = {}
    embedding_partition_dim = 0
    qkv_partition_dim = 0
    gate_up_proj_partition_dim = 0
    down_proj_partition_dim = 1
    o_proj_partition_dim = 1
    layer_name = "layers"
    layer_name_pattern = r"^(model\.layers\.\d+)"

    # (regex, partition_dim attribute or int, stride); first match wins; no match → replicated
    rules: List[Tuple[str, Any, int]] = [
        (r".*(embed_tokens|lm_head|wte)\.weight$", "embedding_partition_dim", 1),
        (r".*(q_proj|k_proj|v_proj)\.(weight|bias)$", "qkv_partition_dim", 1),
        (r".*qkv_proj\.(weight|bias)_(q|k|v)$", "qkv_partition_dim", 1),
        (r".*query_key_value\.(weight|bias)$", "qkv_partition_dim", 1),
        (r".*gate_up_proj\.weight$", "gate_up_proj_partition_dim", 2),
        (r".*(gate_proj|up_proj)\.weight$", "gate_up_proj_partition_dim", 1),
        (r".*down_proj\.weight$", "down_proj_partition_dim", 1),
        (r".*o_proj\.weight$", "o_proj_partition_dim", 1),
    ]

    # ---- small predicates / name maps --------------------------------------------------------------------------------
    def _get_config_value(self, config: Dict[str, Any], key: str):
        """``config[key]``, or the (possibly nested, dotted) name ``attribute_map`` gives it in this architecture."""
        if key in config:
            return config[key]
        mapped = self.attribute_map.get(key)
        if mapped is not None:
            cur: Any = config
            for part in mapped.split("."):
                cur = cur.get(part) if isinstance(cur, dict) else None
                if cur is None:
                    break
            if cur is not None:
                return cur
            if mapped in config:
                return config[mapped]
        raise KeyError(f"Could not find {key} or its mapped name in config")

    def get_partition_rule(self, name: str) -> Optional[Tuple[int, int]]:
        for pat, dim, stride in self.rules:
            if re.match(pat, name):
                return (getattr(self, dim) if isinstance(dim, str) else dim), stride
        return None

    def get_partition_dim(self, name: str) -> int:
        rule = self.get_partition_rule(name if name.endswith(("weight", "bias")) or "weight_" in name else name + ".weight")
        if rule is None:
            raise AssertionError(f"Unknown partition_dim for {name}")
        return rule[0]

    def is_qkv_weight(self, name: str) -> bool:
        return any(t in name for t in ("q_proj", "k_proj", "v_proj", "qkv_proj", "query_key_value"))

    def get_fused_qkv_key(self) -> str:
        return "qkv_proj.weight_qkv"

    def get_hf_to_nxd_model_keys(self, qkv_linear: bool = True, is_gqa: bool = True):
        if qkv_linear:
            m = {"q_proj.weight": "qkv_proj.weight_q", "k_proj.weight": "qkv_proj.weight_k", "v_proj.weight": "qkv_proj.weight_v"}
        elif is_gqa:
            m = {k: k for k in ("q_proj.weight", "k_proj.weight", "v_proj.weight")}
        else:
            m = {k: "qkv_proj.weight" for k in ("q_proj.weight", "k_proj.weight", "v_proj.weight")}
        return m, {v: k for k, v in m.items()}

    def get_weight_key(self, keys_hf_to_nxd, keys_nxd_to_hf, name: str, hf_to_nxd: bool) -> str:
        """Rename the trailing ``<module>.<param>`` of a QKV entry between the HF and the NxD spelling."""
        if not self.is_qkv_weight(name):
            return name
        table = keys_hf_to_nxd if hf_to_nxd else keys_nxd_to_hf
        head, tail = name.split(".")[:-2], ".".join(name.split(".")[-2:])
        return ".".join(head + [table.get(tail, tail)])

    # Megatron-LM naming (reference :171-266): same tensors, different names / fused QKV
    _MEGATRON = [("model.embed_tokens", "language_model.embedding.word_embeddings"), ("model.layers", "language_model.encoder.layers"),
                 ("self_attn.o_proj", "self_attention.dense"), ("self_attn.qkv_proj", "self_attention.query_key_value"),
                 ("self_attn", "self_attention"), ("mlp.gate_up_proj", "mlp.dense_h_to_4h"), ("mlp.down_proj", "mlp.dense_4h_to_h"),
                 ("model.norm", "language_model.encoder.final_layernorm"), ("lm_head", "language_model.output_layer")]

    def rename_keys_for_megatron(self, key: str, model_style: str, hf_to_nxdt: bool = False) -> str:
        if model_style != "megatron":
            return key
        for hf, mg in self._MEGATRON:
            src, dst = (hf, mg) if hf_to_nxdt else (mg, hf)
            if src in key:
                key = key.replace(src, dst)
        return key

    def is_q_or_o_for_megatron(self, args, name: str) -> bool:
        return getattr(args, "model_style", "hf") == "megatron" and ("weight_q" in name or "o_proj" in name)

    def modify_qkv_for_megatron(self, partial_state: Dict[str, torch.Tensor], args) -> None:
        """Megatron keeps one ``query_key_value`` tensor: concatenate this rank's q, k, v shards under that name."""
        if getattr(args, "model_style", "hf") != "megatron":
            return
        for k in [k for k in partial_state if k.endswith("query_key_value.weight_q")]:
            base = k[: -len("weight_q")]
            partial_state[base + "weight"] = torch.cat([partial_state.pop(base + "weight_q"), partial_state.pop(base + "weight_k"),
                                                        partial_state.pop(base + "weight_v")], self.qkv_partition_dim)

    @staticmethod
    def find_size(state: Dict[str, Any]) -> int:
        """Bytes held by the tensors of a (nested) state dict."""
        return sum(v.numel() * v.element_size() if isinstance(v, torch.Tensor) else
                   (CheckpointConverterBase.find_size(v) if isinstance(v, dict) else 0) for v in state.values())

    def coalesce_qkv(self, state_dict: Dict[str, torch.Tensor], config: Dict[str, Any], tp_degree: int) -> Dict[str, torch.Tensor]:
        """MHA only: ``q/k/v_proj.weight`` → one ``qkv_proj.weight`` laid out rank-major (``[q_r; k_r; v_r]`` per rank) so
        that plain dim-0 sharding hands every rank its own q, k, v rows."""
        for i in range(self._get_config_value(config, "num_hidden_layers")):
            p = f"model.{self.layer_name}.{i}.self_attn."
            q, k, v = (state_dict.pop(p + n + "_proj.weight") for n in "qkv")
            per = [torch.cat([t.chunk(tp_degree, 0)[r] for t in (q, k, v)], self.qkv_partition_dim) for r in range(tp_degree)]
            state_dict[p + "qkv_proj.weight"] = torch.cat(per, self.qkv_partition_dim)
        return state_dict

    def convert_partial_state_to_fused_qkv(self, partial_state: Dict[str, torch.Tensor], keys_nxd_to_hf=None, n_layers: int = 0):
        """``weight_q / weight_k / weight_v`` of one rank → ``weight_qkv`` (``GQAQKVColumnParallelLinear(fuse_qkv=True)``)."""
        for k in [k for k in partial_state if k.endswith("qkv_proj.weight_q")]:
            base = k[: -len("weight_q")]
            partial_state[base + "weight_qkv"] = torch.cat([partial_state.pop(base + "weight_q"), partial_state.pop(base + "weight_k"),
                                                            partial_state.pop(base + "weight_v")], self.qkv_partition_dim)
        return partial_state

    def convert_partial_state_to_non_fused_qkv(self, partial_state: Dict[str, torch.Tensor], keys_nxd_to_hf=None, kv_size_multiplier: int = 1,
                                               config: Optional[Dict[str, Any]] = None, tp_size: int = 1):
        """Inverse of :meth:`convert_partial_state_to_fused_qkv`; the split sizes follow the head counts in ``config``."""
        for k in [k for k in partial_state if k.endswith("qkv_proj.weight_qkv")]:
            base, w = k[: -len("weight_qkv")], partial_state.pop(k)
            if config is not None:
                qh, kvh = self._get_config_value(config, "num_attention_heads"), self._get_config_value(config, "num_key_value_heads")
                hd = self._get_config_value(config, "hidden_size") // qh
                sizes = [qh * hd // tp_size, kvh * kv_size_multiplier * hd // tp_size, kvh * kv_size_multiplier * hd // tp_size]
            else:
                sizes = [w.shape[0] // 3] * 3
            partial_state[base + "weight_q"], partial_state[base + "weight_k"], partial_state[base + "weight_v"] = \
                torch.split(w, sizes, self.qkv_partition_dim)
        return partial_state

    # ---- hooks -------------------------------------------------------------------------------------------------------
    def pre_process_full_state_before_tp_conversion(self, full_state: Dict[str, torch.Tensor], args) -> Dict[str, torch.Tensor]:
        """Bring an HF-style full state into the parameter naming of the target model: gate/up → fused ``gate_up_proj``
        (``--fuse_gate_up``), q/k/v → ``qkv_proj.weight_{q,k,v}`` with KV replication and Q / o_proj head permutation
        (``--qkv_linear``).  Subclasses may override or extend."""
        state = full_state      # reference parameter names in the signature
        out = dict(state)
        if getattr(args, "fuse_gate_up", False):
            for k in [k for k in state if k.endswith("gate_proj.weight")]:
                base = k[: -len("gate_proj.weight")]
                out[base + "gate_up_proj.weight"] = torch.cat([out.pop(k), out.pop(base + "up_proj.weight")], 0)
        if getattr(args, "qkv_linear", False):
            mult = getattr(args, "kv_size_multiplier", 1)
            layout = getattr(args, "kv_replication_layout", "tile")
            q_heads = getattr(args, "_q_heads", None)
            kv_heads = getattr(args, "_kv_heads", None)
            for k in [k for k in state if k.endswith("q_proj.weight")]:
                base = k[: -len("q_proj.weight")]
                q, kk, v = out.pop(k), out.pop(base + "k_proj.weight"), out.pop(base + "v_proj.weight")
                if mult > 1:
                    from ..modules.qkv_linear import replicate_kv

                    hd = q.shape[0] // q_heads if q_heads else None
                    if layout != "tile" and hd is None:
                        raise ValueError("the adjacent KV layout needs the head counts: pass --config")
                    kk, v = replicate_kv(kk, mult, hd, layout), replicate_kv(v, mult, hd, layout)
                    if q_heads and kv_heads and layout == "tile":
                        order = gqa_q_head_permutation(q_heads, kv_heads, mult, layout)
                        q = _permute_heads(q, order, hd, 0)
                        okey = base + "o_proj.weight"
                        if okey in out:
                            out[okey] = _permute_heads(out[okey], order, hd, 1)
                out[base + "qkv_proj.weight_q"], out[base + "qkv_proj.weight_k"], out[base + "qkv_proj.weight_v"] = q, kk, v
        return out

    def post_process_full_state_after_tp_conversion(self, full_state: Dict[str, torch.Tensor], args) -> Dict[str, torch.Tensor]:
        """Inverse of the pre-processing for the sharded → full direction (un-permute, de-replicate, HF names)."""
        state = full_state      # reference parameter names in the signature
        if not getattr(args, "qkv_linear", False):
            return state
        out = dict(state)
        mult, layout = getattr(args, "kv_size_multiplier", 1), getattr(args, "kv_replication_layout", "tile")
        q_heads, kv_heads = getattr(args, "_q_heads", None), getattr(args, "_kv_heads", None)
        for k in [k for k in state if k.endswith("qkv_proj.weight_q")]:
            base = k[: -len("qkv_proj.weight_q")]
            q, kk, v = out.pop(k), out.pop(base + "qkv_proj.weight_k"), out.pop(base + "qkv_proj.weight_v")
            if mult > 1:
                hd = q.shape[0] // q_heads if q_heads else None
                if layout == "tile":
                    kk, v = kk[: kk.shape[0] // mult], v[: v.shape[0] // mult]
                    if q_heads and kv_heads:
                        order = gqa_q_head_permutation(q_heads, kv_heads, mult, layout)
                        inv = sorted(range(len(order)), key=order.__getitem__)
                        q = _permute_heads(q, inv, hd, 0)
                        if base + "o_proj.weight" in out:
                            out[base + "o_proj.weight"] = _permute_heads(out[base + "o_proj.weight"], inv, hd, 1)
                else:
                    sel = torch.arange(0, kk.shape[0] // hd, mult)
                    kk, v = (t.reshape(-1, hd, t.shape[1])[sel].reshape(-1, t.shape[1]) for t in (kk, v))
            out[base + "q_proj.weight"], out[base + "k_proj.weight"], out[base + "v_proj.weight"] = q, kk, v
        return out

    # ---- full → sharded ----------------------------------------------------------------------------------------------
    def shard_full_state(self, full: Dict[str, torch.Tensor], tp_size: int, args=None) -> List[Dict[str, torch.Tensor]]:
        """All TP shards of a (pre-processed) full state."""
        shards: List[Dict[str, torch.Tensor]] = [dict() for _ in range(tp_size)]
        for name, w in full.items():
            rule = self.get_partition_rule(name)
            for r in range(tp_size):
                if rule is None or not isinstance(w, torch.Tensor) or w.dim() == 0:
                    shards[r][name] = w
                else:
                    dim, stride = rule
                    shards[r][name] = create_local_weight(w, dim, w.shape[dim] // tp_size, stride, rank=r, world_size=tp_size).clone()
        if bool(getattr(args, "fuse_qkv", False)):
            for r in range(tp_size):
                self.convert_partial_state_to_fused_qkv(shards[r])
        return shards

    def _stage_of(self, name: str, partitions: Sequence[str], pp_size: int) -> Optional[int]:
        """PP rank owning a layer parameter (``None``: not a layer parameter)."""
        m = re.match(self.layer_name_pattern, name) or re.match(rf".*{self.layer_name}\.(\d+)\..*", name)
        if m is None:
            return None
        layer_idx = int(re.findall(r"\d+", m.group(1) if m.lastindex else name)[-1]) if re.match(self.layer_name_pattern, name) \
            else int(m.group(1))
        stage = len(partitions)
        for s, cut in enumerate(partitions):
            if layer_idx <= int(re.findall(r"\d+", cut)[-1]):
                stage = s
                break
        return stage_to_pipeline_parallel_rank(stage, pp_size)

    def convert_full_state_to_tp(self, full_state: Dict[str, torch.Tensor], args, tp_rank: Optional[int] = None, pp_rank: int = 0,
                                 ep_rank: int = 0, partitions: Optional[Sequence[str]] = None, config: Optional[Dict[str, Any]] = None):
        """Reference form (:513-720): the partial state of ONE (tp, pp, ep) coordinate from a pre-processed full state.
        Legacy form of this package: ``convert_full_state_to_tp(full, tp_size, args)`` → list of all TP shards (runs the
        pre-processing itself)."""
        if isinstance(args, int):                                   # legacy form: (full, tp_size, args)
            return self.shard_full_state(self.pre_process_full_state_before_tp_conversion(full_state, tp_rank), args, tp_rank)
        tp, pp, ep = args.tp_size, args.pp_size, getattr(args, "ep_size", 1)
        cache = getattr(self, "_shard_cache", None)
        if cache is None or cache[0] is not full_state:
            cache = (full_state, self.shard_full_state(full_state, tp, args))
            self._shard_cache = cache
        out: Dict[str, torch.Tensor] = {}
        for name, w in cache[1][tp_rank].items():
            if pp_rank != 0 and ("embed_tokens" in name or "wte" in name):
                continue
            if pp_rank != pp - 1 and ("lm_head" in name or name.endswith(("model.norm.weight", "norm_f.weight"))):
                continue
            if ep_rank != 0 and "expert_mlps" not in name:
                continue
            if pp > 1 and partitions is not None:
                owner = self._stage_of(name, partitions, pp)
                if owner is not None and owner != pp_rank:
                    continue
            if "expert_mlps" in name and ep > 1 and isinstance(w, torch.Tensor):
                if w.shape[0] % ep != 0:
                    raise ValueError(f"Expert dimension ({w.shape[0]}) is not divisible by expert parallelism degree ({ep}).")
                n = w.shape[0] // ep
                w = w.narrow(0, n * ep_rank, n).clone()
            out[self.rename_keys_for_megatron(name, getattr(args, "model_style", "hf"), True)] = w
        self.modify_qkv_for_megatron(out, args)
        return out

    # ---- sharded → full ----------------------------------------------------------------------------------------------
    def convert_tp_to_full_state(self, shards: List[Dict[str, torch.Tensor]], args=None) -> Dict[str, torch.Tensor]:
        full: Dict[str, torch.Tensor] = {}
        shards = [self.convert_partial_state_to_non_fused_qkv(dict(s), None, getattr(args, "kv_size_multiplier", 1),
                                                              getattr(args, "_config", None), len(shards)) for s in shards]
        for name in shards[0]:
            rule = self.get_partition_rule(name)
            parts = [s[name] for s in shards]
            if rule is None or not isinstance(parts[0], torch.Tensor) or parts[0].dim() == 0:
                full[name] = parts[0]
            else:
                full[name] = gather_full_weight(parts, rule[0], rule[1])
        return self.post_process_full_state_after_tp_conversion(full, args)

    def merge_tp_checkpoints(self, args) -> Dict[str, torch.Tensor]:
        """Read every (tp, pp, ep) partial checkpoint under ``args.input_dir`` and rebuild the full state."""
        ep = getattr(args, "ep_size", 1)
        load = self.load_partial_xser if getattr(args, "load_xser", False) else self.load_partial_no_xser
        shards = []
        for tp_rank in range(args.tp_size):
            merged: Dict[str, torch.Tensor] = {}
            experts: Dict[str, List[torch.Tensor]] = {}
            for pp_rank in range(args.pp_size):
                for ep_rank in range(ep):
                    part = load(args, tp_rank, pp_rank, ep_rank)
                    part = part.get(getattr(args, "model_key", "model"), part) if isinstance(part, dict) and \
                        isinstance(part.get(getattr(args, "model_key", "model")), dict) else part
                    for k, v in part.items():
                        k = self.rename_keys_for_megatron(k, getattr(args, "model_style", "hf"), False)
                        if "expert_mlps" in k and ep > 1:
                            experts.setdefault(k, []).append(v)
                        else:
                            merged[k] = v
            for k, vs in experts.items():
                merged[k] = torch.cat(vs, 0)
            shards.append(merged)
        return self.convert_tp_to_full_state(shards, args)

    # ---- IO ----------------------------------------------------------------------------------------------------------
    def download_and_save_hf_model(self, model_identifier: str, config_path: Optional[str] = None):
        """Needs network access and ``transformers`` — not available in an offline image."""
        from transformers import AutoConfig, AutoModelForCausalLM

        config = AutoConfig.from_pretrained(config_path) if config_path else None
        return AutoModelForCausalLM.from_pretrained(model_identifier, config=config).state_dict()

    def load_full_state(self, args) -> Dict[str, torch.Tensor]:
        if getattr(args, "hf_model_name", None) and not args.input_dir:
            return self.download_and_save_hf_model(args.hf_model_name, getattr(args, "config", None))
        if not args.input_dir:
            raise ValueError("Error: Please provide either HuggingFace model name or input path to consolidated statedict")
        path = args.input_dir
        if os.path.isdir(path):
            cands = [os.path.join(path, n) for n in ("pytorch_model.bin", "checkpoint.pt", "model.safetensors")]
            path = next((c for c in cands if os.path.isfile(c)), cands[0])
        if path.endswith(".safetensors"):
            from ..utils.safetensors_utils import load_state_dict_safetensors

            return load_state_dict_safetensors(path)
        return torch.load(path, map_location="cpu", weights_only=False)

    def _model_dir(self, root: str, args) -> str:
        tag = getattr(args, "tag", None)
        for cand in ([os.path.join(root, tag, "model")] if tag else []) + [os.path.join(root, "converted", "model"),
                                                                         os.path.join(root, "model"), root]:
            if os.path.isdir(cand):
                return cand
        return root

    def get_input_filename(self, args, tp_rank: int, pp_rank: int, ep_rank: int = 0, xser: bool = False) -> str:
        root = args.input_dir
        legacy = os.path.join(root, f"tp_rank_{tp_rank:02d}_pp_rank_{pp_rank:02d}")
        cands = [legacy if xser else os.path.join(legacy, "checkpoint.pt")]
        d = self._model_dir(root, args)
        cands += [os.path.join(d, f"dp_rank_00_tp_rank_{tp_rank:02d}_pp_rank_{pp_rank:02d}.pt"),
                  os.path.join(d, f"dp_rank_00_ep_rank_{ep_rank:02d}_tp_rank_{tp_rank:02d}_pp_rank_{pp_rank:02d}.pt")]
        for c in cands:
            if os.path.exists(c):
                return c
        raise RuntimeError(f"Error: none of {cands} exist")

    def get_output_filename(self, args, tp_rank: int, pp_rank: int, ep_rank: int = 0, xser: bool = False) -> str:
        if getattr(args, "legacy_format", False):
            return os.path.join(args.output_dir, f"tp_rank_{tp_rank:02d}_pp_rank_{pp_rank:02d}", "checkpoint.pt")
        d = os.path.join(args.output_dir, getattr(args, "tag", None) or "converted", "model") if getattr(args, "_tagged_output", True) \
            else os.path.join(args.output_dir, "model")
        name = f"dp_rank_00_ep_rank_{ep_rank:02d}_tp_rank_{tp_rank:02d}_pp_rank_{pp_rank:02d}.pt" if getattr(args, "ep_size", 1) > 1 \
            else f"dp_rank_00_tp_rank_{tp_rank:02d}_pp_rank_{pp_rank:02d}.pt"
        return os.path.join(d, name)

    @staticmethod
    def prune_state(state: Dict[str, Any], ep_rank: int) -> None:
        """EP ranks > 0 store only expert tensors; drop the ``None`` placeholders of everything else."""
        if ep_rank > 0:
            for k in [k for k, v in state.items() if v is None]:
                state.pop(k)

    def load_partial_no_xser(self, args, tp_rank: int, pp_rank: int, ep_rank: int = 0):
        return torch.load(self.get_input_filename(args, tp_rank, pp_rank, ep_rank, False), map_location="cpu", weights_only=False)

    def load_partial_xser(self, args, tp_rank: int, pp_rank: int, ep_rank: int = 0):
        from ..trainer.checkpoint import _xser_load
        from ..trainer.checkpoint_storage import FilesysCheckpointStorage

        path = self.get_input_filename(args, tp_rank, pp_rank, ep_rank, True)
        d, f = os.path.split(path)
        state = _xser_load(FilesysCheckpointStorage(d), f, None, 1, 0)
        self.prune_state(state, ep_rank)
        return state

    def save_partial_no_xser(self, args, partial_state, tp_rank: int, pp_rank: int, ep_rank: int = 0) -> None:
        f = self.get_output_filename(args, tp_rank, pp_rank, ep_rank, False)
        os.makedirs(os.path.dirname(f), exist_ok=True)
        torch.save({"model": partial_state} if getattr(args, "legacy_format", False) else partial_state, f)

    def save_partial_xser(self, args, partial_state, tp_rank: int, pp_rank: int, ep_rank: int = 0) -> None:
        from ..trainer.checkpoint import CheckpointIOState, _xser_tasks

        f = self.get_output_filename(args, tp_rank, pp_rank, ep_rank, True)
        os.makedirs(os.path.dirname(f), exist_ok=True)
        io = CheckpointIOState(False)
        _xser_tasks(partial_state, f, 1, 0, io)
        for obj, name in io.items:
            os.makedirs(os.path.dirname(name), exist_ok=True)
            if type(obj).__name__ == "_RawBytes":          # the reference-named TensorReference pickle, already serialised
                with open(name, "wb") as fh:
                    fh.write(obj.data)
            else:
                torch.save(obj, name)

    def save_full(self, args, full_state) -> None:
        path = args.output_dir
        if not path.endswith((".pt", ".bin")):
            os.makedirs(path, exist_ok=True)
            path = os.path.join(path, getattr(args, "full_name", "pytorch_model.bin"))
        torch.save(full_state, path)

    def _mark_complete(self, args) -> None:
        if getattr(args, "legacy_format", False):
            return
        root = os.path.dirname(os.path.dirname(self.get_output_filename(args, 0, 0, 0, False)))
        for marker in ("checkpoint", "done"):
            with open(os.path.join(root, marker), "w") as f:
                f.write("1")

    # ---- drivers -----------------------------------------------------------------------------------------------------
    def _each_coord(self, args):
        for tp_rank in range(args.tp_size):
            for pp_rank in range(args.pp_size):
                for ep_rank in range(getattr(args, "ep_size", 1)):
                    yield tp_rank, pp_rank, ep_rank

    def convert_from_xser(self, args) -> None:
        for c in self._each_coord(args):
            self.save_partial_no_xser(args, self.load_partial_xser(args, *c), *c)
        self._mark_complete(args)

    def convert_to_xser(self, args) -> None:
        for c in self._each_coord(args):
            self.save_partial_xser(args, self.load_partial_no_xser(args, *c), *c)
        self._mark_complete(args)

    def _read_config(self, args) -> Dict[str, Any]:
        cfg: Dict[str, Any] = {}
        if getattr(args, "config", None):
            with open(args.config) as f:
                cfg = json.load(f)
            for attr, key in (("_q_heads", "num_attention_heads"), ("_kv_heads", "num_key_value_heads")):
                try:
                    setattr(args, attr, self._get_config_value(cfg, key))
                except KeyError:
                    pass
            args._config = cfg
        return cfg

    def convert_from_full_state(self, args) -> None:
        full_state = self.load_full_state(args)
        full_state = full_state.get(getattr(args, "model_key", "model"), full_state) \
            if isinstance(full_state.get(getattr(args, "model_key", "model"), None), dict) else full_state
        config = self._read_config(args)
        layer_names = sorted({m.group(1) for k in full_state if (m := re.match(self.layer_name_pattern, k))},
                             key=lambda s: int(re.findall(r"\d+", s)[-1]))
        n_layers = len(layer_names) or getattr(args, "n_layers", 0)
        stages = args.pp_size * getattr(args, "virtual_pp_size", 1)
        partitions = [layer_names[i] if layer_names else f"model.{self.layer_name}.{i}" for i in create_partitions(n_layers, stages)] \
            if stages > 1 else []
        if getattr(args, "coalesce_qkv", False):
            full_state = self.coalesce_qkv(full_state, config, args.tp_size)
        full_state = self.pre_process_full_state_before_tp_conversion(full_state, args)
        save = self.save_partial_xser if getattr(args, "save_xser", False) else self.save_partial_no_xser
        for tp_rank, pp_rank, ep_rank in self._each_coord(args):
            save(args, self.convert_full_state_to_tp(full_state, args, tp_rank, pp_rank, ep_rank, partitions, config), tp_rank, pp_rank, ep_rank)
        self._shard_cache = None
        self._mark_complete(args)

    def convert_to_full_state(self, args) -> None:
        self._read_config(args)
        self.save_full(args, self.merge_tp_checkpoints(args))

    # ---- CLI ---------------------------------------------------------------------------------------------------------
    def get_arg_parser(self) -> argparse.ArgumentParser:
        def flag(v) -> bool:
            return str(v).lower() in ("1", "true", "yes", "y")

        ap = argparse.ArgumentParser(description="Convert model checkpoints between full and (TP, PP, EP)-sharded layouts")
        ap.add_argument("--input_dir", type=str, default=None, help="full checkpoint file / directory, or the sharded checkpoint directory")
        ap.add_argument("--hf_model_name", type=str, default=None, help="HuggingFace model identifier (needs network)")
        ap.add_argument("--output_dir", type=str, required=True)
        ap.add_argument("--hw_backend", type=str, default="b200", help="accepted for compatibility (trn1 → tile KV layout, trn2 → adjacent)")
        ap.add_argument("--config", type=str, default=None, help="HF config.json (head counts for GQA)")
        ap.add_argument("--model_key", type=str, default="model")
        ap.add_argument("--tp_size", type=int, default=1)
        ap.add_argument("--pp_size", type=int, default=1)
        ap.add_argument("--ep_size", type=int, default=1)
        ap.add_argument("--virtual_pp_size", type=int, default=1)
        ap.add_argument("--n_layers", type=int, default=0)
        ap.add_argument("--coalesce_qkv", type=flag, nargs="?", const=True, default=False)
        ap.add_argument("--kv_size_multiplier", type=int, default=1)
        ap.add_argument("--kv_replication_layout", choices=["tile", "adjacent"], default=None)
        ap.add_argument("--qkv_linear", type=flag, nargs="?", const=True, default=False)
        ap.add_argument("--fuse_qkv", type=flag, nargs="?", const=True, default=False)
        ap.add_argument("--fuse_gate_up", type=flag, nargs="?", const=True, default=False)
        ap.add_argument("--i_tp_round_factor", type=int, default=0, help="accepted for compatibility (no padding is needed by the grouped GEMM)")
        ap.add_argument("--pad_attn_heads", type=flag, nargs="?", const=True, default=False)
        ap.add_argument("--load_xser", type=flag, nargs="?", const=True, default=False)
        ap.add_argument("--save_xser", type=flag, nargs="?", const=True, default=False)
        ap.add_argument("--convert_from_xser", action="store_true")
        ap.add_argument("--convert_to_xser", action="store_true")
        ap.add_argument("--convert_from_full_state", action="store_true")
        ap.add_argument("--convert_to_full_state", action="store_true")
        ap.add_argument("--model_style", type=str, choices=["hf", "megatron"], default="hf")
        ap.add_argument("--nxdt_yaml_config", type=str, default=None)
        ap.add_argument("--legacy_format", action="store_true", help="tp_rank_XX_pp_rank_XX/checkpoint.pt layout")
        ap.add_argument("--tag", default=None, help="checkpoint tag directory under --output_dir / --input_dir")
        return ap

    def run(self, args) -> None:
        modes = ["convert_from_full_state", "convert_to_full_state", "convert_from_xser", "convert_to_xser"]
        assert sum(int(bool(getattr(args, m, False))) for m in modes) == 1, "Exactly one '--convert_*' flag must be specified"
        if getattr(args, "kv_replication_layout", None) is None:
            args.kv_replication_layout = "adjacent" if str(getattr(args, "hw_backend", "")).lower() == "trn2" else "tile"
        if getattr(args, "nxdt_yaml_config", None):
            from .yaml_converter import convert_yaml_to_json

            args.config = convert_yaml_to_json(args.nxdt_yaml_config)
        for m in modes:
            if getattr(args, m, False):
                return getattr(self, m)(args)


def main(argv=None) -> int:
    c = CheckpointConverterBase()
    c.run(c.get_arg_parser().parse_args(argv))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
