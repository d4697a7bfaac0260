"""Grouped-query attention sharding for inference when the head counts do not divide the TP degree
(role of the reference's ``examples/inference/modules/gqa.py``: ``GQA`` strategies, ``determine_sharding_strategy``,
``get_shardable_head_counts``, ``GroupQueryAttention_QKV`` / ``GroupQueryAttention_O`` and their preshard hooks).

Two strategies (the enum values are a checkpoint / config contract):

* ``REPLICATE_TO_TP_DEGREE`` — K/V heads are repeated until there is exactly one per TP rank; each K/V group's query heads
  are padded (zero heads appended to the GROUP) so every rank holds whole groups:   ``| K1 | K1 | K2 | K2 |`` /
  ``| Q1 Q2 | Q3 pad | Q4 Q5 | Q6 pad |``.  Needs ``tp % kv_heads == 0``.
* ``CONVERT_TO_MHA`` — every K/V head is repeated once per query head of its group, then both are zero-padded at the tail to a
  multiple of TP.  Works for any configuration, costs more KV-cache memory.

Design: a :class:`HeadLayout` computes, once, for every TARGET head slot the SOURCE head it holds (or ``-1`` = zero padding);
every weight / bias / per-channel scale transform is then the same gather along the head axis (:func:`remap_heads`) — the
query projection rows, the K/V projection rows and the output projection columns all use the slot maps of one object, so they
cannot disagree.  Zero query heads produce a uniform softmax over V, but their output-projection columns are zero as well, so
padded heads contribute nothing.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from ..parallel_layers import parallel_state as ps
from ..parallel_layers.layers import ColumnParallelLinear, RowParallelLinear
from ..parallel_layers.pad import get_number_of_extra_heads


class GQA(enum.Enum):
    CONVERT_TO_MHA = "convert-to-mha"
    REPLICATE_TO_TP_DEGREE = "replicate-to-tp-degree"


def determine_sharding_strategy(tp_degree: int, source_key_value_heads: int,
                                desired_sharding_strategy: Optional[GQA] = None) -> GQA:
    """The requested strategy (default: replicate to the TP degree), downgraded to MHA conversion when the K/V heads cannot be
    replicated evenly over the ranks."""
    want = desired_sharding_strategy or GQA.REPLICATE_TO_TP_DEGREE
    if want == GQA.REPLICATE_TO_TP_DEGREE and tp_degree % source_key_value_heads != 0:
        return GQA.CONVERT_TO_MHA
    return want


def get_shardable_head_counts(tp_degree: int, num_attention_heads: int, num_key_value_heads: int,
                              sharding_strategy: GQA) -> Tuple[int, int]:
    """``(query heads, key/value heads)`` after padding / replication, both divisible by ``tp_degree``."""
    q = num_attention_heads + get_number_of_extra_heads(num_attention_heads, tp_degree)
    kv = num_key_value_heads
    if num_attention_heads == num_key_value_heads:                       # MHA: K/V follow Q
        kv = q
    elif num_key_value_heads < tp_degree or num_key_value_heads % tp_degree != 0:
        if sharding_strategy == GQA.REPLICATE_TO_TP_DEGREE:
            assert tp_degree % num_key_value_heads == 0, "REPLICATE_TO_TP_DEGREE needs tp_degree % num_key_value_heads == 0"
            kv = tp_degree
        else:
            kv = q
    return q, kv


@dataclass(frozen=True)
class HeadLayout:
    """Source-head index held by every target head slot (``-1`` = zero padding)."""

    tp_degree: int
    src_q: int
    src_kv: int
    strategy: GQA
    q: int
    kv: int

    @classmethod
    def build(cls, tp_degree: int, num_attention_heads: int, num_key_value_heads: int,
              desired: Optional[GQA] = None) -> "HeadLayout":
        strat = determine_sharding_strategy(tp_degree, num_key_value_heads, desired)
        q, kv = get_shardable_head_counts(tp_degree, num_attention_heads, num_key_value_heads, strat)
        return cls(tp_degree, num_attention_heads, num_key_value_heads, strat, q, kv)

    @property
    def group(self) -> int:
        return self.src_q // self.src_kv

    def q_slots(self) -> List[int]:
        if self.strategy == GQA.REPLICATE_TO_TP_DEGREE and self.src_q != self.src_kv and self.q != self.src_q:
            per = self.q // self.src_kv                      # slots per K/V group after padding
            assert per >= self.group and self.q % self.src_kv == 0
            return [g * self.group + j if j < self.group else -1 for g in range(self.src_kv) for j in range(per)]
        return list(range(self.src_q)) + [-1] * (self.q - self.src_q)

    def kv_slots(self) -> List[int]:
        if self.kv == self.src_kv:
            return list(range(self.src_kv))
        if self.src_q == self.src_kv:                        # MHA: tail padding only
            return list(range(self.src_kv)) + [-1] * (self.kv - self.src_kv)
        rep = (self.tp_degree // self.src_kv) if self.strategy == GQA.REPLICATE_TO_TP_DEGREE else self.group
        rep_slots = [h for h in range(self.src_kv) for _ in range(rep)]
        return rep_slots + [-1] * (self.kv - len(rep_slots))


def remap_heads(tensor: Optional[torch.Tensor], slots: List[int], num_source_heads: int, dim: int = 0) -> Optional[torch.Tensor]:
    """Gather whole heads along ``dim`` (size ``num_source_heads * head_size``) into the slot order; ``-1`` slots are zeros."""
    if tensor is None:
        return None
    assert tensor.shape[dim] % num_source_heads == 0, (tensor.shape, dim, num_source_heads)
    hs = tensor.shape[dim] // num_source_heads
    idx = torch.tensor([s if s >= 0 else 0 for s in slots], dtype=torch.long)
    keep = torch.tensor([s >= 0 for s in slots])
    view = tensor.movedim(dim, 0).reshape(num_source_heads, hs, *tensor.shape[:dim], *tensor.shape[dim + 1:])
    raw = view.view(torch.uint8) if view.dtype in (torch.float8_e4m3fn, torch.float8_e5m2) else view   # no fp8 gather on CPU
    out = raw.index_select(0, idx)
    out[~keep] = 0
    out = out.view(view.dtype) if raw is not view else out
    out = out.reshape(len(slots) * hs, *tensor.shape[:dim], *tensor.shape[dim + 1:]).movedim(0, dim)
    return out.contiguous()


def _remap_scale(scale: Optional[torch.Tensor], weight_before: torch.Tensor, slots: List[int], heads: int, dim: int):
    """Per-channel scales along the remapped axis follow their channels; per-tensor scales are untouched."""
    if scale is None or scale.dim() <= dim or scale.shape[dim] != weight_before.shape[dim]:
        return scale
    return remap_heads(scale, slots, heads, dim)


def _rename(sd: Dict[str, torch.Tensor], old_prefix: str, new_prefix: str) -> None:
    if old_prefix == new_prefix:
        return
    for k in [k for k in sd if k.startswith(old_prefix + ".")]:
        sd[new_prefix + k[len(old_prefix):]] = sd.pop(k)


class BaseGroupQueryAttention(nn.Module):
    def __init__(self, hidden_size: int, head_dim: int, num_attention_heads: int, num_key_value_heads: int, tp_degree: int = 1,
                 dtype: torch.dtype = torch.float32, bias: bool = False, desired_sharding_strategy: Optional[GQA] = None,
                 tensor_model_parallel_group=None):
        super().__init__()
        group = tensor_model_parallel_group
        if group is None and ps.model_parallel_is_initialized():
            group = ps.get_tensor_model_parallel_group()
        self.tensor_model_parallel_group = group
        if group is not None:
            import torch.distributed as dist

            n = dist.get_world_size(group)
            assert tp_degree in (1, n), "tp_degree and the tensor-parallel group size differ"
            tp_degree = n
        self.hidden_size, self.head_dim, self.tp_degree, self.dtype, self.bias = hidden_size, head_dim, tp_degree, dtype, bias
        self.layout = HeadLayout.build(tp_degree, num_attention_heads, num_key_value_heads, desired_sharding_strategy)
        self._src_num_attention_heads, self._src_num_key_value_heads = num_attention_heads, num_key_value_heads
        self.sharding_strategy = self.layout.strategy
        self.num_attention_heads, self.num_key_value_heads = self.layout.q, self.layout.kv

    def get_sharding_strategy(self) -> GQA:
        return self.sharding_strategy

    def get_num_attention_heads(self) -> int:
        return self.num_attention_heads

    def get_num_key_value_heads(self) -> int:
        return self.num_key_value_heads

    @staticmethod
    def _paths(key: str) -> Tuple[str, str]:
        """``…self_attn.qkv_proj.weight`` → (module path ``…self_attn.qkv_proj``, HF parent path ``…self_attn``)."""
        mod = key.rsplit(".", 1)[0]
        return mod, mod.rsplit(".", 1)[0] if "." in mod else ""


class GroupQueryAttention_QKV(BaseGroupQueryAttention):
    """Q / K / V projections with the head layout above; ``fused_qkv`` keeps one ``Wqkv`` weight ``[q + 2·kv, hidden]``.
    The preshard hook accepts HF checkpoints (``…q_proj / k_proj / v_proj`` or ``…Wqkv`` under the parent module)."""

    def __init__(self, hidden_size: int, head_dim: int, num_attention_heads: int, num_key_value_heads: int, tp_degree: int = 1,
                 dtype: torch.dtype = torch.float32, bias: bool = False, desired_sharding_strategy: Optional[GQA] = None,
                 gather_output: bool = True, fused_qkv: bool = False, clip_qkv: Optional[float] = None,
                 sequence_parallel_enabled: bool = False, sequence_dimension: Optional[int] = None,
                 tensor_model_parallel_group=None, device=None):
        super().__init__(hidden_size, head_dim, num_attention_heads, num_key_value_heads, tp_degree, dtype, bias,
                         desired_sharding_strategy, tensor_model_parallel_group)
        self.gather_output, self.fused_qkv, self.clip_qkv = gather_output, fused_qkv, clip_qkv
        q_out, kv_out = self.num_attention_heads * head_dim, self.num_key_value_heads * head_dim
        kw = dict(bias=bias, gather_output=gather_output, dtype=dtype, sequence_parallel_enabled=sequence_parallel_enabled,
                  sequence_dimension=sequence_dimension, tensor_model_parallel_group=self.tensor_model_parallel_group, device=device)
        if self.tensor_model_parallel_group is None:
            mk = lambda o: nn.Linear(hidden_size, o, bias=bias, dtype=dtype, device=device)           # noqa: E731
        else:
            mk = lambda o: ColumnParallelLinear(hidden_size, o, **kw)                                 # noqa: E731
        if fused_qkv:
            # rank-major fused layout: rank r's rows are [q_r ; k_r ; v_r], so a plain dim-0 shard of the fused tensor is correct
            self.Wqkv = mk(q_out + 2 * kv_out)
        else:
            self.q_proj, self.k_proj, self.v_proj = mk(q_out), mk(kv_out), mk(kv_out)

    def forward(self, hidden_states: torch.Tensor):
        if self.fused_qkv:
            qkv = self.Wqkv(hidden_states)
            if self.clip_qkv is not None:
                qkv = qkv.clamp(-self.clip_qkv, self.clip_qkv)
            div = 1 if self.gather_output else self.tp_degree
            qn, kn = self.num_attention_heads * self.head_dim // div, self.num_key_value_heads * self.head_dim // div
            if self.gather_output and self.tp_degree > 1:
                # gathered fused output is rank-major: regroup into [Q ; K ; V]
                parts = qkv.view(*qkv.shape[:-1], self.tp_degree, (qn + 2 * kn) // self.tp_degree)
                a, b = qn // self.tp_degree, kn // self.tp_degree
                q, k, v = parts[..., :a], parts[..., a:a + b], parts[..., a + b:]
                return tuple(t.reshape(*qkv.shape[:-1], -1) for t in (q, k, v))
            return torch.split(qkv, [qn, kn, kn], dim=-1)
        q, k, v = self.q_proj(hidden_states), self.k_proj(hidden_states), self.v_proj(hidden_states)
        if self.clip_qkv is not None:
            q, k, v = (t.clamp(-self.clip_qkv, self.clip_qkv) for t in (q, k, v))
        return q, k, v

    def preshard_hook(self, model_state_dict: Dict[str, torch.Tensor], prefix: str) -> bool:
        mod, parent = self._paths(prefix)
        L, sd = self.layout, model_state_dict
        src: Dict[str, Dict[str, Optional[torch.Tensor]]] = {}
        if any(k.startswith(f"{parent}.Wqkv.") or k.startswith(f"{mod}.Wqkv.") for k in sd) and f"{parent}.q_proj.weight" not in sd \
                and f"{mod}.q_proj.weight" not in sd:
            base = f"{mod}.Wqkv" if f"{mod}.Wqkv.weight" in sd else f"{parent}.Wqkv"
            sizes = [L.src_q * self.head_dim, L.src_kv * self.head_dim, L.src_kv * self.head_dim]
            for name in ("weight", "bias", "scale"):
                t = sd.pop(f"{base}.{name}", None)
                pieces = (None, None, None)
                if t is not None and t.dim() >= 1 and t.shape[0] == sum(sizes):
                    pieces = torch.split(t, sizes, dim=0)
                elif t is not None:                          # per-tensor scale: shared by the three sections
                    pieces = (t, t, t)
                for n, p in zip(("q", "k", "v"), pieces):
                    src.setdefault(n, {})[name] = p
        else:
            for n in ("q", "k", "v"):
                base = f"{mod}.{n}_proj" if f"{mod}.{n}_proj.weight" in sd else f"{parent}.{n}_proj"
                src[n] = {name: sd.pop(f"{base}.{name}", None) for name in ("weight", "bias", "scale")}
        if src["q"]["weight"] is None:
            raise KeyError(f"no q/k/v projection weights under {parent!r} or {mod!r}")
        out: Dict[str, Dict[str, Optional[torch.Tensor]]] = {}
        for n, slots, heads in (("q", L.q_slots(), L.src_q), ("k", L.kv_slots(), L.src_kv), ("v", L.kv_slots(), L.src_kv)):
            w = src[n]["weight"]
            out[n] = {"weight": remap_heads(w, slots, heads, 0), "bias": remap_heads(src[n]["bias"], slots, heads, 0),
                      "scale": _remap_scale(src[n]["scale"], w, slots, heads, 0)}
        if self.fused_qkv:
            tp = self.tp_degree

            def rank_major(name: str) -> Optional[torch.Tensor]:
                ts = [out[n][name] for n in ("q", "k", "v")]
                if any(t is None for t in ts):
                    return None
                if ts[0].dim() == 0 or ts[0].shape[0] in (1,):
                    return ts[0]
                chunks = [t.chunk(tp, 0) for t in ts]
                return torch.cat([torch.cat([c[r] for c in chunks], 0) for r in range(tp)], 0)

            for name in ("weight", "bias", "scale"):
                t = rank_major(name)
                if t is not None:
                    sd[f"{mod}.Wqkv.{name}"] = t
        else:
            for n in ("q", "k", "v"):
                for name, t in out[n].items():
                    if t is not None:
                        sd[f"{mod}.{n}_proj.{name}"] = t
        return True


class GroupQueryAttention_O(BaseGroupQueryAttention):
    """Output projection whose INPUT columns follow the padded query-head order of :class:`GroupQueryAttention_QKV`."""

    def __init__(self, hidden_size: int, head_dim: int, num_attention_heads: int, num_key_value_heads: int, tp_degree: int = 1,
                 dtype: torch.dtype = torch.float32, bias: bool = False, desired_sharding_strategy: Optional[GQA] = None,
                 input_is_parallel: bool = False, layer_name: str = "o_proj", sequence_parallel_enabled: bool = False,
                 sequence_dimension: Optional[int] = None, tensor_model_parallel_group=None, device=None):
        super().__init__(hidden_size, head_dim, num_attention_heads, num_key_value_heads, tp_degree, dtype, bias,
                         desired_sharding_strategy, tensor_model_parallel_group)
        self.input_is_parallel, self.layer_name = input_is_parallel, layer_name
        in_features = self.num_attention_heads * head_dim
        if self.tensor_model_parallel_group is None:
            self.o_proj = nn.Linear(in_features, hidden_size, bias=bias, dtype=dtype, device=device)
        else:
            self.o_proj = RowParallelLinear(in_features, hidden_size, bias=bias, input_is_parallel=input_is_parallel, dtype=dtype,
                                            sequence_parallel_enabled=sequence_parallel_enabled, sequence_dimension=sequence_dimension,
                                            tensor_model_parallel_group=self.tensor_model_parallel_group, device=device)

    def forward(self, attention_output: torch.Tensor) -> torch.Tensor:
        return self.o_proj(attention_output)

    def preshard_hook(self, model_state_dict: Dict[str, torch.Tensor], prefix: str) -> bool:
        mod, parent = self._paths(prefix)
        sd, L = model_state_dict, self.layout
        base = f"{mod}.o_proj" if f"{mod}.o_proj.weight" in sd else f"{parent}.{self.layer_name}"
        if f"{base}.weight" not in sd:
            raise KeyError(f"no output projection weight under {base!r}")
        _rename(sd, base, f"{mod}.o_proj")
        w = sd[f"{mod}.o_proj.weight"]
        sd[f"{mod}.o_proj.weight"] = remap_heads(w, L.q_slots(), L.src_q, 1)
        sc = sd.get(f"{mod}.o_proj.scale")
        if sc is not None:
            sd[f"{mod}.o_proj.scale"] = _remap_scale(sc, w, L.q_slots(), L.src_q, 1)
        return True
