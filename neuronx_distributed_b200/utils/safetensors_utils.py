"""safetensors helpers (reference ``utils/safetensors_utils.py``): de-duplicate tensors that share storage
(tied weights) before ``save_file``, remembering the aliases so ``load`` can re-tie them."""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def remove_shared_tensors(state_dict: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], Dict[str, str]]:
    seen: Dict[Tuple[int, int, tuple], str] = {}
    out: Dict[str, torch.Tensor] = {}
    aliases: Dict[str, str] = {}
    for k, v in state_dict.items():
        if not isinstance(v, torch.Tensor):
            continue
        key = (v.untyped_storage().data_ptr(), v.storage_offset(), tuple(v.shape))
        if v.numel() and key in seen:
            aliases[k] = seen[key]
            continue
        seen[key] = k
        out[k] = v.contiguous()
    return out, aliases


def save_state_dict_safetensors(state_dict: Dict[str, torch.Tensor], path: str) -> None:
    from safetensors.torch import save_file

    sd, aliases = remove_shared_tensors(state_dict)
    save_file({k: v.cpu() for k, v in sd.items()}, path, metadata={f"alias:{k}": v for k, v in aliases.items()})


def load_state_dict_safetensors(path: str) -> Dict[str, torch.Tensor]:
    from safetensors import safe_open

    out: Dict[str, torch.Tensor] = {}
    with safe_open(path, framework="pt") as f:
        for k in f.keys():
            out[k] = f.get_tensor(k)
        meta = f.metadata() or {}
    for k, v in meta.items():
        if k.startswith("alias:"):
            out[k[len("alias:"):]] = out[v]
    return out
