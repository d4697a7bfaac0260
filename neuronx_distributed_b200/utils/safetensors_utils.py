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


def check_for_duplicate_tensors(checkpoint: Dict[str, torch.Tensor], remove_duplicate_tensors: bool = False) -> Dict[str, torch.Tensor]:
    """safetensors refuses tensors that share storage.  Find groups of entries backed by the same storage; with
    ``remove_duplicate_tensors`` keep one name per group — the (alphabetically first) entry that covers the whole storage —
    and delete the others in place; without it only warn (shard-on-load rewrites the tensors anyway).  A group in which no
    entry covers its storage cannot be de-duplicated safely and raises (reference :111-178)."""
    import logging
    from collections import defaultdict

    groups = defaultdict(list)
    for name, t in checkpoint.items():
        if isinstance(t, torch.Tensor) and t.device.type != "meta" and t.numel():
            groups[(t.device, t.untyped_storage().data_ptr())].append(name)
    for names in groups.values():
        if len(names) < 2:
            continue
        complete = sorted(n for n in names if checkpoint[n].storage_offset() == 0
                          and checkpoint[n].numel() * checkpoint[n].element_size() == checkpoint[n].untyped_storage().nbytes())
        if not complete:
            if remove_duplicate_tensors:
                raise RuntimeError(f"shared tensors {names}: none covers the entire storage, refusing to pick one to keep")
            logging.warning("Found shared tensors %s, this can cause checkpoint saving to fail. This warning can be safely "
                            "ignored when using shard-on-load.", names)
            continue
        if remove_duplicate_tensors:
            for n in names:
                if n != complete[0]:
                    del checkpoint[n]
    return checkpoint
