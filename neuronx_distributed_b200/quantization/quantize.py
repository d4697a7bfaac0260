"""``convert`` — swap float parallel layers for their quantised counterparts (reference ``quantization/quantize.py:18-160``)."""
from __future__ import annotations

import copy
from fnmatch import fnmatch
from typing import Any, Callable, Dict, List, Optional, Union

import torch
from torch import nn

from ..utils.logger import get_logger
from .quantization_config import BASE_QCONFIG_DICT_TYPE, get_default_custom_qconfig_dict
from .quantization_mappings import get_default_quant_module_mappings

logger = get_logger()


def convert(module: nn.Module, q_config: Optional[BASE_QCONFIG_DICT_TYPE] = None, inplace: bool = False,
            mapping: Optional[Dict[Callable, Any]] = None, include: Optional[Union[str, List[str]]] = None,
            modules_to_not_convert: Optional[List[str]] = None) -> nn.Module:
    """Replace every layer whose type is in ``mapping`` by ``mapping[type].from_float(layer, q_config)``.

    ``include``: allow-list of ``fnmatch`` patterns over dotted module names (``"*mlp.down_proj"``); a matching module is
    swapped if its type is mapped, otherwise everything mapped underneath it is swapped.  ``modules_to_not_convert``:
    deny-list — a layer is skipped when its own name equals an entry or any entry is a substring of its dotted path.
    The two are mutually exclusive, as in the reference."""
    assert include is None or modules_to_not_convert is None, (
        "Either include and modules_to_not_convert both should be None, or only one of them should be not-None. "
        f"Provided values include: {include} , modules_to_not_convert: {modules_to_not_convert}")
    if not inplace:
        module = copy.deepcopy(module)
    q_config = q_config if q_config is not None else get_default_custom_qconfig_dict()
    mapping = mapping if mapping is not None else get_default_quant_module_mappings()
    if include is None:
        _convert_initialized_float_to_initialized_quantized(module, q_config, mapping, modules_to_not_convert=modules_to_not_convert)
        return module
    patterns = [include] if isinstance(include, str) else list(include)
    # plain names behave like suffix patterns ("down_proj" selects every *.down_proj)
    patterns = [p if any(c in p for c in "*?[") else f"*{p}" for p in patterns]
    for name, sub in list(module.named_modules()):
        if not name or not any(fnmatch(name, p) for p in patterns):
            continue
        if type(sub) in mapping:
            _swap_module(module, sub, name, q_config, mapping)
        else:
            _convert_initialized_float_to_initialized_quantized(sub, q_config, mapping)
    return module


def _swap_module(root_module: nn.Module, module_to_swap: nn.Module, module_name_to_swap: str,
                 q_config: BASE_QCONFIG_DICT_TYPE, mapping: Dict[Callable, Any]) -> None:
    parent_name, _, leaf = module_name_to_swap.rpartition(".")
    parent = root_module.get_submodule(parent_name) if parent_name else root_module
    setattr(parent, leaf, mapping[type(module_to_swap)].from_float(module_to_swap, q_config))


def _convert_initialized_float_to_initialized_quantized(module: nn.Module, q_config: BASE_QCONFIG_DICT_TYPE,
                                                        mapping: Dict[Callable, Any], prefixes: Optional[List[str]] = None,
                                                        modules_to_not_convert: Optional[List[str]] = None) -> nn.Module:
    deny = list(modules_to_not_convert or [])
    prefixes = prefixes if prefixes is not None else []
    for name, child in list(module.named_children()):
        path = ".".join(prefixes + [name])
        if type(child) in mapping:
            if name in deny or any(key in path for key in deny):
                continue
            logger.debug("Quantizing %s to %s", path, q_config.get("quantized_dtype"))
            module._modules[name] = mapping[type(child)].from_float(child, q_config)
        else:
            _convert_initialized_float_to_initialized_quantized(child, q_config, mapping, prefixes + [name], deny)
    return module


def direct_cast_quantize(tensor: torch.Tensor, downcast_dtype: torch.dtype) -> torch.Tensor:
    """Scale-free down-cast (fp8 KV cache with ``direct_cast``)."""
    return tensor.to(downcast_dtype)
