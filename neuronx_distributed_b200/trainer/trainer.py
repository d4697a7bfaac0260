"""Trainer façade: ``neuronx_distributed_config`` → ``initialize_parallel_model`` →
``initialize_parallel_optimizer`` (reference ``trainer/trainer.py:32-315``)."""
from __future__ import annotations

import os
from pprint import pformat
from typing import Any, Callable, Dict, Optional

import torch
import torch.distributed as dist

from ..optimizer.zero_redundancy_optimizer import NeuronEPZero1Optimizer, NeuronZero1Optimizer
from ..parallel_layers import parallel_state as ps
from ..utils import get_device
from ..utils.logger import get_logger
from .post_partition_hooks import hooks
from .model import NxDModel
from .optimizer import NxDOptimizer

logger = get_logger()

_RS_AG_BUCKET_CAP_MB = 130


def _fill(d: Optional[dict], defaults: dict, name: str) -> dict:
    if d is None:
        return dict(defaults)
    assert isinstance(d, dict), f"{name} must be a dict."
    out = dict(d)
    for k, v in defaults.items():
        if k not in out:
            logger.warning("%s[%s] is not set, automatically set it to %s.", name, k, v)
            out[k] = v
    return out


def neuronx_distributed_config(
    tensor_parallel_size: int = 1,
    pipeline_parallel_size: int = 1,
    expert_parallel_size: int = 1,
    context_parallel_size: int = 1,
    pipeline_config: Optional[dict] = None,
    optimizer_config: Optional[dict] = None,
    activation_checkpoint_config: Any = None,
    pad_model: bool = False,
    sequence_parallel: bool = False,
    model_init_config: Optional[dict] = None,
    lora_config: Any = None,
    mixed_precision_config: Optional[dict] = None,
    sequential_move_factor: int = 11,
    lnc_size: int = 1,
) -> Dict[str, Any]:
    """Build the plain-dict config all other entry points consume, and bring up the parallel
    groups if ``torch.distributed`` is initialised."""
    optimizer_config = _fill(optimizer_config, {"zero_one_enabled": False, "grad_clipping": True}, "optimizer_config")
    if optimizer_config["grad_clipping"]:
        optimizer_config.setdefault("max_grad_norm", 1.0)
    z = optimizer_config["zero_one_enabled"]
    mixed_precision_config = _fill(
        mixed_precision_config,
        {"use_master_weights": z, "use_fp32_grad_acc": z, "use_master_weights_in_ckpt": False},
        "mixed_precision_config",
    )
    model_init_config = _fill(
        model_init_config,
        {"sequential_move_factor": sequential_move_factor, "meta_device_init": False, "param_init_fn": None},
        "model_init_config",
    )
    if model_init_config["meta_device_init"] and model_init_config.get("param_init_fn") is None:
        raise ValueError("param_init_fn must be provided when meta_device_init is True")
    config = {
        "tensor_parallel_size": tensor_parallel_size,
        "pipeline_parallel_size": pipeline_parallel_size,
        "expert_parallel_size": expert_parallel_size,
        "context_parallel_size": context_parallel_size,
        "pipeline_config": pipeline_config,
        "optimizer_config": optimizer_config,
        "activation_checkpoint_config": activation_checkpoint_config,
        "pad_model": pad_model,
        "sequence_parallel": sequence_parallel,
        "model_init_config": model_init_config,
        "lora_config": lora_config,
        "mixed_precision_config": mixed_precision_config,
        "lnc_size": lnc_size,
    }
    if dist.is_initialized() and not ps.model_parallel_is_initialized():
        ps.initialize_model_parallel(
            tensor_model_parallel_size=tensor_parallel_size,
            pipeline_model_parallel_size=pipeline_parallel_size,
            expert_model_parallel_size=expert_parallel_size,
            context_parallel_size=context_parallel_size,
        )
    if dist.is_initialized() and ps.is_global_rank_zero():
        logger.info("NxD config:\n%s", pformat(config))
    return config


def initialize_parallel_model(nxd_config: dict, model_fn: Callable, include_buffers: bool = False, *model_args, **model_kwargs):
    """Instantiate (optionally on the meta device), pipeline-partition, move to the device in
    staggered waves, apply LoRA / head padding / activation checkpointing, wrap in
    :class:`NxDModel` (reference trainer.py:147-234)."""
    from ..utils import model_utils
    from ..utils.activation_checkpoint import apply_activation_checkpointing

    meta = nxd_config["model_init_config"]["meta_device_init"]
    if meta:
        with model_utils.init_on_device(torch.device("meta"), include_buffers=include_buffers):
            model = model_fn(*model_args, **model_kwargs)
    else:
        model = model_utils.get_model_sequential(
            lambda: model_fn(*model_args, **model_kwargs), nxd_config["model_init_config"]["sequential_move_factor"],
            move_to_device=False,
        )
    if nxd_config["pipeline_parallel_size"] > 1:
        from ..pipeline.model import NxDPPModel

        pcfg = dict(nxd_config["pipeline_config"] or {})
        pcfg.setdefault("param_init_fn", nxd_config["model_init_config"].get("param_init_fn"))
        if nxd_config["optimizer_config"]["zero_one_enabled"]:
            pcfg.setdefault("use_zero1_optimizer", True)
        model = NxDPPModel(model, **pcfg)
    if nxd_config.get("lora_config") is not None:
        from ..modules.lora import LoraModel

        model = LoraModel(model, nxd_config["lora_config"])
    if nxd_config.get("pad_model"):
        from ..parallel_layers.pad import pad_model

        model = pad_model(model, ps.get_tensor_model_parallel_size(), _num_heads_of(model))
    if meta and nxd_config["pipeline_parallel_size"] == 1:
        model_utils.reinit_model(model, get_device(), nxd_config["model_init_config"]["param_init_fn"])
    if not getattr(model, "_nxd_on_device", False):
        from ..parallel_layers.utils import move_model_to_device

        if nxd_config["pipeline_parallel_size"] > 1:
            model.move_model_to_device()
        else:
            move_model_to_device(model, get_device())
    nxd_model = NxDModel(model, nxd_config)
    acc = nxd_config.get("activation_checkpoint_config")
    if acc is not None:
        if acc == "full":
            target = getattr(nxd_model.original_module(), "_no_split_modules", None)
            if nxd_config["pipeline_parallel_size"] > 1:
                cls = model.transformer_layer_cls
                apply_activation_checkpointing(nxd_model, check_fn=lambda m: isinstance(m, cls))
            elif target:
                names = set(target)
                apply_activation_checkpointing(nxd_model, check_fn=lambda m: type(m).__name__ in names)
            else:
                raise ValueError("activation_checkpoint_config='full' needs `_no_split_modules` on the model")
        elif isinstance(acc, (tuple, list)):
            classes = tuple(acc)
            apply_activation_checkpointing(nxd_model, check_fn=lambda m: isinstance(m, classes))
        elif isinstance(acc, type):
            apply_activation_checkpointing(nxd_model, check_fn=lambda m: isinstance(m, acc))
        else:
            raise ValueError(f"unsupported activation_checkpoint_config: {acc!r}")
    hooks.execute_all_hooks(nxd_model)
    return nxd_model


def _num_heads_of(model) -> int:
    cfg = getattr(model, "config", None)
    for name in ("num_attention_heads", "n_head", "num_heads"):
        if cfg is not None and hasattr(cfg, name):
            return getattr(cfg, name)
    raise ValueError("pad_model=True needs model.config.num_attention_heads")


def initialize_optimizer_from_class(nxd_config: dict, optimizer_class, parameters, model=None, **defaults):
    ocfg, mp = nxd_config["optimizer_config"], nxd_config["mixed_precision_config"]
    if ocfg["zero_one_enabled"]:
        ep = ps.get_expert_model_parallel_size() > 1
        cls = NeuronEPZero1Optimizer if ep else NeuronZero1Optimizer
        cap = int(os.getenv("ALL_GATHER_REDUCE_SCATTER_BUCKET_CAP_MB", _RS_AG_BUCKET_CAP_MB))
        zcfg = dict(
            grad_clipping=ocfg["grad_clipping"],
            max_norm=ocfg.get("max_grad_norm", 1.0),
            sharding_groups=ps.get_zero1_sharding_groups() if ps.get_context_model_parallel_size() > 1
            else ps.get_data_parallel_replica_groups(),
            grad_norm_groups=ps.get_tensor_model_parallel_replica_groups(),
            bucket_cap_mb_reduce_scatter=cap,
            bucket_cap_mb_all_gather=max(1, cap // ps.get_data_parallel_size()),
            use_master_weights=mp["use_master_weights"],
            use_grad_acc_hook=mp["use_fp32_grad_acc"],
            higher_cc_precision=mp["use_fp32_grad_acc"],
            save_master_weights=bool(mp["use_master_weights_in_ckpt"]),
        )
        defaults.pop("decoupled_weight_decay", None)
        return cls(parameters, optimizer_class, **zcfg, **defaults)
    if mp["use_master_weights"]:
        raise RuntimeError("ZeRO-1 optimizer is not enabled, while `use_master_weights` is True.")
    if mp["use_fp32_grad_acc"] or mp["use_master_weights_in_ckpt"]:
        raise RuntimeError("Non Zero-1 optimizer does not support `use_fp32_grad_acc` or `use_master_weights_in_ckpt`.")
    return optimizer_class(parameters, **defaults)


def initialize_parallel_optimizer(nxd_config: dict, optimizer_class, parameters, **defaults) -> NxDOptimizer:
    params = list(parameters)
    return NxDOptimizer(initialize_optimizer_from_class(nxd_config, optimizer_class, params, **defaults), nxd_config)


def filter_to_local_parameter_group(optimizer, model) -> None:
    """Pipeline-parallel models own only their stage's parameters: rewrite ``optimizer.param_groups`` in place so each group
    keeps just the parameters that are materialised on this rank (reference trainer.py:325-335).  Works with the mapping a
    partitioned model exposes (``meta_device_parameter_map``: original → local parameter) and, without one, by dropping
    parameters that are still on the meta device."""
    mapping = getattr(model, "meta_device_parameter_map", None)
    for group in optimizer.param_groups:
        kept = []
        for p in group["params"]:
            q = mapping.get(p, None) if mapping is not None else p
            if q is not None and q.device.type != "meta":
                kept.append(q)
        group["params"] = kept
