"""Sanity checks of a training config's MoE section (reference ``modules/moe/moe_config_validator.py:14-136``):
dropless mode needs a SiLU-family GLU MLP and ``capacity_factor == 0``; token dropping needs a positive capacity factor."""
from __future__ import annotations

import logging
from typing import Any, Dict

from ...utils.utils import get_dict_from_json

logger = logging.getLogger(__name__)


class MoeConfigValidator:
    """``cfg`` is the (OmegaConf-like) training config: ``cfg.model_source`` ∈ {"hf", "megatron"}, ``cfg.model.moe`` with
    ``dropless`` / ``capacity_factor`` / ``glu_mlp``, ``cfg.model.model_config`` (path of the HF ``config.json``) or
    ``cfg.model.activation`` (Megatron)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.hf_model_config: Dict[Any, Any] = {}

    def _load_hf_config(self) -> Dict[Any, Any]:
        return get_dict_from_json(self.cfg.model.model_config)

    def _validate_hf_activation(self, dropless: bool) -> None:
        if not dropless:
            return
        if self.hf_model_config.get("model_type") == "dbrx":
            act = self.hf_model_config.get("ffn_config", {}).get("ffn_act_fn", {}).get("name")
            if act != "silu":
                raise ValueError("For DBRX models, dropless mode is only supported with SiLU activation function. "
                                 f"Current activation function: {act}. Please adjust your configuration.")
        elif self.hf_model_config.get("hidden_act") != "silu":
            raise ValueError("Dropless mode is only supported with SiLU activation function. Current activation "
                             f"function: {self.hf_model_config.get('hidden_act')}. Please adjust your configuration.")

    def _validate_megatron_activation(self, dropless: bool) -> None:
        if not dropless:
            return
        act = getattr(self.cfg.model, "activation", None)
        if act not in ("silu", "swiglu"):
            raise ValueError("For Megatron models, dropless mode is only supported with SiLU or SwiGLU activation "
                             f"functions. Current activation function: {act}. Please adjust your configuration.")

    def validate_moe_config(self) -> None:
        if not hasattr(self.cfg.model, "moe"):
            raise AttributeError("MoE configuration is missing in model config. Please ensure 'moe' attribute is present "
                                 "in the model configuration.")
        moe = self.cfg.model.moe
        dropless = getattr(moe, "dropless", False)
        capacity_factor = moe.capacity_factor
        glu_mlp = getattr(moe, "glu_mlp", True)
        if self.cfg.model_source == "hf":
            self.hf_model_config = self._load_hf_config()
            self._validate_hf_activation(dropless)
        elif self.cfg.model_source == "megatron":
            self._validate_megatron_activation(dropless)
        if dropless:
            if not glu_mlp:
                raise ValueError("Dropless mode requires GLU_MLP to be True.")
            if capacity_factor is None or capacity_factor > 0.0:
                logger.warning("Dropless mode expects a capacity_factor set to 0.0. Current value: %s. Setting "
                               "capacity_factor to 0.0.", capacity_factor)
                moe.capacity_factor = 0.0
        elif capacity_factor is not None and capacity_factor <= 0.0:
            raise ValueError("Dropping requires a capacity factor greater than 0.0 Please adjust your configuration.")
