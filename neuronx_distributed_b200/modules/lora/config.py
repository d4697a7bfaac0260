"""LoRA configuration (reference ``modules/lora/config.py:6-148``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Union


@dataclass
class LoraConfig:
    enable_lora: bool = True
    lora_rank: int = 16
    lora_alpha: float = 32.0
    lora_dropout: float = 0.0
    bias: str = "none"                              # "none" | "all" | "lora_only"
    target_modules: Optional[Union[List[str], str]] = None   # names / suffixes, or one regex string
    use_rslora: bool = False
    init_lora_weights: Union[bool, str] = "default"  # "default" (kaiming A, zero B) | "gaussian"
    lora_verbose: bool = False
    # checkpointing
    load_lora_from_ckpt: bool = False
    save_lora_base: bool = False
    merge_lora: bool = False
    save_lora_config_adapter: bool = True
    lora_save_path: Optional[str] = None
    lora_load_tag: Optional[str] = None
    merge_sharded_lora: bool = False
    # multi-adapter serving
    max_loras: int = 1
    lora_memory_transpose: bool = False

    @property
    def scaling(self) -> float:
        return self.lora_alpha / (self.lora_rank ** 0.5 if self.use_rslora else self.lora_rank)

    def to_dict(self) -> Dict:
        return dict(self.__dict__)
