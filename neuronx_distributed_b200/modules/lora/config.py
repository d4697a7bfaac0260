"""LoRA configuration (reference ``modules/lora/config.py:6-148``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Union


@dataclass
class LoraConfig:
    enable_lora: bool = True
    lora_rank: int = 16
    lora_alpha: float = 8.0                         # reference default (config.py:53)
    lora_dropout: float = 0.0
    bias: str = "none"                              # "none" | "all" | "lora_only"
    target_modules: Optional[Union[List[str], str]] = None   # names / suffixes, or one regex string
    use_rslora: bool = False
    init_lora_weights: Union[bool, str] = "default"  # "default" (kaiming A, zero B) | "gaussian"
    modules_to_save: Optional[List[str]] = None      # non-LoRA modules that stay trainable and are saved with the adapter
    lora_verbose: bool = False
    # checkpointing
    load_lora_from_ckpt: bool = False
    save_lora_base: bool = False
    merge_lora: bool = False
    save_lora_config_adapter: bool = True            # config inside the adapter file (False → adapter_config.json next to it)
    lora_save_dir: Optional[str] = "lora_adapter"
    lora_save_path: Optional[str] = None             # older name in this package; wins over ``lora_save_dir`` when set
    lora_load_tag: Optional[str] = None
    merge_sharded_lora: bool = False                 # gather TP-sharded adapter halves into full matrices on ``state_dict()``
    # multi-adapter serving
    max_loras: int = 1
    lora_memory_transpose: bool = False

    @property
    def scaling(self) -> float:
        return self.lora_alpha / (self.lora_rank ** 0.5 if self.use_rslora else self.lora_rank)

    def __post_init__(self) -> None:
        if isinstance(self.target_modules, list):
            self.target_modules = sorted(set(self.target_modules))
        if self.lora_save_path is None:
            self.lora_save_path = self.lora_save_dir
        else:
            self.lora_save_dir = self.lora_save_path

    def to_dict(self) -> Dict:
        return dict(self.__dict__)

    @staticmethod
    def get_selected_fields() -> List[str]:
        """Fields that define the adapter (saved with it; everything else is run-time policy) — reference :120-133."""
        return ["bias", "init_lora_weights", "lora_alpha", "lora_dropout", "lora_rank", "use_rslora", "target_modules",
                "modules_to_save", "save_lora_base", "merge_lora", "save_lora_config_adapter"]

    def selected_fields_to_save(self) -> Dict:
        d = {k: v for k, v in self.to_dict().items() if k in self.get_selected_fields()}
        d["r"] = d["lora_rank"]                       # HF-PEFT spelling
        return d
