"""``LoraModel`` — inject adapters into a model, save / load them, merge / unmerge
(reference ``modules/lora/model.py:74-746``).  Adapter checkpoints use HF-PEFT-compatible key names
(``base_model.model.<module>.lora_A.weight``) plus ``adapter_config.json`` content under ``"lora_config"``."""
from __future__ import annotations

import os
import re
from typing import Any, Dict, List, Optional

import torch
from torch import nn

from ...parallel_layers.layers import ColumnParallelLinear, RowParallelLinear
from ..qkv_linear import GQAQKVColumnParallelLinear
from .config import LoraConfig
from .layer import LoraConv2d, LoraEmbedding, LoraLayer, LoraLinear
from .tp_layer import LoraGQAQKVParallelLinear, LoraParallelLinear

_DEFAULT_TARGETS = ["q_proj", "k_proj", "v_proj", "o_proj", "qkv_proj", "gate_up_proj", "down_proj", "up_proj", "gate_proj"]


def _by_targets(groups):
    return {arch: list(t) for t, archs in groups for arch in archs}


# Default adapter placement per Hugging Face ``config.model_type`` when ``LoraConfig.target_modules`` is not given (the table
# of reference lora/model.py:35-67, which follows peft): attention query / value projections, or the fused QKV where the
# architecture has one.
MODELS_TO_LORA_TARGET_MODULES_MAPPING = _by_targets([
    (("q_proj", "v_proj"), ("llama", "mistral", "mixtral", "gemma", "stablelm", "opt", "gptj", "gpt_neo", "bart")),
    (("query_key_value",), ("gpt_neox", "bloom", "falcon", "chatglm", "RefinedWeb", "RefinedWebModel")),
    (("query", "value"), ("bert", "roberta", "xlm-roberta", "electra", "layoutlm")),
    (("q", "v"), ("t5", "mt5")),
    (("c_attn",), ("gpt2", "gpt_bigcode")),
    (("q", "v", "q_proj", "v_proj"), ("blip-2",)),
    (("query_proj", "value_proj"), ("deberta-v2",)),
    (("in_proj",), ("deberta",)),
    (("Wqkv",), ("mpt",)),
    (("c_proj", "c_attn"), ("btlm",)),
    (("qkv_proj",), ("codegen",)),
    (("q_proj", "v_proj", "fc1", "fc2"), ("phi",)),
])


def _wrap(module: nn.Module, cfg: LoraConfig) -> Optional[nn.Module]:
    if isinstance(module, GQAQKVColumnParallelLinear):
        return LoraGQAQKVParallelLinear(module, cfg)
    if isinstance(module, (ColumnParallelLinear, RowParallelLinear)):
        return LoraParallelLinear(module, cfg)
    if isinstance(module, nn.Linear):
        return LoraLinear(module, cfg)
    if isinstance(module, nn.Embedding):
        return LoraEmbedding(module, cfg)
    if isinstance(module, nn.Conv2d):
        return LoraConv2d(module, cfg)
    return None


WEIGHTS_NAME = "adapter_model.pt"
CONFIG_NAME = "adapter_config.json"


class LoraModel(nn.Module):
    ColumnParallelLinear_lora_type = "ColumnParallelLinear"
    RowParallelLinear_lora_type = "RowParallelLinear"
    GQAQKVParallelLinear_lora_type = "GQAQKVParallelLinear"

    def __init__(self, module: nn.Module, config: LoraConfig):
        super().__init__()
        self.module = module
        self.lora_config = config
        self.modules_to_save: List[str] = list(config.modules_to_save or [])
        self.lora_module_parallel_types: Dict[str, str] = {}
        self.lora_ckpt: Optional[Dict[str, Any]] = None
        self.is_config_saved = self.is_checkpoint_loaded = self.is_base_model_loaded = False
        if config.enable_lora:
            self.inject_adapter()
        if config.load_lora_from_ckpt and config.lora_save_path and os.path.isdir(config.lora_save_path):
            self.load_lora(config.lora_save_path, config.lora_load_tag)

    # reference-named flags
    @property
    def is_lora_enabled(self) -> bool:
        return any(isinstance(m, LoraLayer) for m in self.module.modules())

    @property
    def is_lora_merged(self) -> bool:
        return any(isinstance(m, LoraLayer) and m.merged for m in self.module.modules())

    @property
    def is_verbose_enabled(self) -> bool:
        return bool(self.lora_config.lora_verbose)

    # ------------------------------------------------------------------ injection
    def _is_target(self, name: str) -> bool:
        t = self.lora_config.target_modules
        if t is None:
            t = self._default_targets()
        if isinstance(t, str):
            return re.fullmatch(t, name) is not None
        return any(name == x or name.endswith("." + x) for x in t)

    def _default_targets(self):
        """No ``target_modules``: the per-architecture default when the base model says what it is (``config.model_type``) and
        that default matches modules of this model (a fused-QKV build of a "q_proj / v_proj" architecture has ``qkv_proj``
        instead); otherwise every projection of the attention and MLP blocks."""
        cached = self.__dict__.get("_auto_targets")
        if cached is None:
            cfg = getattr(self.module, "config", None)
            mt = (cfg.get("model_type") if isinstance(cfg, dict) else getattr(cfg, "model_type", None)) or "unknown"
            want = MODELS_TO_LORA_TARGET_MODULES_MAPPING.get(mt) or MODELS_TO_LORA_TARGET_MODULES_MAPPING.get(str(mt).lower())
            names = [n for n, _ in self.module.named_modules()]
            if want and any(n == x or n.endswith("." + x) for n in names for x in want):
                cached = list(want)
            else:
                cached = _DEFAULT_TARGETS
            self.__dict__["_auto_targets"] = cached
        return cached

    def inject_adapter(self) -> None:
        replaced = 0
        for name, child in list(self.module.named_modules()):
            if not name or isinstance(child, LoraLayer) or not self._is_target(name):
                continue
            wrapped = _wrap(child, self.lora_config)
            if wrapped is None:
                continue
            parent_name, _, leaf = name.rpartition(".")
            parent = self.module.get_submodule(parent_name) if parent_name else self.module
            setattr(parent, leaf, wrapped)
            if isinstance(wrapped, LoraGQAQKVParallelLinear):
                self.lora_module_parallel_types[name] = self.GQAQKVParallelLinear_lora_type
                self.lora_kv_size_multiplier = getattr(child, "kv_size_multiplier", 1)
            elif isinstance(child, ColumnParallelLinear):
                self.lora_module_parallel_types[name] = self.ColumnParallelLinear_lora_type
            elif isinstance(child, RowParallelLinear):
                self.lora_module_parallel_types[name] = self.RowParallelLinear_lora_type
            replaced += 1
        if replaced == 0:
            raise ValueError(f"no module matched LoRA target_modules={self.lora_config.target_modules}")
        self.mark_only_lora_as_trainable()

    def mark_only_lora_as_trainable(self) -> None:
        bias = self.lora_config.bias
        for n, p in self.module.named_parameters():
            is_lora = "lora_" in n
            keep = any(t in n for t in self.modules_to_save)
            p.requires_grad_(is_lora or keep or (bias == "all" and n.endswith("bias")))
        if bias == "lora_only":
            for m in self.module.modules():
                if isinstance(m, LoraLayer) and getattr(m.base_layer, "bias", None) is not None:
                    m.base_layer.bias.requires_grad_(True)

    # ------------------------------------------------------------------ forward / passthrough
    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def __getattr__(self, name: str) -> Any:
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("module"), name)

    # ------------------------------------------------------------------ merge
    def merge_lora(self) -> None:
        for m in self.module.modules():
            if isinstance(m, LoraLayer) and not isinstance(m, LoraGQAQKVParallelLinear):
                m.merge()

    def unmerge_lora(self) -> None:
        for m in self.module.modules():
            if isinstance(m, LoraLayer) and not isinstance(m, LoraGQAQKVParallelLinear):
                m.unmerge()

    # ------------------------------------------------------------------ state dicts
    def lora_state_dict(self) -> Dict[str, torch.Tensor]:
        sd = {}
        for k, v in self.module.state_dict().items():
            if "lora_" in k:
                sd["base_model.model." + k] = v
        return sd

    def base_state_dict(self) -> Dict[str, torch.Tensor]:
        return {k.replace(".base_layer", ""): v for k, v in self.module.state_dict().items() if "lora_" not in k}

    def save_lora(self, save_dir: Optional[str] = None, adapter_tag: Optional[str] = None, path: Optional[str] = None,
                  tag: Optional[str] = None) -> str:
        """Write this rank's adapter file under ``save_dir[/adapter_tag]`` (argument names of reference lora/model.py:461;
        ``path`` / ``tag`` are this package's earlier names).  Unlike the reference — which refuses here once model parallelism
        is initialised and sends the user to ``nxd.save_checkpoint`` — every (tp, pp) rank writes its own shard file."""
        from ...parallel_layers import parallel_state as ps

        path = save_dir or path or getattr(self.lora_config, "lora_save_dir", None) or getattr(self.lora_config, "lora_save_path", None)
        tag = adapter_tag or tag
        assert path is not None, "no save_dir given and lora_config.lora_save_dir is not set"
        d = os.path.join(path, tag) if tag else path
        os.makedirs(d, exist_ok=True)
        tp = ps.get_tensor_model_parallel_rank() if ps.model_parallel_is_initialized() else 0
        pp = ps.get_pipeline_model_parallel_rank() if ps.model_parallel_is_initialized() else 0
        f = os.path.join(d, f"adapter_tp_rank_{tp:02d}_pp_rank_{pp:02d}.pt")
        torch.save({"lora_config": self.lora_config.to_dict(), "state_dict": {k: v.cpu() for k, v in self.lora_state_dict().items()}}, f)
        return f

    def load_lora(self, save_dir: Optional[str] = None, adapter_tag: Optional[str] = None, ckpt_path: Optional[str] = None,
                  adapter_only: bool = True, path: Optional[str] = None, tag: Optional[str] = None) -> None:
        """Load this rank's adapter file from ``save_dir[/adapter_tag]`` (reference lora/model.py:555-595).  With
        ``adapter_only=False`` the base model's weights are loaded first from ``ckpt_path`` (a plain ``state_dict`` file;
        default ``<save_dir>[/<tag>]/adapter_model.pt``) unless the adapter checkpoint itself carries the base
        (``save_lora_base``)."""
        from ...parallel_layers import parallel_state as ps

        path = save_dir or path or getattr(self.lora_config, "lora_save_dir", None)
        tag = adapter_tag or tag
        assert path is not None, "no save_dir given and lora_config.lora_save_dir is not set"
        if not adapter_only and not self.lora_config.save_lora_base:
            base = ckpt_path or os.path.join(os.path.join(path, tag) if tag else path, WEIGHTS_NAME)
            if not os.path.exists(base):
                raise FileNotFoundError(f"The checkpoint file {base} is not found.")
            self.load_state_dict(torch.load(base, map_location="cpu", weights_only=False), strict=False)
        self.is_base_model_loaded = True
        d = os.path.join(path, tag) if tag else path
        tp = ps.get_tensor_model_parallel_rank() if ps.model_parallel_is_initialized() else 0
        pp = ps.get_pipeline_model_parallel_rank() if ps.model_parallel_is_initialized() else 0
        ck = torch.load(os.path.join(d, f"adapter_tp_rank_{tp:02d}_pp_rank_{pp:02d}.pt"), map_location="cpu", weights_only=False)
        sd = {k[len("base_model.model."):]: v for k, v in ck["state_dict"].items()}
        missing, unexpected = self.module.load_state_dict(sd, strict=False)
        bad = [k for k in unexpected]
        if bad:
            raise RuntimeError(f"unexpected adapter keys: {bad}")

    def state_dict(self, *a, **k):
        if self.lora_config.save_lora_base:
            return self.module.state_dict(*a, **k)
        sd = self.lora_state_dict()
        for key, v in self.module.state_dict().items():                       # extra trainable modules travel with the adapter
            if self.modules_to_save and any(t in key for t in self.modules_to_save):
                sd["base_model.model." + key] = v
        if self.lora_config.merge_sharded_lora:
            sd = self.merge_sharded_lora_weights(sd)
        return sd

    # ------------------------------------------------------------------ reference-named surface (model.py:366-746)
    def get_base_model(self) -> nn.Module:
        return self.module

    def generate(self, *args, **kwargs):
        return self.module.generate(*args, **kwargs)

    def named_parameters(self, *args, **kwargs):
        yield from self.module.named_parameters(*args, **kwargs)

    def module_state_dict(self) -> Dict[str, Any]:
        return self.module.state_dict()

    @property
    def dtype(self) -> torch.dtype:
        return next(self.module.parameters()).dtype

    @property
    def config(self):
        return getattr(self.module, "config", None) or self.lora_config

    def update_state_dict_keys(self, state_dict: Dict[str, Any]) -> Dict[str, Any]:
        """Keys of the un-adapted model → keys of the adapted one (``x.weight`` → ``x.base_layer.weight``), in place."""
        for mkey in self.module_state_dict():
            if ".base_layer" in mkey:
                plain = mkey.replace(".base_layer", "")
                if plain in state_dict:
                    state_dict[mkey] = state_dict.pop(plain)
        return state_dict

    def merge_sharded_lora_weights(self, state_dict: Dict[str, Any]) -> Dict[str, Any]:
        """Turn TP-sharded adapter halves into full matrices (all-gather over the TP group): ``lora_B`` of column-parallel
        bases along dim 0, ``lora_A`` of row-parallel bases along the last dim, GQA-QKV ``lora_B`` along dim 0 minus the
        KV replicas — the result is a single-device (HF-PEFT loadable) adapter."""
        from ...parallel_layers import comm
        from ...parallel_layers import parallel_state as ps

        if not ps.model_parallel_is_initialized() or ps.get_tensor_model_parallel_size() == 1:
            return state_dict
        for name, w in list(state_dict.items()):
            if name == "lora_config" or ".lora_" not in name or not isinstance(w, torch.Tensor):
                continue
            base = name.split(".lora_")[0]
            base = base[len("base_model.model."):] if base.startswith("base_model.model.") else base
            kind = self.lora_module_parallel_types.get(base)
            if kind == self.ColumnParallelLinear_lora_type and ".lora_B" in name:
                state_dict[name] = comm.all_gather(w.contiguous(), dim=0)
            elif kind == self.RowParallelLinear_lora_type and ".lora_A" in name:
                state_dict[name] = comm.all_gather(w.contiguous(), dim=w.dim() - 1)
            elif kind == self.GQAQKVParallelLinear_lora_type and ".lora_B" in name:
                full = comm.all_gather(w.contiguous(), dim=0)
                state_dict[name] = torch.chunk(full, getattr(self, "lora_kv_size_multiplier", 1))[0] \
                    if ("lora_B_k" in name or "lora_B_v" in name) else full
        return state_dict

    def save_config(self, save_dir: Optional[str] = None) -> str:
        """``adapter_config.json`` with the adapter-defining fields (HF-PEFT compatible ``r`` included)."""
        import json

        d = save_dir or self.lora_config.lora_save_dir
        assert d, "no save directory"
        os.makedirs(d, exist_ok=True)
        f = os.path.join(d, CONFIG_NAME)
        with open(f, "w") as w:
            json.dump(self.lora_config.selected_fields_to_save(), w, indent=2, sort_keys=True)
        self.is_config_saved = True
        return f

    def load_checkpoint(self, lora_config: LoraConfig) -> None:
        """Parse the adapter checkpoint named by ``lora_config`` (single-device file ``<dir>[/<tag>]/adapter_model.pt`` or this
        rank's ``adapter_tp_rank_XX_pp_rank_YY.pt``): adopt the stored adapter configuration and keep the tensors for
        :meth:`load_lora_adapter`."""
        import json
        from dataclasses import replace

        d = lora_config.lora_save_dir
        assert d, "lora_save_dir is not set"
        out = os.path.join(d, lora_config.lora_load_tag) if lora_config.lora_load_tag else d
        cands = [os.path.join(out, WEIGHTS_NAME)]
        from ...parallel_layers import parallel_state as ps
        if ps.model_parallel_is_initialized():
            cands.insert(0, os.path.join(out, f"adapter_tp_rank_{ps.get_tensor_model_parallel_rank():02d}_pp_rank_"
                                              f"{ps.get_pipeline_model_parallel_rank():02d}.pt"))
        path = next((c for c in cands if os.path.isfile(c)), None)
        if path is None:
            raise FileNotFoundError(f"{cands[-1]} is not found.")
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        tensors = ckpt.get("state_dict", ckpt)
        stored = ckpt.get("lora_config")
        if stored is None:
            cfg_file = os.path.join(d, CONFIG_NAME)
            if not os.path.isfile(cfg_file):
                raise FileNotFoundError(f"Please name the file for LoRA confiugration as {CONFIG_NAME}.")
            stored = json.load(open(cfg_file))
        fields = {k: v for k, v in stored.items() if k in LoraConfig.__dataclass_fields__}
        self.lora_config = replace(lora_config, **fields)
        self.lora_ckpt = {k: v for k, v in tensors.items() if k != "lora_config"}
        self.is_checkpoint_loaded = True

    def load_lora_adapter(self):
        """Inject adapters (if not yet) and load the tensors parsed by :meth:`load_checkpoint`."""
        assert self.lora_ckpt is not None, "call load_checkpoint() first"
        if not self.is_lora_enabled and not (self.lora_config.save_lora_base and self.lora_config.merge_lora):
            self.inject_adapter()
        sd = {(k[len("base_model.model."):] if k.startswith("base_model.model.") else k): v for k, v in self.lora_ckpt.items()}
        res = self.module.load_state_dict(sd, strict=False)
        self.print_model_info()
        return res

    def get_nb_trainable_parameters(self):
        trainable = sum(p.numel() for p in self.module.parameters() if p.requires_grad)
        return trainable, sum(p.numel() for p in self.module.parameters())

    def print_trainable_parameters(self) -> None:
        from ...utils.logger import get_logger

        t, a = self.get_nb_trainable_parameters()
        get_logger().info(f"trainable params: {t:,d} || all params: {a:,d} || trainable%: {100 * t / max(a, 1)}")

    def print_model_info(self) -> None:
        if self.is_verbose_enabled:
            from ...utils.logger import get_logger

            get_logger().info("LoRA model: %s", self.module)
            get_logger().info("LoRA configuration: %s", self.lora_config)
            self.print_trainable_parameters()

    def load_state_dict(self, state_dict=None, strict: bool = False, assign: bool = False):
        """Adapter and / or base tensors into the wrapped module.  Keys of a base-model checkpoint written before the adapters
        were injected (``….weight`` where the module now holds ``….base_layer.weight``) are redirected, as the reference's
        ``update_state_dict_keys`` does; missing adapter or base entries are tolerated (``strict`` only reports them in the
        returned result)."""
        sd = {(k[len("base_model.model."):] if k.startswith("base_model.model.") else k): v for k, v in (state_dict or {}).items()}
        have = set(self.module.state_dict().keys())
        for key in list(sd):
            if key not in have:
                head, _, leaf = key.rpartition(".")
                moved = f"{head}.base_layer.{leaf}" if head else key
                if moved in have:
                    sd[moved] = sd.pop(key)
        return self.module.load_state_dict(sd, strict=False, assign=assign)


def get_lora_model(model: nn.Module, lora_config: LoraConfig) -> LoraModel:
    return LoraModel(model, lora_config)
