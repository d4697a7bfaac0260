"""Turn a training-launcher YAML into the small ``config.json`` the checkpoint converter reads
(reference ``scripts/yaml_converter.py:5-40``): ``model.num_layers / num_attention_heads / hidden_size / num_kv_heads``
(and ``model.moe.num_experts`` → ``num_local_experts``)."""
from __future__ import annotations

import json
from typing import Any, Dict, Optional


def load_yaml_file(file_path: str) -> Optional[Dict[str, Any]]:
    """Parsed YAML, or ``None`` (with a message) when the file is missing or malformed."""
    import yaml

    try:
        with open(file_path, "r") as f:
            return yaml.safe_load(f)
    except FileNotFoundError:
        print(f"Error: File '{file_path}' not found.")
    except yaml.YAMLError as e:
        print(f"Error parsing YAML file: {e}")
    return None


def convert_yaml_to_json(yaml_path: str, filename: str = "yaml_config.json") -> str:
    data = load_yaml_file(yaml_path)
    if data is None or "model" not in data:
        raise ValueError(f"{yaml_path} is not a readable training config with a 'model' section")
    m = data["model"]
    cfg = {"num_hidden_layers": m["num_layers"], "num_attention_heads": m["num_attention_heads"], "hidden_size": m["hidden_size"],
           "num_key_value_heads": m.get("num_kv_heads", m["num_attention_heads"])}
    if "moe" in m:
        cfg["num_local_experts"] = m["moe"]["num_experts"]
    with open(filename, "w") as f:
        json.dump(cfg, f)
    return filename
