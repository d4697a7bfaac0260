"""Model hyper-parameters of the sample (Llama-3.2-1B shapes by default; ``tiny()`` for CPU smoke runs)."""
from dataclasses import dataclass

import torch


@dataclass
class Config:
    dim: int = 2048
    n_layers: int = 16
    n_heads: int = 32
    n_kv_heads: int = 8
    vocab_size: int = 128256
    ffn_dim_multiplier: float = 1.5
    multiple_of: int = 256
    norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    use_scaled_rope: bool = True
    max_batch_size: int = 2
    max_seq_len: int = 128
    dtype: torch.dtype = torch.bfloat16

    @property
    def head_dim(self) -> int:
        return self.dim // self.n_heads

    @property
    def hidden_dim(self) -> int:
        h = int(2 * (4 * self.dim) / 3)
        h = int(self.ffn_dim_multiplier * h)
        return self.multiple_of * ((h + self.multiple_of - 1) // self.multiple_of)

    @staticmethod
    def tiny(**kw) -> "Config":
        base = dict(dim=64, n_layers=2, n_heads=4, n_kv_heads=2, vocab_size=260, multiple_of=32, max_seq_len=32,
                    use_scaled_rope=False, dtype=torch.float32)
        base.update(kw)
        return Config(**base)
