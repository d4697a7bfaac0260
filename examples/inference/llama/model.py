"""A self-contained Llama-3 style decoder written against the public building blocks of the package — the counterpart of the
reference's ``examples/inference/llama/model.py``.

What to look at:

* tensor parallelism comes from three layer types only: ``ParallelEmbedding`` (vocab-sharded), ``ColumnParallelLinear``
  (output features sharded: wq / wk / wv / w1 / w3 / the LM head) and ``RowParallelLinear`` (input features sharded, result
  all-reduced: wo / w2).  Everything between a Column and a Row layer works on this rank's heads / features;
* the KV cache is a pair of *buffers* per layer.  ``forward`` writes them in place, which is what makes them "state" for the
  inference builder: every bucket (prefill, decode) is traced from the same module and shares the same cache tensors;
* one ``forward`` serves both phases, selected by the input shape — ``[B, S]`` prompt tokens (S > 1: prefill, causal flash
  attention over the prompt, cache slots 0..S-1 written) or ``[B, 1]`` (decode: one slot written at ``last_pos``, attention over
  the cache with the split-KV kernel).  The builder captures each shape as its own CUDA graph.
"""
import torch
from torch import nn

from neuronx_distributed_b200 import ops
from neuronx_distributed_b200.modules.attention.utils import precompute_freqs_cis
from neuronx_distributed_b200.parallel_layers import parallel_state as ps
from neuronx_distributed_b200.parallel_layers.layers import ColumnParallelLinear, ParallelEmbedding, RowParallelLinear

from config import Config


def rotate(x: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """Interleaved-pair RoPE (Meta checkpoint convention).  ``x`` [B, S, H, D]; ``freqs`` [B, S, D/2, 2] = (cos, sin)."""
    xf = x.float().reshape(*x.shape[:-1], -1, 2)
    c, s = freqs[:, :, None, :, 0], freqs[:, :, None, :, 1]
    out = torch.stack([xf[..., 0] * c - xf[..., 1] * s, xf[..., 0] * s + xf[..., 1] * c], dim=-1)
    return out.flatten(-2).to(x.dtype)


class Attention(nn.Module):
    def __init__(self, cfg: Config):
        super().__init__()
        tp = ps.get_tensor_model_parallel_size()
        assert cfg.n_kv_heads % tp == 0, f"this sample shards KV heads: tp ({tp}) must divide n_kv_heads ({cfg.n_kv_heads})"
        self.n_heads, self.n_kv, self.hd = cfg.n_heads // tp, cfg.n_kv_heads // tp, cfg.head_dim
        kw = dict(bias=False, dtype=cfg.dtype)
        self.wq = ColumnParallelLinear(cfg.dim, cfg.n_heads * self.hd, gather_output=False, **kw)
        self.wk = ColumnParallelLinear(cfg.dim, cfg.n_kv_heads * self.hd, gather_output=False, **kw)
        self.wv = ColumnParallelLinear(cfg.dim, cfg.n_kv_heads * self.hd, gather_output=False, **kw)
        self.wo = RowParallelLinear(cfg.n_heads * self.hd, cfg.dim, input_is_parallel=True, **kw)
        shape = (cfg.max_batch_size, cfg.max_seq_len, self.n_kv, self.hd)
        self.register_buffer("cache_k", torch.zeros(shape, dtype=cfg.dtype), persistent=False)
        self.register_buffer("cache_v", torch.zeros(shape, dtype=cfg.dtype), persistent=False)

    def forward(self, x: torch.Tensor, freqs: torch.Tensor, last_pos: torch.Tensor) -> torch.Tensor:
        B, S, _ = x.shape
        q = self.wq(x).view(B, S, self.n_heads, self.hd)
        k = self.wk(x).view(B, S, self.n_kv, self.hd)
        v = self.wv(x).view(B, S, self.n_kv, self.hd)
        q, k = rotate(q, freqs), rotate(k, freqs)
        if S > 1:                                                      # prefill: slots 0..S-1
            self.cache_k[:B, :S].copy_(k)
            self.cache_v[:B, :S].copy_(v)
            o = ops.attention.flash_attention(q, k, v, causal=True)
        else:                                                          # decode: slot last_pos[b]
            b = torch.arange(B, device=x.device)
            self.cache_k[b, last_pos] = k[:, 0]
            self.cache_v[b, last_pos] = v[:, 0]
            o = ops.attention.decode_attention(q, self.cache_k[:B], self.cache_v[:B], last_pos)
        return self.wo(o.reshape(B, S, self.n_heads * self.hd))


class FeedForward(nn.Module):
    def __init__(self, cfg: Config):
        super().__init__()
        kw = dict(bias=False, dtype=cfg.dtype)
        self.w1 = ColumnParallelLinear(cfg.dim, cfg.hidden_dim, gather_output=False, **kw)
        self.w3 = ColumnParallelLinear(cfg.dim, cfg.hidden_dim, gather_output=False, **kw)
        self.w2 = RowParallelLinear(cfg.hidden_dim, cfg.dim, input_is_parallel=True, **kw)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.w2(torch.nn.functional.silu(self.w1(x)) * self.w3(x))


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float, dtype: torch.dtype):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim, dtype=dtype))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.norm.rms_norm(x, self.weight, self.eps)             # fp32 statistics, one kernel on CUDA


class TransformerBlock(nn.Module):
    def __init__(self, cfg: Config):
        super().__init__()
        self.attention, self.feed_forward = Attention(cfg), FeedForward(cfg)
        self.attention_norm = RMSNorm(cfg.dim, cfg.norm_eps, cfg.dtype)
        self.ffn_norm = RMSNorm(cfg.dim, cfg.norm_eps, cfg.dtype)

    def forward(self, x, freqs, last_pos):
        h = x + self.attention(self.attention_norm(x), freqs, last_pos)
        return h + self.feed_forward(self.ffn_norm(h))


class Transformer(nn.Module):
    def __init__(self, cfg: Config):
        super().__init__()
        self.cfg = cfg
        self.tok_embeddings = ParallelEmbedding(cfg.vocab_size, cfg.dim, dtype=cfg.dtype)
        self.layers = nn.ModuleList(TransformerBlock(cfg) for _ in range(cfg.n_layers))
        self.norm = RMSNorm(cfg.dim, cfg.norm_eps, cfg.dtype)
        self.output = ColumnParallelLinear(cfg.dim, cfg.vocab_size, bias=False, gather_output=True, dtype=cfg.dtype)
        self.register_buffer("freqs", precompute_freqs_cis(cfg.head_dim, cfg.max_seq_len, cfg.rope_theta, cfg.use_scaled_rope),
                             persistent=False)

    def forward(self, tokens: torch.Tensor, last_pos: torch.Tensor):
        """``tokens`` [B, S] (right-padded prompts) with ``last_pos`` [B] = index of each prompt's last real token, or
        ``tokens`` [B, 1] with ``last_pos`` [B] = the cache slot of that token.  Returns (greedy next token [B], logits [B, V])."""
        B, S = tokens.shape
        if S > 1:
            pos = torch.arange(S, device=tokens.device)[None, :].expand(B, S)
        else:
            pos = last_pos[:, None]
        freqs = self.freqs[pos]                                        # [B, S, D/2, 2]
        h = self.tok_embeddings(tokens)
        for layer in self.layers:
            h = layer(h, freqs, last_pos)
        h = self.norm(h)
        last = h[torch.arange(B, device=h.device), last_pos] if S > 1 else h[:, 0]
        logits = self.output(last).float()                             # [B, V] (all-gathered over TP)
        return logits.argmax(-1), logits


def meta_to_sample_state_dict(sd):
    """Meta ``consolidated.00.pth`` keys are already the names used here; only the rope table is not a parameter."""
    return {k: v for k, v in sd.items() if k != "rope.freqs"}
