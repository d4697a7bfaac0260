"""Llama inference model: context encoding (prefill) + token generation (decode) over a persistent KV cache, with
on-device sampling — role of reference ``examples/inference`` (``NeuronBaseModel`` / ``ModelWrapper`` /
``KVCacheManager`` / ``hf_adapter``), re-using the training model's parameter names so the same checkpoints load.
No sequence parallelism at inference: Row-parallel outputs are all-reduced."""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from .. import ops
from ..inference.kv_cache import KVCacheManager
from ..utils.sampling import Sampler
from .llama import LlamaConfig, LlamaForCausalLM


class LlamaForInference(nn.Module):
    def __init__(self, cfg: LlamaConfig, batch_size: int = 1, max_seq_len: int = 2048, on_device_sampling: bool = True,
                 sampler: Optional[Sampler] = None, lm_cls=None, flash_decoding: bool = False):
        """``lm_cls``: any decoder built from Llama attention blocks (``LlamaForCausalLM`` default; ``MixtralForCausalLM`` for
        the MoE families — its layers' ``mlp`` returns ``(y, router_logits)``, handled in ``_body``)."""
        super().__init__()
        cfg.sequence_parallel_enabled = False
        cfg.activation_checkpointing = "none"
        self.cfg = cfg
        self.lm = (lm_cls or LlamaForCausalLM)(cfg)
        self.batch_size, self.max_seq_len = batch_size, max_seq_len
        attn0 = self._core.layers[0].self_attn
        # flash decoding (reference examples/inference ``flash_decoding_enabled``): TP ranks that hold replicas of the same KV
        # heads shard the cache along the SEQUENCE inside that replica group instead of storing it ``m`` times
        self.flash_decoding, self.kv_group, self.kv_rank, shards = bool(flash_decoding), None, 0, 1
        if flash_decoding:
            import torch.distributed as dist

            self.kv_group = getattr(attn0.qkv_proj, "kv_group", None)
            assert self.kv_group is not None and dist.get_world_size(self.kv_group) > 1, \
                "flash decoding needs replicated KV heads (tensor-parallel size > number of KV heads)"
            shards, self.kv_rank = dist.get_world_size(self.kv_group), dist.get_rank(self.kv_group)
            assert max_seq_len % shards == 0, f"max_seq_len {max_seq_len} must divide by the KV replication factor {shards}"
        self.kv = KVCacheManager(cfg.num_hidden_layers, batch_size, max_seq_len, attn0.num_kv_heads_local, cfg.head_dim,
                                 dtype=cfg.dtype, device=cfg.device, num_cores_per_group=shards)
        self.sampler = sampler or Sampler(top_k=1, vocab_parallel=True)
        self.on_device_sampling = on_device_sampling
        cos, sin = ops.rope.rope_tables(max_seq_len, cfg.head_dim, cfg.rope_theta, cfg.device, 0, cfg.rope_scaling_factor)
        self.register_buffer("rope_cos", cos, persistent=False)
        self.register_buffer("rope_sin", sin, persistent=False)

    @property
    def _core(self):
        return getattr(self.lm, "model", self.lm)        # Llama nests the decoder under .model, Mixtral/DBRX do not

    def load_state_dict(self, sd, strict: bool = True):
        return self.lm.load_state_dict(sd, strict=strict)

    def state_dict(self, *a, **k):
        return self.lm.state_dict(*a, **k)

    # ------------------------------------------------------------------ shared pieces
    def _attn_block(self, layer_idx: int, layer, x: torch.Tensor, positions: Optional[torch.Tensor], prefill: bool,
                    kv_len: Optional[int], tree=None) -> torch.Tensor:
        att = layer.self_attn
        h = layer.input_layernorm(x)
        q, k, v = att.qkv_proj(h)                          # [S, B, h*D]
        S, B = q.shape[0], q.shape[1]
        D = att.head_dim
        q = q.reshape(S, B, att.num_heads_local, D).transpose(0, 1)
        k = k.reshape(S, B, att.num_kv_heads_local, D).transpose(0, 1)
        v = v.reshape(S, B, att.num_kv_heads_local, D).transpose(0, 1).contiguous()
        if self.flash_decoding and not prefill:
            assert S == 1 and tree is None, "flash decoding serves plain token generation (no speculation window)"
            from ..modules.attention.flash_decode import flash_decode_attention, write_decode_sharded

            cos = self.rope_cos[positions].unsqueeze(1)
            sin = self.rope_sin[positions].unsqueeze(1)
            q, k = _rope_per_batch(q, cos, sin), _rope_per_batch(k, cos, sin)
            write_decode_sharded(self.kv.k[layer_idx], self.kv.v[layer_idx], k, v, positions, self.kv_rank)   # owner-only write
            o = flash_decode_attention(q, self.kv.k[layer_idx], self.kv.v[layer_idx], positions, self.kv_group)
        elif not prefill and S > 1:
            # speculation window: W new tokens per sequence at positions p..p+W-1, causal among themselves, full cache before
            W = S
            # linear window: node w sits at sequence position p+w; Medusa tree: node w sits at p+depth[w] (its cache SLOT is
            # still p+w) and sees only its ancestors inside the window
            depth = torch.arange(W, device=q.device) if tree is None else tree[1]
            idx = (positions.unsqueeze(1) + depth.unsqueeze(0)).clamp(max=self.max_seq_len - 1)
            cos, sin = self.rope_cos[idx], self.rope_sin[idx]                     # [B, W, D/2]
            q, k = _rope_window(q, cos, sin), _rope_window(k, cos, sin)
            self.kv.write_window(layer_idx, k, v, positions)
            kc, vc = self.kv.get(layer_idx, kv_len)
            o = _window_attention(q, kc, vc, positions, None if tree is None else tree[0])
        elif prefill:
            cos, sin = self.rope_cos[:S], self.rope_sin[:S]
            q, k = ops.rope.apply_rotary(q, cos, sin), ops.rope.apply_rotary(k, cos, sin)
            if self.flash_decoding:
                self.kv.write_prefill_sharded(layer_idx, k, v, self.kv_rank)
            else:
                self.kv.write_prefill(layer_idx, k, v)
            o = ops.attention.flash_attention(q, k, v, causal=True)
        elif (q.is_cuda and q.dtype == torch.bfloat16 and self.kv.kv_quant is None and self.kv.seq_shards == 1
              and ops._ext.ext() is not None and hasattr(ops._ext.ext(), "decode_rope_kv")):
            # one launch: RoPE of q and k at each sequence's position + append of k / v to the cache (csrc/decode.cu)
            ops._ext.count_launch()
            q = ops._ext.ext().decode_rope_kv(q, k, v, positions.to(torch.long).contiguous(), self.rope_cos, self.rope_sin,
                                              self.kv.k[layer_idx], self.kv.v[layer_idx])
            kc, vc = self.kv.get(layer_idx, kv_len)
            o = _decode_attention(q, kc, vc, positions)
        else:
            # one new token per sequence at its own position
            cos = self.rope_cos[positions].unsqueeze(1)     # [B, 1, D/2] → per-batch tables
            sin = self.rope_sin[positions].unsqueeze(1)
            q, k = _rope_per_batch(q, cos, sin), _rope_per_batch(k, cos, sin)
            self.kv.write_decode(layer_idx, k, v, positions)
            kc, vc = self.kv.get(layer_idx, kv_len)
            o = _decode_attention(q, kc, vc, positions)
        o = o.transpose(0, 1).reshape(S, B, att.num_heads_local * D)
        return _row_linear_plus_residual(att.o_proj, o, x)

    def _tkg_block(self, layer_idx: int, layer):
        """Decode-time fused view of a MoE layer (shares its router / experts / norm; owns no parameters).  Kept outside the
        module tree so that state dicts and ``named_modules`` of the served model do not change."""
        blocks = self.__dict__.setdefault("_tkg_blocks", {})
        if layer_idx not in blocks:
            from ..modules.moe.moe_fused_tkg import MoEFusedTKG

            mlp = layer.mlp
            blocks[layer_idx] = MoEFusedTKG(mlp.router, mlp.expert_mlps, getattr(mlp, "shared_experts", None),
                                            layer.post_attention_layernorm, sequence_dimension=0,
                                            tensor_model_parallel_group=getattr(mlp, "tensor_parallel_group", None)).eval()
        return blocks[layer_idx]

    def _body(self, input_ids: torch.Tensor, positions: Optional[torch.Tensor], prefill: bool, kv_len: Optional[int], tree=None):
        core = self._core
        x = core.embed_tokens(input_ids).transpose(0, 1).contiguous()               # [S, B, H]
        for i, layer in enumerate(core.layers):
            x = self._attn_block(i, layer, x, positions, prefill, kv_len, tree)
            mlp = layer.mlp
            if hasattr(mlp, "gate_up_proj") and hasattr(mlp, "down_proj") and not prefill and x.shape[0] * x.shape[1] <= 8:
                # dense decode MLP: the residual add rides in the down-projection's GEMV(+all-reduce) epilogue
                hmid = ops.act.swiglu(mlp.gate_up_proj(layer.post_attention_layernorm(x)))
                x = _row_linear_plus_residual(mlp.down_proj, hmid, x)
                continue
            if not prefill and x.shape[0] * x.shape[1] <= 8 and hasattr(mlp, "expert_mlps") and hasattr(mlp, "router"):
                # MoE decode: norm → router → top-k → chosen experts → weighted sum as ONE launch when the kernel applies
                # (modules/moe/moe_fused_tkg.py, csrc/moe_tkg.cu); otherwise the MoE layer's own dispatch below
                fused = self._tkg_block(i, layer)
                if fused._can_use_kernel(x):
                    x = x + fused(x)[0]
                    continue
            y = mlp(layer.post_attention_layernorm(x))
            x = x + (y[0] if isinstance(y, tuple) else y)                           # MoE blocks also return router logits
        return core.norm(x)

    # ------------------------------------------------------------------ entry points
    @torch.no_grad()
    def context_encoding(self, input_ids: torch.Tensor, last_token_index: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``input_ids`` [B, S_bucket] (right-padded) → next token ids [B] (or vocab-parallel logits)."""
        h = self._body(input_ids, None, True, None)                              # [S, B, H]
        B = input_ids.shape[0]
        if last_token_index is None:
            last = h[-1]
        else:
            last = h[last_token_index, torch.arange(B, device=h.device)]
        logits = self.lm.lm_head(last.unsqueeze(0))[0].float()                   # [B, V/tp]
        return self.sampler.sample(logits) if self.on_device_sampling else logits

    @torch.no_grad()
    def token_generation(self, input_ids: torch.Tensor, positions: torch.Tensor, kv_len: Optional[int] = None) -> torch.Tensor:
        """``input_ids`` [B, 1], ``positions`` [B] → next token ids [B]."""
        h = self._body(input_ids, positions, False, kv_len)
        logits = self.lm.lm_head(h)[0].float()
        return self.sampler.sample(logits) if self.on_device_sampling else logits

    @torch.no_grad()
    def speculation_forward(self, input_ids: torch.Tensor, positions: torch.Tensor, kv_len: Optional[int] = None,
                            tree_mask: Optional[torch.Tensor] = None, tree_depth: Optional[torch.Tensor] = None,
                            return_hidden: bool = False):
        """Verify a window: ``input_ids`` [B, W] placed at ``positions[b] .. positions[b]+W-1`` → greedy next token after
        each of the W positions, ``[B, W]`` (reference ``examples/inference`` speculation model, ``speculation_length``).
        With ``tree_mask [W, W]`` (True = node row sees node column) and ``tree_depth [W]`` the window is a Medusa candidate
        tree instead of a chain.  ``return_hidden`` also returns the final hidden states ``[B, W, H]`` (Medusa heads input)."""
        tree = None if tree_mask is None else (tree_mask.bool(), tree_depth.long())
        h = self._body(input_ids, positions, False, kv_len, tree)                # [W, B, H]
        logits = self.lm.lm_head(h).float()                                       # [W, B, V/tp]
        W, B = logits.shape[:2]
        out = self.sampler.sample(logits.reshape(W * B, -1)).view(W, B).t() if self.on_device_sampling else logits.transpose(0, 1)
        return (out, h.transpose(0, 1)) if return_hidden else out

    @torch.no_grad()
    def generate(self, prompt_ids: torch.Tensor, max_new_tokens: int, prompt_lens: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, S = prompt_ids.shape
        lens = prompt_lens if prompt_lens is not None else torch.full((B,), S, device=prompt_ids.device, dtype=torch.long)
        tok = self.context_encoding(prompt_ids, lens - 1)
        out = [tok]
        pos = lens.clone()
        for _ in range(max_new_tokens - 1):
            tok = self.token_generation(tok.view(B, 1), pos)
            out.append(tok)
            pos = pos + 1
        return torch.stack(out, dim=1)


def _row_linear_plus_residual(lin, inp: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
    """``residual + RowParallelLinear(inp)``.  At decode (<= 8 rows, bf16, no bias) this is ONE kernel: the GEMV with the residual
    in its epilogue (tp = 1) or the fused GEMV + in-switch all-reduce + residual (tp > 1, csrc/nvls_coll.cu)."""
    M = inp.numel() // inp.shape[-1]
    w = lin.weight
    e = ops._ext.ext() if inp.is_cuda else None
    if (e is not None and M <= 8 and inp.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and getattr(lin, "bias", None) is None
            and not torch.is_grad_enabled() and getattr(lin, "reduce_output", True) and w.shape[1] % 8 == 0 and w.shape[0] % 8 == 0
            and not getattr(lin, "sequence_parallel_enabled", False)):
        x2 = inp.reshape(M, inp.shape[-1]).contiguous()
        r2 = residual.reshape(M, residual.shape[-1]).contiguous()
        tp = lin.tensor_model_parallel_size
        if tp == 1 and hasattr(e, "gemv"):
            ops._ext.count_launch()
            return e.gemv(x2, w, r2).view_as(residual)
        if tp > 1 and ops.tp_fused.get_backend() == "fused" and ops.nvls.gemv_all_reduce_eligible(x2, w):
            return ops.nvls.gemv_all_reduce(x2, w, lin.tensor_parallel_group, residual=r2).view_as(residual)
    return residual + lin(inp)


def _rope_per_batch(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x [B,1,H,D]; cos/sin [B,1,D/2] — tiny decode-time rotation (elementwise, graph-capturable)."""
    d2 = x.shape[-1] // 2
    xf = x.float()
    x1, x2 = xf[..., :d2], xf[..., d2:]
    c, s = cos.unsqueeze(2), sin.unsqueeze(2)
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1).to(x.dtype)


def _rope_window(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x [B,W,H,D]; cos/sin [B,W,D/2]."""
    d2 = x.shape[-1] // 2
    xf = x.float()
    x1, x2 = xf[..., :d2], xf[..., d2:]
    c, s = cos.unsqueeze(2), sin.unsqueeze(2)
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1).to(x.dtype)


def _window_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, positions: torch.Tensor,
                      tree_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [B,W,H,D] in cache slots p..p+W-1 vs cache k/v [B,L,Hkv,D].  Chain: token w sees slots ≤ p+w.  Tree (``tree_mask``
    [W,W]): node w sees every slot < p (the committed sequence) and the window slots of its ancestors."""
    B, W, H, D = q.shape
    Hkv, L = k.shape[2], k.shape[1]
    qt, kt, vt = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    if H != Hkv:
        kt, vt = kt.repeat_interleave(H // Hkv, 1), vt.repeat_interleave(H // Hkv, 1)
    cols = torch.arange(L, device=q.device)
    if tree_mask is None:
        lim = positions[:, None] + torch.arange(W, device=q.device)[None, :]                   # [B, W]
        mask = (cols[None, None, :] <= lim[:, :, None])[:, None]                               # [B,1,W,L]
    else:
        rel = cols[None, :] - positions[:, None]                                               # [B, L] slot index inside the window
        inside = (rel >= 0) & (rel < W)
        tm = tree_mask.to(q.device)[:, rel.clamp(0, W - 1)].permute(1, 0, 2)                   # [B, W, L]
        mask = ((rel < 0)[:, None, :] | (inside[:, None, :] & tm))[:, None]
    return torch.nn.functional.scaled_dot_product_attention(qt, kt, vt, attn_mask=mask).transpose(1, 2)


def _decode_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
    """q [B,1,H,D] vs cache k/v [B,L,Hkv,D]; keys beyond each sequence's position are masked."""
    B, _, H, D = q.shape
    Hkv, L = k.shape[2], k.shape[1]
    e = ops._ext.ext() if q.is_cuda else None
    if (e is not None and hasattr(e, "decode_attention") and D == 128 and q.dtype == torch.bfloat16 and k.dtype == torch.bfloat16
            and H // Hkv in (1, 2, 4, 8) and k.stride(3) == 1 and v.stride(3) == 1):
        # flash-decoding kernel (csrc/decode.cu): split over the cache length, GQA group shares each K/V read
        ops._ext.count_launch(2)
        return e.decode_attention(q.contiguous(), k, v, positions.to(torch.long).contiguous(), 1.0 / math.sqrt(D))
    qt = q.transpose(1, 2)                                   # [B,H,1,D]
    kt, vt = k.transpose(1, 2), v.transpose(1, 2)
    if H != Hkv:
        kt, vt = kt.repeat_interleave(H // Hkv, 1), vt.repeat_interleave(H // Hkv, 1)
    mask = (torch.arange(L, device=q.device)[None, :] <= positions[:, None])[:, None, None, :]
    o = torch.nn.functional.scaled_dot_product_attention(qt, kt, vt, attn_mask=mask)
    return o.transpose(1, 2)
