"""HF-style ``generate()`` for the serving models (role of the reference's ``examples/inference/modules/hf_adapter.py``
``HuggingFaceGenerationAdapter``): the argument names and return layout of ``transformers``' ``generate`` — ``input_ids`` +
``attention_mask`` (left- or right-padded batches), ``max_new_tokens`` / ``max_length``, ``do_sample`` / ``top_k`` / ``top_p`` /
``temperature``, ``eos_token_id`` / ``pad_token_id``, ``logits_processor`` / ``stopping_criteria`` callables — on top of the
``context_encoding`` / ``token_generation`` programs (bucketed CUDA graphs when wrapped by ``ModelBuilder``).

Differences from running HF's own loop: the KV cache is the model's persistent on-device cache (no ``past_key_values``
objects), prompts are re-packed right-padded for the prefill program, sampling runs on the vocab-parallel logits
(distributed top-k) and the sampled ids are broadcast inside the TP group so every rank continues with the same token."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Union

import torch
import torch.distributed as dist

from ..parallel_layers import parallel_state as ps
from ..utils.sampling import Sampler


class HuggingFaceGenerationAdapter:
    def __init__(self, model, eos_token_id: Optional[Union[int, Sequence[int]]] = None, pad_token_id: Optional[int] = None):
        """``model``: ``LlamaForInference``-like object (``context_encoding``, ``token_generation``, ``kv``, ``sampler``,
        ``on_device_sampling``, ``batch_size``, ``max_seq_len``) or a routed ``NxDModel`` built from one (pass the module as
        ``model`` and the router through :meth:`use_programs`)."""
        self.model = model
        self.eos_token_id, self.pad_token_id = eos_token_id, pad_token_id
        self._ctx: Callable = model.context_encoding
        self._tkg: Callable = model.token_generation

    def use_programs(self, context_encoding: Callable, token_generation: Callable) -> "HuggingFaceGenerationAdapter":
        """Route the two phases through captured programs (``nxd_model`` callables) instead of the eager module."""
        self._ctx, self._tkg = context_encoding, token_generation
        return self

    # ---- helpers ------------------------------------------------------------------------------------------------------
    @staticmethod
    def _pack_right(input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], pad: int):
        """Rows with their real tokens moved to the front (stable), and the real lengths."""
        if attention_mask is None:
            B, S = input_ids.shape
            return input_ids, torch.full((B,), S, dtype=torch.long, device=input_ids.device)
        mask = attention_mask.bool()
        lens = mask.sum(1)
        order = torch.argsort((~mask).to(torch.int8), dim=1, stable=True)            # real tokens first, original order kept
        packed = input_ids.gather(1, order)
        keep = torch.arange(input_ids.shape[1], device=input_ids.device)[None, :] < lens[:, None]
        return torch.where(keep, packed, torch.full_like(packed, pad)), lens

    def _sync(self, tok: torch.Tensor) -> torch.Tensor:
        if dist.is_initialized() and ps.model_parallel_is_initialized() and ps.get_tensor_model_parallel_size() > 1:
            dist.broadcast(tok, src=ps.get_tensor_model_parallel_src_rank(), group=ps.get_tensor_model_parallel_group())
        return tok

    def _pick(self, out: torch.Tensor, sampler: Optional[Sampler], processors, history: torch.Tensor, generator) -> torch.Tensor:
        """``out``: token ids [B] (on-device sampling) or vocab-parallel logits [B, V/tp]."""
        if out.dim() == 1:
            return out
        if processors:                                                               # HF processors want the full vocabulary
            from ..parallel_layers.mappings import gather_from_tensor_model_parallel_region

            logits = gather_from_tensor_model_parallel_region(out) if sampler.vocab_parallel else out
            for proc in processors:
                logits = proc(history, logits)
            full = Sampler(top_k=sampler.top_k, top_p=sampler.top_p, temperature=sampler.temperature, do_sample=sampler.do_sample,
                           vocab_parallel=False)
            return self._sync(full.sample(logits, generator=generator))
        return self._sync(sampler.sample(out, generator=generator))

    # ---- API ----------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, max_new_tokens: Optional[int] = None,
                 max_length: Optional[int] = None, do_sample: bool = False, top_k: int = 50, top_p: float = 1.0, temperature: float = 1.0,
                 eos_token_id: Optional[Union[int, Sequence[int]]] = None, pad_token_id: Optional[int] = None,
                 logits_processor: Optional[List[Callable]] = None, stopping_criteria: Optional[List[Callable]] = None,
                 generator: Optional[torch.Generator] = None, **unused) -> torch.Tensor:
        """Returns ``[B, S + n]``: the prompt rows as given followed by the generated tokens (``pad_token_id`` after a row's
        EOS), like ``transformers``."""
        m = self.model
        B, S = input_ids.shape
        assert B == m.batch_size, f"batch {B} != compiled batch size {m.batch_size}"
        eos = eos_token_id if eos_token_id is not None else self.eos_token_id
        eos_ids = [] if eos is None else ([eos] if isinstance(eos, int) else list(eos))
        pad = pad_token_id if pad_token_id is not None else (self.pad_token_id if self.pad_token_id is not None else (eos_ids[0] if eos_ids else 0))
        packed, lens = self._pack_right(input_ids, attention_mask, pad)
        room = m.max_seq_len - int(lens.max())
        n_new = max_new_tokens if max_new_tokens is not None else ((max_length - S) if max_length is not None else room)
        n_new = max(0, min(n_new, room))
        if n_new == 0:
            return input_ids
        custom = do_sample or bool(logits_processor)
        saved = (m.sampler, m.on_device_sampling)
        sampler = Sampler(top_k=top_k if do_sample else 1, top_p=top_p, temperature=temperature, do_sample=do_sample,
                          vocab_parallel=getattr(saved[0], "vocab_parallel", True)) if custom else saved[0]
        if custom:                                 # get logits back and sample here (per-call parameters, synchronised over TP)
            m.on_device_sampling = False
        try:
            if hasattr(m, "kv"):
                m.kv.reset()
            history = input_ids
            tok = self._pick(self._ctx(packed, lens - 1), sampler, logits_processor, history, generator)
            done = torch.zeros(B, dtype=torch.bool, device=input_ids.device)
            eos_t = torch.tensor(eos_ids, device=input_ids.device) if eos_ids else None
            out, pos = [], lens.clone()
            for i in range(n_new):
                tok = torch.where(done, torch.full_like(tok, pad), tok)
                out.append(tok)
                history = torch.cat([history, tok[:, None]], 1)
                if eos_t is not None:
                    done = done | torch.isin(tok, eos_t)
                if i == n_new - 1 or (eos_t is not None and bool(done.all())) or \
                        (stopping_criteria and any(bool(torch.as_tensor(c(history, None)).all()) for c in stopping_criteria)):
                    break
                tok = self._pick(self._tkg(tok.view(B, 1), pos), sampler, logits_processor, history, generator)
                pos = pos + 1
            return torch.cat([input_ids, torch.stack(out, 1)], 1)
        finally:
            m.sampler, m.on_device_sampling = saved

    __call__ = generate
