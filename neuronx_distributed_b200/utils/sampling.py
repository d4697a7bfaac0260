"""On-device token sampling (reference ``utils/sampling.py:6-77``): greedy, top-k / top-p / temperature
multinomial, all on the GPU with no host sync; with vocab-parallel logits the top-k candidates come from
:func:`operators.topk` (one all-gather of k·tp pairs)."""
from __future__ import annotations

from typing import Optional

import torch

from ..operators import argmax as dist_argmax
from ..operators import topk as dist_topk


class Sampler:
    def __init__(self, top_k=1, top_p: float = 1.0, temperature: float = 1.0, do_sample: bool = False,
                 dynamic: bool = False, deterministic: bool = False, on_device: bool = True, vocab_parallel: bool = False):
        """Explicit settings, or — the reference's constructor (sampling.py:12-22) — one ``neuron_config`` object carrying
        ``on_device_sampling`` and ``hf_config.{do_sample, num_beams, top_k[, top_p, temperature]}``; as there, anything but
        single-beam sampling is refused and ``sample`` is the inverse-CDF multinomial (greedy for ``top_k == 1``)."""
        self._reference_style = False
        if not isinstance(top_k, int):
            nc = top_k
            hf = getattr(nc, "hf_config", nc)
            if not (getattr(hf, "do_sample", False) and getattr(hf, "num_beams", 1) == 1):
                raise Exception("Selected sampling method is not supported.")
            on_device = bool(getattr(nc, "on_device_sampling", True))
            self.is_medusa = bool(getattr(nc, "is_medusa", False))
            top_k, do_sample = int(hf.top_k), True
            top_p, temperature = float(getattr(hf, "top_p", 1.0) or 1.0), float(getattr(hf, "temperature", 1.0) or 1.0)
            self._reference_style = True
        self.on_device_sampling = on_device
        self.top_k, self.top_p, self.temperature = top_k, top_p, temperature
        self.do_sample, self.dynamic, self.deterministic = do_sample, dynamic, deterministic
        self.vocab_parallel = vocab_parallel

    def sample(self, token_logits: torch.Tensor, rank_id: Optional[torch.Tensor] = None,
               generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """``logits`` [B, V] (or [B, V/tp] when ``vocab_parallel``) → token ids [B]."""
        logits = token_logits
        if self._reference_style and not self.vocab_parallel and self.top_p >= 1.0:
            return self.multinomial(logits, generator)
        if self.top_k == 1 or not self.do_sample:
            if self.vocab_parallel:
                return dist_argmax(logits, dim=-1, rank_id=rank_id)
            return torch.argmax(logits, dim=-1)
        k = self.top_k if self.top_k > 0 else logits.shape[-1]
        if self.vocab_parallel:
            vals, idx = dist_topk(logits, k, dim=-1, rank_id=rank_id)
        else:
            vals, idx = torch.topk(logits, min(k, logits.shape[-1]), dim=-1)
        vals = vals.float() / max(self.temperature, 1e-6)
        probs = torch.softmax(vals, dim=-1)
        if self.top_p < 1.0:
            cum = probs.cumsum(-1)
            keep = (cum - probs) < self.top_p          # always keeps the first candidate
            probs = probs * keep
            probs = probs / probs.sum(-1, keepdim=True)
        if self.deterministic:
            choice = probs.argmax(-1, keepdim=True)
        else:
            choice = torch.multinomial(probs, 1, generator=generator)
            if self.vocab_parallel:
                # every TP rank holds the same candidates but its own RNG stream: continue with rank 0's draw everywhere
                from ..parallel_layers import parallel_state as ps

                if ps.model_parallel_is_initialized() and ps.get_tensor_model_parallel_size() > 1:
                    torch.distributed.broadcast(choice, src=ps.get_tensor_model_parallel_src_rank(),
                                                group=ps.get_tensor_model_parallel_group())
        return idx.gather(-1, choice).squeeze(-1)


def _multinomial(self, token_logits: torch.Tensor, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Top-k multinomial sampling by inverse CDF (reference sampling.py:27-77): softmax over the k best logits, one uniform
    draw per row, the sample is the number of CDF entries below the draw.  Branch-free and free of host synchronisation
    (``torch.multinomial`` validates its input on the host), hence CUDA-graph capturable."""
    if self.top_k == 1:
        return torch.argmax(token_logits, dim=1)
    vals, idx = torch.topk(token_logits, min(self.top_k, token_logits.shape[1]), dim=1)
    cdf = torch.softmax(vals.float() / max(self.temperature, 1e-6), dim=1).cumsum(dim=1)
    u = torch.rand((cdf.shape[0], 1), device=token_logits.device, generator=generator)
    counts = ((cdf - u) < 0).sum(dim=1).clamp(max=idx.shape[1] - 1)
    return idx.gather(1, counts.unsqueeze(1)).flatten()


Sampler.multinomial = _multinomial


def create_sampler(neuron_config=None, **kw) -> Sampler:
    if neuron_config is not None:
        kw = {**{k: getattr(neuron_config, k) for k in ("top_k", "top_p", "temperature", "do_sample") if hasattr(neuron_config, k)}, **kw}
    return Sampler(**kw)
