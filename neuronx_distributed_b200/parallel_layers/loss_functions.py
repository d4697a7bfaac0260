"""Vocab-parallel cross entropy and log-prob extraction.

Capability parity with reference ``parallel_layers/loss_functions.py`` (``parallel_cross_entropy``
:217, ``_ParallelCrossEntropy`` :10-129, ``DistributedLogprob`` / ``from_parallel_logits_to_logprobs``
:131-215).

B200 design: one pass over the local ``[T, V/tp]`` logits produces three per-row statistics
(local max, local sum-exp relative to that max, target logit if owned); the three reference
all-reduces (MAX, SUM, SUM) collapse into ONE all-gather of a ``[T, 3]`` fp32 tensor followed by
a log-sum-exp merge, and the softmax is *recomputed* in backward from the saved logits and the
row log-normaliser instead of being stored in fp32 (saves T·V/tp·4 bytes).  On CUDA the two
passes are the hand-written kernels ``ops.cross_entropy.ce_stats / ce_backward``.

Deviation (documented): label smoothing averages log-probs over the *global* vocabulary; the
reference averages over the local partition only (its ``vocab_size`` is the shard size), which
coincides with this for tp=1.
"""
from __future__ import annotations


import torch
import torch.distributed as dist

from . import comm
from . import parallel_state as ps
from .utils import EmbeddingUtility


def _row_stats(logits2d: torch.Tensor, target1d: torch.Tensor, vocab_start: int):
    """(local max, local sumexp wrt local max, owned target logit or 0, owned mask, sum of logits) per row, fp32."""
    from .. import ops

    return ops.cross_entropy.ce_stats(logits2d, target1d, vocab_start)


class _ParallelCrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vocab_parallel_logits, target, label_smoothing=0.0, group=None):
        from .. import ops

        group = group if group is not None else ps.get_tensor_model_parallel_group()
        n, r = dist.get_world_size(group), dist.get_rank(group)
        vp = vocab_parallel_logits.shape[-1]
        start, _ = EmbeddingUtility.range_from_per_partition_vocab_size(vp, r, n)
        logits2d = vocab_parallel_logits.reshape(-1, vp)
        tgt = target.reshape(-1)
        # stats: [T, 4] = (max, sumexp, target_logit, sum_logits)
        stats = ops.cross_entropy.ce_stats(logits2d, tgt, start)
        if n > 1:
            allstats = comm.all_gather(stats.unsqueeze(0), dim=0, group=group)  # [n, T, 4]
            gmax = allstats[..., 0].max(dim=0).values
            sumexp = (allstats[..., 1] * torch.exp(allstats[..., 0] - gmax)).sum(0)
            pred = allstats[..., 2].sum(0)
            sum_logits = allstats[..., 3].sum(0)
        else:
            gmax, sumexp, pred, sum_logits = stats[:, 0], stats[:, 1], stats[:, 2], stats[:, 3]
        lse = gmax + torch.log(sumexp)  # log normaliser per row
        loss = lse - pred
        vocab = vp * n
        if label_smoothing > 0:
            assert 1.0 > label_smoothing > 0.0
            smoothing = label_smoothing * vocab / (vocab - 1)
            mean_log_probs = sum_logits / vocab - lse
            loss = (1.0 - smoothing) * loss - smoothing * mean_log_probs
        ctx.label_smoothing, ctx.vocab, ctx.start = label_smoothing, vocab, start
        ctx.save_for_backward(logits2d, tgt, lse)
        ctx.shape = vocab_parallel_logits.shape
        return loss.view(target.shape)

    @staticmethod
    def backward(ctx, grad_output):
        from .. import ops

        logits2d, tgt, lse = ctx.saved_tensors
        if ctx.label_smoothing > 0:
            smoothing = ctx.label_smoothing * ctx.vocab / (ctx.vocab - 1)
        else:
            smoothing = 0.0
        g = ops.cross_entropy.ce_backward(
            logits2d, tgt, lse, grad_output.reshape(-1).float(), ctx.start, smoothing, ctx.vocab
        )
        return g.view(ctx.shape), None, None, None


def parallel_cross_entropy(vocab_parallel_logits, target, label_smoothing: float = 0.0, process_group=None):
    """Per-token cross-entropy for logits sharded along the vocab (last) dim across TP."""
    return _ParallelCrossEntropy.apply(vocab_parallel_logits, target, label_smoothing, process_group)


class DistributedLogprob(torch.autograd.Function):
    """log p(target) from vocab-parallel logits (DPO/ORPO-style objectives; reference :131-204)."""

    @staticmethod
    def forward(ctx, vocab_parallel_logits, target, inference_only=False, group=None):
        from .. import ops

        group = group if group is not None else ps.get_tensor_model_parallel_group()
        n, r = dist.get_world_size(group), dist.get_rank(group)
        vp = vocab_parallel_logits.shape[-1]
        start, _ = EmbeddingUtility.range_from_per_partition_vocab_size(vp, r, n)
        logits2d = vocab_parallel_logits.reshape(-1, vp)
        tgt = target.reshape(-1)
        stats = ops.cross_entropy.ce_stats(logits2d, tgt, start)
        if n > 1:
            allstats = comm.all_gather(stats.unsqueeze(0), dim=0, group=group)
            gmax = allstats[..., 0].max(dim=0).values
            sumexp = (allstats[..., 1] * torch.exp(allstats[..., 0] - gmax)).sum(0)
            pred = allstats[..., 2].sum(0)
        else:
            gmax, sumexp, pred = stats[:, 0], stats[:, 1], stats[:, 2]
        lse = gmax + torch.log(sumexp)
        if not inference_only:
            ctx.save_for_backward(logits2d, tgt, lse)
            ctx.start, ctx.shape, ctx.vocab = start, vocab_parallel_logits.shape, vp * n
        return (pred - lse).view(target.shape)

    @staticmethod
    def backward(ctx, grad_output):
        from .. import ops

        logits2d, tgt, lse = ctx.saved_tensors
        # d(logp)/dlogits = onehot - softmax = -(softmax - onehot)
        g = ops.cross_entropy.ce_backward(
            logits2d, tgt, lse, (-grad_output).reshape(-1).float(), ctx.start, 0.0, ctx.vocab
        )
        return g.view(ctx.shape), None, None, None


def from_parallel_logits_to_logprobs(vocab_parallel_logits, target, inference_only: bool = False, process_group=None):
    """Log-probabilities of ``target[:, 1:]`` under ``logits[:, :-1]`` (next-token shift as in the
    reference :206-215)."""
    target = target.roll(shifts=-1, dims=-1)
    probs = DistributedLogprob.apply(vocab_parallel_logits, target, inference_only, process_group)
    return probs[:, :-1].contiguous()
