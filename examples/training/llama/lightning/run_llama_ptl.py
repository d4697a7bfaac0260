#!/usr/bin/env python
"""Llama pre-training through the Lightning integration — counterpart of the reference's
``examples/training/llama/lightning/run_llama_nxd_ptl.py`` (+ ``module_llama.py`` / ``data_module.py``).

``NeuronLTModule`` owns the model / optimizer / scheduler factories and the manual-optimisation step (works with and without
pipeline parallelism), ``NxDStrategy`` brings up the process groups and tells the sampler about the DP topology,
``NeuronCheckpointIO`` writes the same sharded checkpoint layout as the plain loops, ``NeuronTensorBoardLogger`` logs from
the rank that owns the loss.  With ``lightning`` installed the same objects plug into ``lightning.pytorch.Trainer``; without
it (this image) the package's own fit loop drives them.

  torchrun --nproc-per-node 8 examples/training/llama/lightning/run_llama_ptl.py --tensor_parallel_size 8 --model 7b
  NXD_CPU_MODE=1 torchrun --nproc-per-node 2 examples/training/llama/lightning/run_llama_ptl.py --tensor_parallel_size 2 --model tiny --max_steps 6
"""
import argparse
import math
import os
import sys

import torch
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples", "training"))

import neuronx_distributed_b200 as nxd  # noqa: E402
from neuronx_distributed_b200.lightning import (NeuronCheckpointIO, NeuronHooksCallback, NeuronLTModule,  # noqa: E402
                                                NeuronTensorBoardLogger, NeuronTQDMProgressBar, NxDStrategy)
from neuronx_distributed_b200.lightning._compat import HAVE_LIGHTNING  # noqa: E402
from neuronx_distributed_b200.models.llama import (LlamaConfig, LlamaForCausalLM, llama2_7b_config, llama2_13b_config,  # noqa: E402
                                                  llama2_70b_config)
from neuronx_distributed_b200.utils import get_device  # noqa: E402


VOCAB_TINY = 2048


class SyntheticTokens(Dataset):
    """Deterministic token stream (there is no dataset access in the sandbox); replace with a tokenised corpus."""

    def __init__(self, vocab: int, seq_len: int, n: int = 1 << 14, seed: int = 1234):
        g = torch.Generator().manual_seed(seed)
        self.data = torch.randint(0, vocab, (n, seq_len), generator=g)

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, i):
        return {"input_ids": self.data[i], "labels": self.data[i]}


class LlamaDataModule:
    """``setup`` runs after the strategy initialised the groups, so the sampler can ask it for the DP rank / size."""

    def __init__(self, vocab: int, seq_len: int, micro_batch: int):
        self.vocab, self.seq_len, self.micro_batch, self.trainer = vocab, seq_len, micro_batch, None

    def setup(self, stage=None):
        self.ds = SyntheticTokens(self.vocab, self.seq_len)

    def train_dataloader(self):
        kw = self.trainer.strategy.distributed_sampler_kwargs
        return DataLoader(self.ds, batch_size=self.micro_batch, sampler=DistributedSampler(self.ds, shuffle=True, seed=7, **kw),
                          drop_last=True, pin_memory=torch.cuda.is_available())


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--model", default="tiny")
    p.add_argument("--tensor_parallel_size", type=int, default=1)
    p.add_argument("--pipeline_parallel_size", type=int, default=1)
    p.add_argument("--num_microbatches", type=int, default=1)
    p.add_argument("--seq_len", type=int, default=128)
    p.add_argument("--micro_batch", type=int, default=2)
    p.add_argument("--grad_accum_usteps", type=int, default=2)
    p.add_argument("--max_steps", type=int, default=10)
    p.add_argument("--lr", type=float, default=3e-4)
    p.add_argument("--warmup_steps", type=int, default=2)
    p.add_argument("--use_zero_1", type=int, default=1)
    p.add_argument("--sequence_parallel", type=int, default=0)
    p.add_argument("--checkpoint_dir", default=None)
    p.add_argument("--checkpoint_freq", type=int, default=0)
    p.add_argument("--resume_from", default=None)
    p.add_argument("--log_dir", default="/tmp/nxd_ptl_logs")
    # flags of the reference script (run_llama_nxd_ptl.py): checkpoint switches and the activation / gradient dump hooks
    p.add_argument("--save_checkpoint", action="store_true", help="same as --checkpoint_freq > 0 (every --checkpoint_freq or 1 steps)")
    p.add_argument("--load_step", type=int, default=-1, help="resume from <checkpoint_dir>/step_<n> (-1: --resume_from decides)")
    p.add_argument("--load_epoch", type=int, default=-1, help="accepted: checkpoints are tagged by optimizer step")
    p.add_argument("--log_rank0", action="store_true", help="log from global rank 0 instead of the last pipeline stage")
    p.add_argument("--hooks", action="store_true", help="record (input, output) of --target_layers during training")
    p.add_argument("--target_layers", default="", help="comma-separated module names (see --master_print_model_layers)")
    p.add_argument("--hooks_interval", type=int, default=1)
    p.add_argument("--hooks_dump_base_directory", default="./hooks_outputs/local/hooks_dumps")
    p.add_argument("--enable_activation_dumps", action="store_true")
    p.add_argument("--enable_grad_dumps", action="store_true")
    p.add_argument("--dump_only_norms", action="store_true")
    p.add_argument("--dump_only_master_rank", action="store_true")
    p.add_argument("--master_print_model_layers", action="store_true")
    p.add_argument("--enable_tb_logging_master_rank_activation_norms", action="store_true",
                   help="accepted: dumped norms are files under --hooks_dump_base_directory")
    p.add_argument("--enable_tb_logging_master_rank_grad_norms", action="store_true", help="accepted, as above")
    a = p.parse_args()
    if a.save_checkpoint and a.checkpoint_freq <= 0:
        a.checkpoint_freq = 1
    if a.load_step >= 0 and a.checkpoint_dir and not a.resume_from:
        a.resume_from = os.path.join(a.checkpoint_dir, f"step_{a.load_step}")

    def warmup_cosine(step):
        if step < a.warmup_steps:
            return (step + 1) / max(1, a.warmup_steps)
        frac = min(1.0, (step - a.warmup_steps) / max(1, a.max_steps - a.warmup_steps))
        return 0.1 + 0.9 * 0.5 * (1 + math.cos(math.pi * frac))

    nxd_config = nxd.neuronx_distributed_config(
        tensor_parallel_size=a.tensor_parallel_size, pipeline_parallel_size=a.pipeline_parallel_size,
        pipeline_config={"num_microbatches": a.num_microbatches, "input_names": ["input_ids", "labels"], "output_loss_value_spec": True}
        if a.pipeline_parallel_size > 1 else None,
        optimizer_config={"zero_one_enabled": bool(a.use_zero_1), "grad_clipping": True, "max_grad_norm": 1.0},
        sequence_parallel=bool(a.sequence_parallel))

    def model_fn():                                  # runs inside module.setup(), i.e. after the strategy created the groups
        dev = get_device()
        kw = dict(sequence_parallel_enabled=bool(a.sequence_parallel), dtype=torch.bfloat16 if dev.type == "cuda" else torch.float32,
                  device=dev, max_position_embeddings=a.seq_len)
        mcfg = {"7b": llama2_7b_config, "13b": llama2_13b_config, "70b": llama2_70b_config}.get(a.model, lambda **k: LlamaConfig(
            vocab_size=VOCAB_TINY, hidden_size=256, intermediate_size=704, num_hidden_layers=2, num_attention_heads=8, **k))(**kw)
        torch.manual_seed(0)
        return LlamaForCausalLM(mcfg)

    module = NeuronLTModule(nxd_config, model_fn, torch.optim.AdamW, scheduler_cls=torch.optim.lr_scheduler.LambdaLR,
                            opt_kwargs={"lr": a.lr, "betas": (0.9, 0.95), "weight_decay": 0.1},
                            scheduler_args=(warmup_cosine,),
                            grad_accum_steps=a.grad_accum_usteps, train_batch_size=a.micro_batch, log_rank0=a.log_rank0)
    dm = LlamaDataModule(VOCAB_TINY if a.model == "tiny" else 32000, a.seq_len, a.micro_batch)
    strategy = NxDStrategy(nxd_config=nxd_config)
    callbacks = [NeuronTQDMProgressBar()]
    if a.hooks:
        callbacks.append(NeuronHooksCallback(a))                        # the argument namespace carries the hook settings
    kwargs = dict(strategy=strategy, callbacks=callbacks, logger=NeuronTensorBoardLogger(a.log_dir, "llama_ptl", log_rank0=a.log_rank0),
                  max_steps=a.max_steps, log_every_n_steps=1, plugins=[NeuronCheckpointIO(async_save=False, num_kept_ckpts=2)])
    if HAVE_LIGHTNING:                                                   # pragma: no cover - not in the offline image
        import lightning.pytorch as pl

        trainer = pl.Trainer(enable_checkpointing=False, num_sanity_val_steps=0, **kwargs)
    else:
        from neuronx_distributed_b200.lightning import Trainer

        trainer = Trainer(default_root_dir=a.checkpoint_dir, every_n_train_steps=a.checkpoint_freq, **kwargs)
    trainer.fit(module, datamodule=dm, ckpt_path=a.resume_from)
    if strategy.is_global_zero:
        print(f"finished {trainer.global_step} optimizer steps; last logged: "
              + ", ".join(f"{k}={float(v):.4g}" for k, v in getattr(trainer, "callback_metrics", {}).items()))
    strategy.teardown()


if __name__ == "__main__":
    main()
