#!/usr/bin/env python
"""Llama pre-training with TP x PP (1F1B / interleaved) — counterpart of the reference's
``examples/training/llama/tp_pp_llama_hf_pretrain/run_llama_nxd.py``.  Same driver, flags, checkpointing, schedule and metrics
as ``tp_zero1_llama_pretrain.py``; this front end only changes the defaults to a pipelined run.

  torchrun --nproc-per-node 8 examples/training/llama/tp_pp_llama_pretrain.py --model 13b --tensor_parallel_size 4 \
      --pipeline_parallel_size 2 --num_microbatches 8 --seq_len 4096 --checkpoint_freq 100 --checkpoint_dir ckpt
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from tp_zero1_llama_pretrain import main  # noqa: E402

_DEFAULTS = {"--model": "tiny", "--tensor_parallel_size": "2", "--pipeline_parallel_size": "2", "--num_microbatches": "4",
             "--seq_len": "512", "--max_steps": "10", "--grad_accum_usteps": "1", "--warmup_steps": "2"}


def _with_defaults(argv):
    given = {a.split("=", 1)[0] for a in argv if a.startswith("--")}
    extra = []
    for k, v in _DEFAULTS.items():
        if k not in given:
            extra += [k, v]
    return extra + list(argv)


if __name__ == "__main__":
    main(_with_defaults(sys.argv[1:]))
