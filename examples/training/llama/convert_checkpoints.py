#!/usr/bin/env python
"""Llama checkpoint conversion CLI — counterpart of the reference's ``examples/training/llama/convert_checkpoints.py``.

HF full state ↔ the (TP, PP)-sharded checkpoint layout ``trainer.load_checkpoint`` reads.  The built-in Llama uses the HF key
names (``model.layers.N…``); pass ``--qkv_linear --fuse_qkv --fuse_gate_up`` for its fused QKV / gate-up parameters:

  python examples/training/llama/convert_checkpoints.py --convert_from_full_state --input_dir hf_llama/ --config hf_llama/config.json \\
      --output_dir ckpt/ --tp_size 8 --pp_size 1 --kv_size_multiplier 1 --qkv_linear --fuse_qkv --fuse_gate_up --save_xser
  python examples/training/llama/convert_checkpoints.py --convert_to_full_state --input_dir ckpt/ --output_dir merged/ --tp_size 8 ...

(To load an HF checkpoint directly into a live sharded model without files, use ``models.hf_compat.load_hf_checkpoint``.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))

from neuronx_distributed_b200.scripts.checkpoint_converter import CheckpointConverterBase  # noqa: E402


class CheckpointConverterLlama(CheckpointConverterBase):
    pass


if __name__ == "__main__":
    converter = CheckpointConverterLlama()
    args, _ = converter.get_arg_parser().parse_known_args()
    converter.run(args)
