#!/usr/bin/env python
"""DBRX checkpoint conversion CLI — counterpart of the reference's ``examples/training/dbrx/convert_checkpoints.py``.
Same converter as Mixtral (``models.hf_compat`` recognises the ``transformer.blocks.N.…`` spelling on the way in); the
sharded → full direction writes DBRX names again."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mixtral"))

from convert_checkpoints import CheckpointConverterMixtral  # noqa: E402


class CheckpointConverterDbrx(CheckpointConverterMixtral):
    hf_style = "dbrx"


if __name__ == "__main__":
    converter = CheckpointConverterDbrx()
    args, _ = converter.get_arg_parser().parse_known_args()
    converter.run(args)
