#!/usr/bin/env python
"""DBRX inference (fine-grained MoE: 16 experts, top-4, clipped QKV) — counterpart of the reference's
``examples/inference/run_dbrx.py``.  Thin front-end of ``run_mixtral.py --family dbrx``.

  torchrun --nproc-per-node 2 examples/inference/run_dbrx.py --tp_degree 2
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

if __name__ == "__main__":
    import run_mixtral

    if "--family" not in sys.argv:
        sys.argv += ["--family", "dbrx"]
    run_mixtral.main()
