#!/usr/bin/env python
"""Numerical bisection with tensor replacement — counterpart of the reference's
``examples/inference/tensor_replacement/tensor_replacement_padding_example.py``.

A "golden" (fp32) run is captured layer by layer; the low-precision model is then prepared ONCE for replacement
(``modify_model_for_tensor_replacement``) and re-run with the golden output injected after layer 0, 1, 2, … through boolean
masks.  The layer after which the final error collapses is the one that loses the precision.  Because the injection is
data (tensors + masks appended to the call), all variants run the same program.

  python examples/inference/tensor_replacement_example.py
"""
import os
import sys

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from neuronx_distributed_b200.utils.tensor_capture import (disable_tensor_capture, enable_tensor_capture,  # noqa: E402
                                                           get_captured_tensors_dict)
from neuronx_distributed_b200.utils.tensor_replacement import RuntimeRegister, modify_model_for_tensor_replacement  # noqa: E402


class Toy(nn.Module):
    def __init__(self, depth=4, width=64):
        super().__init__()
        self.layers = nn.ModuleList([nn.Sequential(nn.Linear(width, width), nn.GELU()) for _ in range(depth)])
        self.head = nn.Linear(width, 8)

    def forward(self, x):
        for layer in self.layers:
            x = x + layer(x)
        return self.head(x)


def main():
    torch.manual_seed(0)
    golden = Toy().eval()
    lowp = Toy().eval()
    lowp.load_state_dict(golden.state_dict())
    with torch.no_grad():                                   # the "bug": layer 2 of the low-precision model is badly quantised
        w = lowp.layers[2][0].weight
        w.copy_((w * 8).round() / 8)
    x = torch.randn(16, 64)
    names = [f"layers.{i}" for i in range(len(golden.layers))]
    enable_tensor_capture(golden, names)
    with torch.no_grad():
        ref = golden(x)
    cap = get_captured_tensors_dict()
    gold = [cap[f"{n}.outputs"] for n in names]
    disable_tensor_capture(golden)

    RuntimeRegister.module_superset = names
    lowp, hooks = modify_model_for_tensor_replacement(lowp)
    off, on = torch.zeros((), dtype=torch.bool), torch.ones((), dtype=torch.bool)
    print("replace output of   final max error vs golden")
    with torch.no_grad():
        base_err = float((lowp(x, *gold, *([off] * len(names))) - ref).abs().max())
        print(f"  (nothing)         {base_err:.3e}")
        for i, n in enumerate(names):
            masks = [on if j == i else off for j in range(len(names))]
            err = float((lowp(x, *gold, *masks) - ref).abs().max())
            print(f"  {n:16s}  {err:.3e}{'   <-- error disappears: the fault is at or before this layer' if err < 1e-6 <= base_err else ''}")
    for h in hooks.values():
        h.remove()
    RuntimeRegister.module_superset = []


if __name__ == "__main__":
    main()
