"""AdamW that keeps its moments (and the update math) in fp32 even for bf16 params — role of
reference ``utils/adamw_fp32_optim_params.py:31-155``.  All tensors of a param group are updated
by ONE fused multi-tensor kernel launch (``ops.optim.fused_adamw_``)."""
from __future__ import annotations

from typing import Callable, Iterable, Optional, Tuple

import torch
from torch.optim import Optimizer

from .. import ops


class AdamW_FP32OptimParams(Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-6, weight_decay: float = 0.0, correct_bias: bool = True, **_unused):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid betas: {betas}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias))
        self.grad_scale: Optional[torch.Tensor] = None  # device scalar folded into the update (clip coeff)
        self.supports_lowp_view = True  # fp32 params may carry `_lowp_view` (bf16 copy refreshed in-kernel)

    @torch.no_grad()
    def step(self, closure: Optional[Callable] = None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            ps_, gs, ms, vs, lowp, steps = [], [], [], [], [], []
            for p in group["params"]:
                g = p.grad if p.grad is not None else getattr(p, "main_grad", None)
                if g is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, dtype=torch.float32)
                    st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32)
                    if p.dtype != torch.float32:
                        st["master"] = p.detach().float().clone()
                st["step"] += 1
                steps.append(st["step"])
                if "master" in st:
                    ps_.append(st["master"])
                    lowp.append(p.data)
                else:
                    ps_.append(p.data)
                    lowp.append(getattr(p, "_lowp_view", None))
                gs.append(g)
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
            if not ps_:
                continue
            b1, b2 = group["betas"]
            has_low = [x is not None for x in lowp]
            # one fused multi-tensor launch per (low-precision copy?, step count): the bias correction is a per-parameter
            # quantity — parameters that joined later (un-frozen, or without a gradient in earlier steps) have their own count
            for want_low in (True, False):
                for step in sorted(set(steps)):
                    idx = [i for i, h in enumerate(has_low) if h == want_low and steps[i] == step]
                    if not idx:
                        continue
                    ops.optim.fused_adamw_(
                        [ps_[i] for i in idx], [gs[i] for i in idx], [ms[i] for i in idx], [vs[i] for i in idx],
                        group["lr"], b1, b2, group["eps"], group["weight_decay"],
                        step, self.grad_scale,
                        [lowp[i] for i in idx] if want_low else None,
                        hf_form=True, correct_bias=group["correct_bias"],
                    )
        return loss
