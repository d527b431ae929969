"""ORACLE (test infrastructure, never shipped) — the optimizer end of the reference train step:

  clip_grad_norm_(net.parameters(), 1.0)        rsuper_train/train_ddp.py:352   (torch/nn/utils/clip_grad.py semantics)
  AdamW(lr, betas, eps=1e-5, weight_decay)      rsuper_train/training/utils.py:46-51 (torch/optim/adamw.py, single-tensor path)
  update_ema_variables(model, ema, alpha, step) rsuper_train/training/utils.py:154-158

restated as explicit fp32 element-wise arithmetic (what csrc/train_glue.cu computes per element).  Pinned against outputs
of the REAL reference functions `get_optimizer` / `update_ema_variables` driven in train_epoch's order
(tests/golden/make_golden_glue.py -> tests/golden/reference_train_glue.npz).
"""
from __future__ import annotations

import math
from typing import List, Sequence

import torch


def glue_inputs(step: int = -1, device="cpu") -> List[torch.Tensor]:
    """Deterministic tensors from formulas (nothing stored): step = -1 -> initial parameters, step >= 0 -> that step's
    gradients (large for steps 0-1 so that clipping is active, small afterwards)."""
    shapes = [(5, 3, 3, 3, 3), (4097,), (2, 7), (1,)]
    out = []
    for i, s in enumerate(shapes):
        n = 1
        for d in s:
            n *= d
        k = torch.arange(n, dtype=torch.float64)
        v = torch.sin(k * (0.37 + 0.11 * i) + 1.3 * (step + 2)) * torch.cos(k * 0.013 * (i + 1) + 0.7 * step)
        scale = 0.5 if step < 0 else (2.0 if step < 2 else 2e-3)
        out.append((scale * v).float().reshape(s).to(device))
    return out


def clip_adamw_ema_step(params: Sequence[torch.Tensor], grads: Sequence[torch.Tensor], exp_avg: Sequence[torch.Tensor],
                        exp_avg_sq: Sequence[torch.Tensor], ema: Sequence[torch.Tensor], step: int, global_step: int, lr: float,
                        betas=(0.9, 0.999), eps: float = 1e-5, weight_decay: float = 0.05, max_norm: float = 1.0,
                        ema_alpha: float = 0.99) -> torch.Tensor:
    """In-place update of every list (fp32 tensors); `step` counts from 1.  Returns the total gradient norm before clipping."""
    f = torch.float32
    norms = torch.stack([g.to(f).norm(2) for g in grads])
    total = norms.norm(2)                                             # clip_grad_norm_: norm of the per-tensor norms
    clip = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    b1, b2 = betas
    step_size = lr / (1.0 - b1 ** step)
    bc2_sqrt = math.sqrt(1.0 - b2 ** step)
    alpha = min(1.0 - 1.0 / (global_step + 1), ema_alpha)             # training/utils.py:156
    for p, g, m, v, e in zip(params, grads, exp_avg, exp_avg_sq, ema):
        g.mul_(clip)
        p.mul_(1.0 - lr * weight_decay)
        m.add_((g - m) * (1.0 - b1))
        v.mul_(b2).add_(g * g * (1.0 - b2))
        p.sub_(step_size * (m / (v.sqrt() / bc2_sqrt + eps)))
        e.mul_(alpha).add_(p * (1.0 - alpha))
    return total
