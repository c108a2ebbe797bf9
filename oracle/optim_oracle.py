"""TEST INFRASTRUCTURE ONLY -- not part of the product path (imported by tests/, __graft_entry__.smoke() and bench.py's
CPU leg only).

CPU restatement, in elementwise fp32 torch arithmetic, of the optimizer-side sweep the reference trainer runs after
backward():

    grad norm + clip        training/trainer.py:1044  (accelerate -> torch.nn.utils.clip_grad_norm_)
    AdamW step              training/trainer.py:461-473,1062  (torch.optim.AdamW, decoupled decay)
    EMA / feedback lerps    training/ema.py:284-313   (torch._foreach_lerp_ per EMA, in config order)
    normalize_weights       training/trainer.py:1107-1108 -> modules/mp_tools.py:375-378 (normalize over dims 1..)

Parity is PINNED: tests/golden/optim_small.pt was produced by the unmodified reference classes (MPConv,
DualDiffusionModule.normalize_weights, EMA_Manager.update) + the torch optimizer they call
(tests/golden/make_golden_optim.py); tests/test_oracle_optim.py checks this restatement against it.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

Tensor = torch.Tensor


def lerp(a: Tensor, b: Tensor, w: float) -> Tensor:
    """torch.lerp's formula (ATen/native/Lerp.h): a + w (b - a) for w < 0.5, else b - (b - a)(1 - w)."""
    return torch.lerp(a, b, w)


def clip_coef(grads: Sequence[Tensor], max_norm: float):
    """torch.nn.utils.clip_grad_norm_ (norm_type 2): total norm and the clamped coefficient max_norm / (norm + 1e-6)."""
    norm = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g.float()) for g in grads]))
    coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
    return norm, coef


def normalize_rows(w: Tensor, eps: float = 1e-4) -> Tensor:
    """mp_tools.py:42-49 with dim=None: per output row, w / (eps + ||row|| / sqrt(fan_in))."""
    flat = w.reshape(w.shape[0], -1).float()
    n = torch.linalg.vector_norm(flat, dim=1, keepdim=True)
    n = eps + n * math.sqrt(1.0 / flat.shape[1])
    return (flat / n).reshape(w.shape).to(w.dtype)


def train_update(params: Dict[str, Tensor], grads: Dict[str, Tensor], exp_avg: Dict[str, Tensor],
                 exp_avg_sq: Dict[str, Tensor], step: int, *, lr: float, betas, eps: float, weight_decay: float,
                 max_norm: Optional[float], emas: Sequence[Dict[str, Tensor]] = (), ema_betas: Sequence[float] = (),
                 feedback_betas: Sequence[Optional[float]] = (), fan_in: Optional[Dict[str, int]] = None) -> Tensor:
    """One optimizer step, in place on every dict of tensors; `step` is the 1-based step count.  Returns the gradient
    norm.  fan_in[name] > 0 marks a weight-normalised tensor."""
    names = list(params)
    if max_norm is not None:
        norm, coef = clip_coef([grads[n] for n in names], max_norm)
    else:
        norm, coef = clip_coef([grads[n] for n in names], float("inf"))
    b1, b2 = betas
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for n in names:
        p, m, v = params[n], exp_avg[n], exp_avg_sq[n]
        g = grads[n].float() * coef                                   # clip_grad_norm_: g.mul_(coef)
        if weight_decay != 0:
            p.mul_(1.0 - lr * weight_decay)                           # torch/optim/adam.py: decoupled decay
        m.copy_(lerp(m, g, 1.0 - b1))                                 # exp_avg.lerp_(grad, 1 - beta1)
        v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
        denom = (v.sqrt() / bc2_sqrt).add_(eps)
        p.addcdiv_(m, denom, value=-step_size)
        for k, ema in enumerate(emas):                                # ema.py:300-313
            e = ema[n]
            e.copy_(lerp(e, p.to(e.dtype), 1.0 - ema_betas[k]))
            fb = feedback_betas[k] if k < len(feedback_betas) else None
            if fb is not None:
                p.copy_(lerp(p, e.to(p.dtype), 1.0 - fb))
        if fan_in is not None and fan_in.get(n, 0) > 0:               # trainer.py:1107-1108
            p.copy_(normalize_rows(p))
    return norm
