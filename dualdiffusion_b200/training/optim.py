"""Optimizer-side sweep of the train step on the B200: a drop-in for the `torch.optim.AdamW`
the reference trainer builds (training/trainer.py:461-473) that can also absorb the three parameter sweeps the trainer
runs around `optimizer.step()`:

    accelerator.clip_grad_norm_(params, max_norm)     trainer.py:1044      -> FusedAdamW.clip_grad_norm_(max_norm)
    optimizer.step()                                  trainer.py:1062      -> FusedAdamW.step()
    ema_manager.update()  (_foreach_lerp_ per EMA)    training/ema.py:284-313   } folded into step() once
    module.normalize_weights()                        trainer.py:1107-1108      } attach_emas / attach_module were called

`clip_grad_norm_` is two launches (deterministic two-stage reduction; the norm and the clip coefficient stay on the
device, the coefficient is consumed by the next `step()` so the gradients are never rewritten), `step()` is ONE launch
per parameter group (`dd_optim_step_batched`, a CTA per weight row, 36 B of HBM traffic per parameter + 8 B per fp32
EMA copy instead of ~56 B + 12 B per EMA for the separate sweeps).  The optimizer state uses torch AdamW's keys
(`step`, `exp_avg`, `exp_avg_sq`), so `load_state_dict` of a reference checkpoint's optimizer state works.

There is no CPU or PyTorch fallback: parameters, gradients and EMA copies must be contiguous fp32 (EMA: fp32 or fp64)
CUDA tensors, anything else raises.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from .. import _lib as L
from .. import ops

Tensor = torch.Tensor


def weight_norm_fan_in(module: torch.nn.Module) -> Dict[int, int]:
    """{id(weight): fan_in} for every weight `module.normalize_weights()` would re-normalise: the submodules that have a
    `normalize_weights` method, a `weight` and no `disable_weight_norm` (MPConv, mp_tools.py:375-378: one norm per output row).  Host only."""
    out: Dict[int, int] = {}
    for m in module.modules():
        if m is module or not hasattr(m, "normalize_weights") or not isinstance(getattr(m, "weight", None), Tensor):
            continue
        if getattr(m, "disable_weight_norm", False):
            continue
        if getattr(m, "norm_dim", None) is not None:    # MPConv3D(norm_dim=1): strided (o, tap) vectors over Cin
            raise NotImplementedError(f"fused normalize_weights: norm_dim={m.norm_dim} is not a per-output-row norm "
                                      "(DAE training is outside the built path)")
        w = m.weight
        if w.ndim < 2:
            raise NotImplementedError("fused normalize_weights: weight without an input dimension")
        out[id(w)] = w.numel() // w.shape[0]
    return out


class FusedAdamW(torch.optim.Optimizer):
    """AdamW with decoupled weight decay, torch.optim.AdamW's constructor arguments and state layout."""

    def __init__(self, params: Iterable, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, amsgrad: bool = False, *, maximize: bool = False, foreach=None,
                 capturable: bool = False, differentiable: bool = False, fused=None) -> None:
        if isinstance(lr, Tensor):
            lr = float(lr)
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid betas: {betas}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        if amsgrad or maximize or capturable or differentiable:
            raise NotImplementedError("FusedAdamW: amsgrad / maximize / capturable / differentiable are not supported "
                                      "(the reference trainer uses none of them, trainer.py:466-473)")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self._fan_in: Dict[int, int] = {}
        self._emas: List[List[Tensor]] = []          # per EMA: tensors in the order of self._all_params()
        self._ema_betas: List[float] = []
        self._ema_feedback: List[Optional[float]] = []
        self._norm_coef: Optional[Tensor] = None     # device [2]: {grad norm, clip coefficient}
        self._coef_pending = False
        self._cache: dict = {}

    # ---- what to fold into step() ------------------------------------------------------------------------
    def _all_params(self) -> List[Tensor]:
        return [p for g in self.param_groups for p in g["params"]]

    def attach_module(self, *modules: torch.nn.Module) -> None:
        """Fold `module.normalize_weights()` (trainer.py:1107-1108) of these modules into step()."""
        for m in modules:
            self._fan_in.update(weight_norm_fan_in(m))
        self._cache.clear()

    def attach_emas(self, ema_params: Sequence[Sequence[Tensor]], betas: Sequence[float],
                    feedback_betas: Optional[Sequence[Optional[float]]] = None) -> None:
        """Fold `EMA_Manager.update()` (ema.py:284-313) into step().  `ema_params[k]` are the parameters of EMA copy k
        in the same order as the optimizer's parameters (both come from `module.parameters()` of deep copies)."""
        n = len(self._all_params())
        if len(ema_params) > L.OPTIM_MAX_EMA:
            raise ValueError(f"at most {L.OPTIM_MAX_EMA} EMA copies can be folded into one step()")
        emas = [list(e) for e in ema_params]
        for e in emas:
            if len(e) != n:
                raise ValueError(f"EMA copy has {len(e)} parameters, the optimizer {n}")
        self._emas = emas
        self.set_ema_betas(betas, feedback_betas)
        self._cache.clear()

    def set_ema_betas(self, betas: Sequence[float], feedback_betas: Optional[Sequence[Optional[float]]] = None) -> None:
        """Per-step effective betas (power-function EMAs and warm-up change them every step, ema.py:300-304)."""
        if len(betas) != len(self._emas):
            raise ValueError("one beta per attached EMA copy")
        fb = list(feedback_betas) if feedback_betas is not None else (self._ema_feedback or [None] * len(betas))
        if len(fb) != len(betas):
            raise ValueError("one feedback beta (or None) per attached EMA copy")
        self._ema_betas = [float(b) for b in betas]
        self._ema_feedback = [None if b is None else float(b) for b in fb]

    # ---- clip_grad_norm_ ---------------------------------------------------------------------------------
    @torch.no_grad()
    def clip_grad_norm_(self, max_norm: float) -> Tensor:
        """torch.nn.utils.clip_grad_norm_ over every parameter of the optimizer: returns the total norm (0-d device
        tensor).  The gradients are left untouched; the clip coefficient is applied inside the next step()."""
        grads = [p.grad for p in self._all_params() if p.grad is not None]
        if not grads:
            return torch.zeros((), dtype=torch.float32)
        for g in grads:
            _require_f32_cuda(g, "gradient")
        dev = grads[0].device
        key = ("gnorm", tuple(g.data_ptr() for g in grads), tuple(g.numel() for g in grads))
        st = self._cache.get("gnorm")
        if st is None or st[0] != key:
            arr, chunks = ops.pack_gnorm_descs(grads)
            st = (key, ops.descs_to_device(arr, dev), len(grads), chunks,
                  torch.empty(chunks, device=dev, dtype=torch.float32))
            self._cache["gnorm"] = st
        if self._norm_coef is None or self._norm_coef.device != dev:
            self._norm_coef = torch.empty(2, device=dev, dtype=torch.float32)
        ops.grad_norm_clip(st[1], st[2], st[3], st[4], float(max_norm), self._norm_coef)
        self._coef_pending = True
        return self._norm_coef[0]

    # ---- step --------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        index = {id(p): i for i, p in enumerate(self._all_params())}
        coef = self._norm_coef if self._coef_pending else None
        for gi, group in enumerate(self.param_groups):
            by_step: Dict[float, List[Tensor]] = {}
            for p in group["params"]:
                if p.grad is None:
                    # torch's AdamW skips it; EMA_Manager.update() / normalize_weights() still cover it (g = NULL)
                    if self._emas or id(p) in self._fan_in:
                        _require_f32_cuda(p, "parameter")
                        by_step.setdefault(-1.0, []).append(p)
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdamW does not support sparse gradients")
                _require_f32_cuda(p, "parameter")
                _require_f32_cuda(p.grad, "gradient")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.zeros((), dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                elif isinstance(st["step"], Tensor) and st["step"].is_cuda:      # a checkpoint written by fused=True
                    st["step"] = st["step"].detach().cpu().float()
                st["step"] += 1
                by_step.setdefault(float(st["step"]), []).append(p)
            for step_t, plist in by_step.items():
                self._launch(gi, group, step_t, plist, index, coef)
        self._coef_pending = False
        return loss

    def _launch(self, gi: int, group: dict, step_t: float, plist: List[Tensor], index: Dict[int, int],
                coef: Optional[Tensor]) -> None:
        dev = plist[0].device
        emas = [[e[index[id(p)]] for p in plist] for e in self._emas]
        for e in emas:
            for t, p in zip(e, plist):
                if not t.is_cuda or t.dtype not in (torch.float32, torch.float64) or not t.is_contiguous() \
                        or t.numel() != p.numel():
                    raise RuntimeError("FusedAdamW: EMA copies must be contiguous fp32/fp64 CUDA tensors of the "
                                       "parameter's size (ema cpu_offload is not supported)")
        is_f64 = [int(e[0].dtype == torch.float64) for e in emas]
        for e, f in zip(emas, is_f64):
            if any(int(t.dtype == torch.float64) != f for t in e):
                raise RuntimeError("FusedAdamW: one EMA copy mixes fp32 and fp64 tensors")
        no_grad = step_t < 0                 # the group of parameters without a gradient: EMA + re-normalisation only
        if no_grad:
            step_t = 1.0
            key = (tuple(p.data_ptr() for p in plist), tuple(t.data_ptr() for e in emas for t in e))
        else:
            key = (tuple(p.data_ptr() for p in plist), tuple(p.grad.data_ptr() for p in plist),
                   tuple(self.state[p]["exp_avg"].data_ptr() for p in plist),
                   tuple(self.state[p]["exp_avg_sq"].data_ptr() for p in plist),
                   tuple(t.data_ptr() for e in emas for t in e))
        slot = ("step", gi, len(plist), no_grad)
        st = self._cache.get(slot)
        if st is None or st[0] != key:
            entries = [dict(p=p, g=None if no_grad else p.grad, m=None if no_grad else self.state[p]["exp_avg"],
                            v=None if no_grad else self.state[p]["exp_avg_sq"],
                            emas=[e[i] for e in emas], fan_in=self._fan_in.get(id(p), 0))
                       for i, p in enumerate(plist)]
            arr, rows = ops.pack_optim_descs(entries)
            st = (key, ops.descs_to_device(arr, dev), len(entries), rows)
            self._cache[slot] = st
        hyper = make_hyper(group["lr"], group["betas"], group["eps"], group["weight_decay"], step_t,
                           self._ema_betas, self._ema_feedback, is_f64)
        ops.optim_step_batched(st[1], st[2], st[3], hyper, coef)
        for p in plist:                      # the raw-pointer write must invalidate cached prepared weights
            torch.autograd.graph.increment_version(p)


def make_hyper(lr: float, betas: Tuple[float, float], eps: float, weight_decay: float, step: float,
               ema_betas: Sequence[float] = (), ema_feedback: Sequence[Optional[float]] = (),
               ema_is_f64: Sequence[int] = ()) -> "L.OptimHyper":
    """dd_optim_hyper for one launch (bias corrections in double, as torch/optim/adam.py computes them).  Host only."""
    if isinstance(lr, Tensor):
        lr = float(lr)
    h = L.OptimHyper()
    h.lr, h.beta1, h.beta2, h.eps, h.weight_decay = float(lr), float(betas[0]), float(betas[1]), float(eps), \
        float(weight_decay)
    h.bias_correction1 = 1.0 - float(betas[0]) ** step
    h.bias_correction2 = 1.0 - float(betas[1]) ** step
    h.n_ema = len(ema_betas)
    for k in range(L.OPTIM_MAX_EMA):
        h.ema_beta[k] = float(ema_betas[k]) if k < len(ema_betas) else 0.0
        fb = ema_feedback[k] if k < len(ema_feedback) else None
        h.feedback_beta[k] = -1.0 if fb is None else float(fb)
        h.ema_is_f64[k] = int(ema_is_f64[k]) if k < len(ema_is_f64) else 0
    return h


def _require_f32_cuda(t: Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"dualdiffusion_b200 has no CPU path: {what} must live on a CUDA device")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise RuntimeError(f"FusedAdamW: {what} must be a contiguous fp32 tensor (got {t.dtype})")
