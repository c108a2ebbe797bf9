"""Noise-level schedules (reference src/sampling/schedule.py:30-79): same names, arguments and formulas.
Host-side scalar math: the sampler immediately converts the schedule to Python floats (pipeline.py:630-635)."""
from __future__ import annotations

import inspect
from typing import Any, Optional

import numpy as np
import torch


class SamplingSchedule:

    @staticmethod
    @torch.no_grad()
    def get_schedule(name: str, steps: int, t_start: float = 1.0, device: Optional[torch.device] = None,
                     **kwargs) -> torch.Tensor:
        fn = getattr(SamplingSchedule, f"schedule_{name}")
        t = torch.linspace(t_start, 0, int(steps) + 1, device=device)
        return fn(t, **kwargs)

    @staticmethod
    def get_schedule_params(name: str) -> dict[str, type[Any]]:
        params = {k: v.annotation
                  for k, v in inspect.signature(getattr(SamplingSchedule, f"schedule_{name}")).parameters.items()}
        for k in ("t", "_", "sigma_max", "sigma_min"):
            params.pop(k, None)
        return params

    @classmethod
    def get_schedules_list(cls) -> list[str]:
        return [a.removeprefix("schedule_") for a in dir(cls)
                if callable(getattr(cls, a)) and a.startswith("schedule_")]

    @staticmethod
    def schedule_edm2(t: torch.Tensor, sigma_max: float, sigma_min: float, rho: float = 7.0, **_) -> torch.Tensor:
        return (sigma_max ** (1 / rho) + (1 - t) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho

    @staticmethod
    def schedule_ln_linear(t: torch.Tensor, sigma_max: float, sigma_min: float, **_) -> torch.Tensor:
        return (np.log(sigma_min) + (np.log(sigma_max) - np.log(sigma_min)) * t).exp()

    @staticmethod
    def schedule_linear(t: torch.Tensor, sigma_max: float, sigma_min: float, rho: float = 1.0, **_) -> torch.Tensor:
        t = (sigma_max ** (1 / rho) - sigma_min ** (1 / rho)) * t + sigma_min ** (1 / rho)
        return t ** rho

    @staticmethod
    def schedule_cos(t: torch.Tensor, sigma_max: float, sigma_min: float, rho: float = 1.0, **_) -> torch.Tensor:
        theta_max = np.pi / 2 - np.arctan(sigma_max / rho)
        theta_min = np.pi / 2 - np.arctan(sigma_min / rho)
        theta = (1 - t) * (theta_min - theta_max) + theta_max
        return theta.cos() / theta.sin() * rho

    @staticmethod
    def schedule_scale_invariant(t: torch.Tensor, sigma_max: float, sigma_min: float, rho: float = 1.0,
                                 **_) -> torch.Tensor:
        return sigma_min / ((1 - t) ** rho + sigma_min / sigma_max)
