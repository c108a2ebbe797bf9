"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's EDM sampler loop
(`DualDiffusionPipeline.diffusion_decode`, src/pipelines/dual_diffusion_pipeline.py:589-752) and sigma schedule
(src/sampling/schedule.py:34-37,58-59) on top of oracle/unet_oracle.py.  Pinned by tests/golden/sampler_small.pt,
which is written by tests/golden/make_golden.py from the unmodified reference run on CPU.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from . import unet_oracle as uo

Tensor = torch.Tensor


def schedule_edm2(steps: int, sigma_max: float, sigma_min: float, rho: float = 7.0) -> Tensor:
    """schedule.py:34-37 + :58-59 (fp32 tensor arithmetic with double Python scalars)."""
    t = torch.linspace(1.0, 0, int(steps) + 1)
    return (sigma_max ** (1 / rho) + (1 - t) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho


def diffusion_decode(sd: Dict[str, Tensor], spec: uo.UNetSpec, audio_embedding: Tensor, sample_shape, *, seed: int,
                     num_steps: int = 100, batch_size: int = 1, cfg_scale: float = 1.5, rho: float = 7.0,
                     use_heun: bool = True, input_perturbation: float = 1.0, input_perturbation_offset: float = 0.0,
                     sigma_max: Optional[float] = None, sigma_min: Optional[float] = None,
                     x_ref: Optional[Tensor] = None, record: Optional[dict] = None, seamless_loop: bool = False,
                     stereo_fix: float = 0.0, stereo_noise: Optional[Tensor] = None) -> Tensor:
    """pipeline.py:598-752 with a CPU torch.Generator (what the reference uses when the module is on CPU).
    `seamless_loop` (:651-656, :729-732; needs x_ref, as in the reference) and `stereo_fix` (:638-640; `stereo_noise`
    stands for the reference's un-seeded `torch.randn_like` draw, taken from the global RNG when None)."""
    sigma_max = sigma_max or spec.sigma_max
    sigma_min = sigma_min or spec.sigma_min
    sigma_data = spec.sigma_data
    B = batch_size
    gen = torch.Generator(device="cpu").manual_seed(seed)
    np_gen = np.random.default_rng(seed)                                                   # :606
    mask = torch.cat((torch.ones(B, dtype=torch.bool), torch.zeros(B, dtype=torch.bool)))
    emb = uo.get_embeddings(sd, audio_embedding, mask)                                     # :608-609
    ref2 = None if x_ref is None else x_ref.repeat(2, 1, 1, 1)
    sig = schedule_edm2(num_steps, sigma_max, sigma_min, rho).tolist()                     # :630-634
    noise = torch.randn(tuple(sample_shape), generator=gen)                                # :637
    if record is not None:
        record["initial_noise"] = noise.clone()
        record["step_noise"] = []
        record["loop_shifts"] = []
    if stereo_fix > 0:                                                                     # :638-640
        noise = noise.clone()
        noise[:, ::2] = noise[:, 1::2]
        fresh = torch.randn_like(noise) if stereo_noise is None else stereo_noise
        if record is not None:
            record["stereo_noise"] = fresh.clone()
        noise = uo.mp_sum(fresh, noise, stereo_fix)
    sample = noise * (sig[0] ** 2 + sigma_data ** 2) ** 0.5                                # :642
    for i, (sigma_curr, sigma_next) in enumerate(zip(sig[:-1], sig[1:])):
        loop_shift = None
        if seamless_loop:                                                                  # :651-656
            loop_shift = int(np_gen.integers(0, sample.shape[-1]))
            if record is not None:
                record["loop_shifts"].append(loop_shift)
            sample = torch.roll(sample, shifts=loop_shift, dims=-1)
            sample = torch.cat((sample[..., -32:], sample, sample[..., :32]), dim=-1)
            ref2 = torch.roll(ref2, shifts=loop_shift, dims=-1)
            ref2 = torch.cat((ref2[..., -32:], ref2, ref2[..., :32]), dim=-1)
        old_sigma_next = sigma_next
        ipo = np.log(sigma_curr) + input_perturbation_offset
        eff = (np.tanh(ipo) / 2 + 0.5) * float(input_perturbation)                        # :691
        sigma_next *= (1 - (max(min(eff, 1), 0)))                                          # :694
        x2 = sample.repeat(2, 1, 1, 1)
        out = uo.unet_forward(sd, spec, x2, torch.tensor([sigma_curr] * 2 * B), emb, ref2)
        cfg = out[B:].lerp(out[:B], cfg_scale)                                             # :701
        if use_heun:
            sigma_hat = max(old_sigma_next, sigma_min)                                     # :706
            t_hat = sigma_hat / sigma_curr
            xh = torch.lerp(cfg, sample, t_hat).repeat(2, 1, 1, 1)                         # :712
            out_h = uo.unet_forward(sd, spec, xh, torch.tensor([t_hat * sigma_curr] * 2 * B), emb, ref2)
            cfg_h = out_h[B:].lerp(out_h[:B], cfg_scale)
            cfg = torch.lerp(cfg, cfg_h, 0.5)                                              # :721
        t = sigma_next / sigma_curr if (i + 1) < num_steps else 0                          # :723
        sample = torch.lerp(cfg, sample, t)
        if loop_shift is not None:                                                         # :729-732
            sample = torch.roll(sample[..., 32:-32], shifts=-loop_shift, dims=-1)
            ref2 = torch.roll(ref2[..., 32:-32], shifts=-loop_shift, dims=-1)
        if i + 1 < num_steps:
            p = max(old_sigma_next ** 2 - sigma_next ** 2, 0) ** 0.5                       # :735
            nz = torch.randn(sample.shape, generator=gen, dtype=sample.dtype)
            if record is not None:
                record["step_noise"].append(nz.clone())
            sample = sample + p * nz
    return sample


def diffusion_decode_unconditional(model_fn, sample_shape, *, seed: int, sigma_max: float, sigma_min: float,
                                   sigma_data: float = 1.0, num_steps: int = 100, rho: float = 7.0, use_heun: bool = True,
                                   input_perturbation: float = 1.0, input_perturbation_offset: float = 0.0,
                                   record: Optional[dict] = None) -> Tensor:
    """pipeline.py:598-752 for a module without class embeddings (`get_embeddings` -> None, e.g. the ddec UNets): one copy
    of the batch, no classifier-free guidance (:659-664, :703-704, :713-720).  `model_fn(x, sigma[1]) -> D_x`."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sig = schedule_edm2(num_steps, sigma_max, sigma_min, rho).tolist()
    noise = torch.randn(tuple(sample_shape), generator=gen)
    if record is not None:
        record["initial_noise"] = noise.clone()
        record["step_noise"] = []
    sample = noise * (sig[0] ** 2 + sigma_data ** 2) ** 0.5
    for i, (sigma_curr, sigma_next) in enumerate(zip(sig[:-1], sig[1:])):
        old_sigma_next = sigma_next
        ipo = np.log(sigma_curr) + input_perturbation_offset
        eff = (np.tanh(ipo) / 2 + 0.5) * float(input_perturbation)
        sigma_next *= (1 - (max(min(eff, 1), 0)))
        cfg = model_fn(sample, torch.tensor([sigma_curr]))
        if use_heun:
            sigma_hat = max(old_sigma_next, sigma_min)
            t_hat = sigma_hat / sigma_curr
            cfg_h = model_fn(torch.lerp(cfg, sample, t_hat), torch.tensor([t_hat * sigma_curr]))
            cfg = torch.lerp(cfg, cfg_h, 0.5)
        t = sigma_next / sigma_curr if (i + 1) < num_steps else 0
        sample = torch.lerp(cfg, sample, t)
        if i + 1 < num_steps:
            p = max(old_sigma_next ** 2 - sigma_next ** 2, 0) ** 0.5
            nz = torch.randn(sample.shape, generator=gen, dtype=sample.dtype)
            if record is not None:
                record["step_noise"].append(nz.clone())
            sample = sample + p * nz
    return sample
