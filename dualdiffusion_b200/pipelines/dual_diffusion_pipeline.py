"""EDM sampler of the drop-in path: `DualDiffusionPipeline.diffusion_decode` with the reference's signature and
step arithmetic (src/pipelines/dual_diffusion_pipeline.py:48-110 SampleParams, :589-752 diffusion_decode).

Per step (Heun + classifier-free guidance) the reference issues 2 UNet calls at batch 2B plus ~10 eager
elementwise kernels and 4 host syncs.  Here a step is: UNet graph replay, `dd_sampler_cfg_lerp`, UNet graph
replay, `dd_sampler_update` — the cond/uncond `.repeat(2,1,1,1)` is folded into the two glue kernels (they
write both halves of the next UNet input), per-step sigmas come from one pre-uploaded device table, and the
debug statistics (`.item()` x4 per step, :740-744) are only computed when `collect_debug_info` is set.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional, Union

import numpy as np
import torch

from .. import ops
from ..sampling.schedule import SamplingSchedule


@dataclass
class SampleParams:
    """pipeline.py:48-70 (same fields and defaults)."""
    seed: Optional[int] = None
    num_steps: int = 100
    batch_size: int = 1
    length: Optional[int] = None
    seamless_loop: bool = False
    cfg_scale: float = 1.5
    sigma_max: Optional[float] = None
    sigma_min: Optional[float] = None
    sigma_data: Optional[float] = None
    rho: float = 7.0
    schedule: Optional[str] = "edm2"
    prompt: Optional[str] = None
    use_heun: bool = True
    input_perturbation: float = 1.0
    input_perturbation_offset: float = 0.0
    stereo_fix: float = 0
    img2img_strength: float = 0.5
    input_audio: Optional[Union[str, torch.Tensor]] = None
    input_audio_pre_encoded: bool = False
    inpainting_mask: Optional[torch.Tensor] = None

    def sanitize(self) -> "SampleParams":
        self.seed = int(self.seed) if self.seed is not None else None
        self.length = int(self.length) if self.length is not None else None
        self.num_steps = int(self.num_steps)
        self.batch_size = int(self.batch_size)
        self.stereo_fix = float(self.stereo_fix)
        return self


def step_scalars(i: int, sigma_curr: float, sigma_next: float, params: SampleParams) -> dict:
    """Host-side per-step scalars of the sampler loop (pipeline.py:680-724): perturbed sigma_next, Heun's
    sigma_hat / t_hat, the final lerp weight t and the re-noise amplitude p."""
    old_sigma_next = sigma_next
    ipo = np.log(sigma_curr) + params.input_perturbation_offset
    eff = (np.tanh(ipo) / 2 + 0.5) * float(params.input_perturbation)             # :691
    sigma_next = sigma_next * (1 - (max(min(eff, 1), 0)))                          # :694
    sigma_hat = max(old_sigma_next, params.sigma_min)                              # :706
    t_hat = sigma_hat / sigma_curr
    last = (i + 1) >= params.num_steps
    t = 0.0 if last else sigma_next / sigma_curr                                   # :723
    p = 0.0 if last else max(old_sigma_next ** 2 - sigma_next ** 2, 0) ** 0.5      # :735
    return dict(sigma_next=float(sigma_next), old_sigma_next=float(old_sigma_next), sigma_hat=float(sigma_hat),
                t_hat=float(t_hat), t=float(t), p=float(p), effective_input_perturbation=float(old_sigma_next - sigma_next))


class EDMSamplerState:
    """Device-resident state of one `diffusion_decode` run, advanced one sampler step at a time.  A step is the
    body of the loop at pipeline.py:649-737: UNet (cond|uncond) -> CFG + Heun predictor -> UNet -> corrector,
    final lerp and re-noise.  `sample2` holds the current sample twice ([cond ; uncond] halves of the UNet batch)."""

    def __init__(self, unet, params: SampleParams, emb: torch.Tensor, sig: list, sample: torch.Tensor,
                 input_ref: Optional[torch.Tensor], fmt, collect_debug_info: bool = False) -> None:
        self.unet, self.params, self.emb, self.sig, self.fmt = unet, params, emb, sig, fmt
        self.input_ref = input_ref
        device = sample.device
        B = params.batch_size
        self.B = B
        # unconditional modules (get_embeddings -> None, e.g. the ddec UNets) run one copy of the batch and no CFG
        # (pipeline.py:659-664, :703-704, :713-720)
        self.uncond = emb is None
        rep = 1 if self.uncond else 2
        self.flags = 2 if self.uncond else 1          # dd_sampler_* `dup` bits: 0 = duplicate halves, 1 = unconditional
        self.steps = [step_scalars(i, sig[i], sig[i + 1], params) for i in range(params.num_steps)]
        # device table of the per-call sigma vectors: [step][0 = sigma_curr | 1 = sigma_hat][rep * B]
        self.table = torch.tensor([[[sig[i]] * (rep * B), [st["t_hat"] * sig[i]] * (rep * B)]
                                   for i, st in enumerate(self.steps)], dtype=torch.float32).to(device)
        self.sample2 = torch.empty((rep * B,) + tuple(sample.shape[1:]), device=device, dtype=torch.float32)
        self.sample2[:B] = sample
        if not self.uncond:
            self.sample2[B:] = sample
        self.xhat2 = torch.empty_like(self.sample2)
        self.cfg1 = torch.empty_like(sample)
        self.cfg_out = torch.empty_like(sample) if collect_debug_info else None

    @property
    def sample(self) -> torch.Tensor:
        return self.sample2[:self.B]

    def set_sample(self, sample: torch.Tensor) -> None:
        """Overwrite the current sample (both UNet-batch halves), e.g. from a pinned host buffer."""
        self.sample2[:self.B].copy_(sample, non_blocking=True)
        if not self.uncond:
            self.sample2[self.B:].copy_(sample, non_blocking=True)

    def step(self, i: int, noise: Optional[torch.Tensor]) -> None:
        st, p = self.steps[i], self.params
        d1 = self.unet(self.sample2, self.table[i, 0], self.fmt, self.emb, self.input_ref)
        ops.sampler_cfg_lerp(d1, self.sample2, p.cfg_scale, st["t_hat"], self.cfg1,
                             self.xhat2 if p.use_heun else None, dup=self.flags)
        d2 = self.unet(self.xhat2, self.table[i, 1], self.fmt, self.emb, self.input_ref) if p.use_heun else None
        ops.sampler_update(self.cfg1, d2, p.cfg_scale, p.use_heun, st["t"], st["p"] if noise is not None else 0.0,
                           noise, self.sample2, self.cfg_out, dup=self.flags)


class DualDiffusionPipeline(torch.nn.Module):
    """Minimal module container exposing the sampler.  Construct with the modules that a model directory's
    model_index.json would load (`unet`, optionally `format`, ...)."""

    def __init__(self, pipeline_modules: dict) -> None:
        super().__init__()
        for name, module in pipeline_modules.items():
            setattr(self, name, module)
        self.collect_debug_info = False
        self.last_debug_info: dict = {}

    @torch.inference_mode()
    def prepare_sampler(self, params: SampleParams, audio_embedding: torch.Tensor, sample_shape=None,
                        x_ref: Optional[torch.Tensor] = None, module=None,
                        initial_noise: Optional[torch.Tensor] = None, stereo_noise: Optional[torch.Tensor] = None):
        """Everything `diffusion_decode` does before its loop (pipeline.py:598-646).  Returns (state, generator)."""
        unet = module or getattr(self, "unet")
        params = SampleParams(**params.__dict__).sanitize()
        params.seed = params.seed or int(np.random.randint(100000, 999999))
        params.sigma_max = params.sigma_max or unet.config.sigma_max
        params.sigma_min = params.sigma_min or unet.config.sigma_min
        params.sigma_data = params.sigma_data or unet.config.sigma_data
        if params.seamless_loop and x_ref is None:
            # the reference rolls `input_ref_sample` unconditionally (pipeline.py:655) and fails on None
            raise ValueError("seamless_loop needs x_ref (pipeline.py:655 rolls input_ref_sample)")
        if sample_shape is None and x_ref is None:                          # pipeline.py:615-622
            fmt = getattr(self, "format", None)
            if fmt is None:
                raise ValueError("sample_shape, x_ref or a `format` module is required")
            length = params.length or fmt.config.default_raw_length
            sample_shape = fmt.get_mel_spec_shape(bsz=params.batch_size, raw_length=length)
            dae = getattr(self, "dae", None)
            if dae is not None:
                sample_shape = dae.get_latent_shape(sample_shape)
        device = torch.device(unet.device)
        B = params.batch_size
        generator = torch.Generator(device=device).manual_seed(params.seed)
        conditioning_mask = torch.cat((torch.ones(B, dtype=torch.bool), torch.zeros(B, dtype=torch.bool)))
        emb = unet.get_embeddings(audio_embedding, conditioning_mask.to(device))
        input_ref = None
        if x_ref is not None:
            sample_shape = sample_shape or x_ref.shape
            input_ref = x_ref.to(device=device, dtype=torch.float32)
            if emb is not None:
                input_ref = input_ref.repeat(2, 1, 1, 1)                     # pipeline.py:626-629
        sample_shape = tuple(sample_shape)
        schedule = SamplingSchedule.get_schedule(params.schedule, params.num_steps, 1, device="cpu",
                                                 sigma_max=params.sigma_max, sigma_min=params.sigma_min, rho=params.rho)
        sig = schedule.tolist()
        if initial_noise is None:
            noise = torch.randn(sample_shape, device=device, generator=generator)
        else:
            noise = initial_noise.to(device=device, dtype=torch.float32)
        if params.stereo_fix > 0:                                             # pipeline.py:638-640
            # the reference draws this second noise from the global RNG (`torch.randn_like`, no generator)
            fresh = (torch.randn_like(noise) if stereo_noise is None
                     else stereo_noise.to(device=device, dtype=torch.float32))
            noise = ops.stereo_fix_noise(noise.contiguous(), fresh.contiguous(), params.stereo_fix)
        sample = noise * (sig[0] ** 2 + params.sigma_data ** 2) ** 0.5
        state = EDMSamplerState(unet, params, emb, sig, sample, input_ref, getattr(self, "format", None),
                                self.collect_debug_info)
        return state, generator

    @torch.inference_mode()
    def diffusion_decode(self, params: SampleParams, quiet: bool = False,
                         audio_embedding: Optional[torch.Tensor] = None, sample_shape: Optional[torch.Size] = None,
                         x_ref: Optional[torch.Tensor] = None, module=None,
                         initial_noise: Optional[torch.Tensor] = None,
                         step_noise: Optional[Any] = None, stereo_noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pipeline.py:589-752.  `initial_noise` / `step_noise` (callable i -> tensor) / `stereo_noise` optionally inject
        the noise draws so that parity tests can feed the oracle's values; by default they come from a device
        `torch.Generator` seeded with params.seed exactly as in the reference (:605, :637, :736) and, for stereo_fix,
        from the global RNG (:640).

        seamless_loop (:651-656, :729-732): every step runs on the sample rolled by a random shift (numpy generator
        seeded with params.seed, :606) and circularly padded by 32 columns per side.  The UNet calls, CFG, Heun and the
        final lerp + re-noise all run in that padded frame (`loop` state: its own CUDA graph at width W + 64); the
        re-noise commutes with the index map, so the step's noise is rolled / padded the same way and the result is
        cropped and rolled back -- elementwise identical to the reference's order (crop, roll back, add noise)."""
        state, generator = self.prepare_sampler(params, audio_embedding, sample_shape, x_ref, module, initial_noise,
                                                stereo_noise)
        p = state.params
        debug_info: dict = {"sigma_schedule": state.sig}
        shape = tuple(state.sample.shape)
        device = state.sample2.device
        loop = None
        if p.seamless_loop:
            pad, W = 32, shape[-1]
            if W < pad:
                raise ValueError(f"seamless_loop: sample width {W} is smaller than the 32-column loop padding")
            np_generator = np.random.default_rng(p.seed)                                  # :606
            rep = 1 if state.uncond else 2
            ref_src = x_ref.to(device=device, dtype=torch.float32).contiguous()
            ref_pad = torch.empty((rep * ref_src.shape[0],) + tuple(ref_src.shape[1:-1]) + (W + 2 * pad,),
                                  device=device, dtype=torch.float32)
            loop = EDMSamplerState(state.unet, p, state.emb, state.sig,
                                   torch.zeros(shape[:-1] + (W + 2 * pad,), device=device, dtype=torch.float32),
                                   ref_pad, state.fmt, self.collect_debug_info)
            cfg_crop = torch.empty(shape, device=device, dtype=torch.float32) if self.collect_debug_info else None
        for i in range(p.num_steps):
            nz = None
            if (i + 1) < p.num_steps:
                nz = (step_noise(i).to(device=device, dtype=torch.float32) if step_noise is not None
                      else torch.randn(shape, generator=generator, device=device, dtype=torch.float32))
            if loop is None:
                state.step(i, nz)
                cfg_dbg = state.cfg_out
            else:
                shift = int(np_generator.integers(0, W))                                  # :652
                ops.roll_pad_w(state.sample, shift, pad, copies=rep, out=loop.sample2)    # :653-654, :661
                ops.roll_pad_w(ref_src, shift, pad, copies=rep, out=ref_pad)              # :655-656
                loop.step(i, ops.roll_pad_w(nz.contiguous(), shift, pad) if nz is not None else None)
                ops.crop_unroll_w(loop.sample, shift, pad, out=state.sample)              # :730
                cfg_dbg = ops.crop_unroll_w(loop.cfg_out, shift, pad, out=cfg_crop) if self.collect_debug_info else None
            if self.collect_debug_info:   # pipeline.py:740-744 (forces host syncs)
                debug_info.setdefault("sample_std", []).append(state.sample.std().item())
                debug_info.setdefault("cfg_output_mean", []).append(cfg_dbg.mean().item())
                debug_info.setdefault("cfg_output_std", []).append(cfg_dbg.std().item())
                debug_info.setdefault("effective_input_perturbation", []).append(
                    state.steps[i]["effective_input_perturbation"])
        self.last_debug_info = debug_info
        return state.sample.clone()
