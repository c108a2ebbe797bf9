#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native dualdiffusion denoising hot path.

Workload (BASELINE.json configs[1]): EDM sampler (Heun + classifier-free guidance) on the 45 s @ 32 kHz stereo
latent shape (1 x 4 x 32 x 688), default 293 M-parameter EDM2 UNet, bf16 tensor-core compute, fp32 sampler state.
A *step* is one sampler step = 2 UNet evaluations at batch 2 (cond | uncond) + the CFG/Heun glue
(reference src/pipelines/dual_diffusion_pipeline.py:649-737).  Weights are random-init (seeded), inputs synthetic.

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, C-ABI library)
  python bench.py --impl reference --steps K --warmup W    # reference arm: the CPU oracle port on the host cores

Prints ONE JSON line.  `value` = sampler steps/s with the state resident in HBM (CUDA events on the launching
stream, max over ranks); `e2e` = the same step driven with HOST buffers (pinned H2D of the sample in, D2H of the
new sample out, inside the timed region); `roofline` = the dominant kernel (tcgen05 implicit-GEMM MPConv) timed
live with CUDA events; `cpu_baseline` = the oracle port timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from dualdiffusion_b200.replicas import max_over_ranks  # noqa: E402

LATENT = (1, 4, 32, 688)             # 45 s @ 32 kHz stereo -> mel (2,256,5504) -> 8x-downsampled 4-channel latent
FLOP_PER_SAMPLE_FWD = 0.489e12       # default UNet forward at (4,32,688)
FLOP_PER_STEP = 4 * FLOP_PER_SAMPLE_FWD
METRIC = "UNet denoise steps/sec on 45s@32kHz-stereo latents (EDM sampler step: Heun + CFG = 2 UNet calls x batch 2)"
UNIT = "steps/s"
# identical in both arms (the driver compares the two `config` objects); arm-specific details go to `config_details`
CONFIG = {"workload": "EDM sampler step (Heun+CFG), default EDM2 UNet 293M, latent 1x4x32x688 (45 s stereo)"}


def peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {"hbm_gbs": d.get("hbm_gbs"), "tflops": d.get("bf16_tflops_sustained") or d.get("bf16_tflops"),
                "tflops_burst": d.get("bf16_tflops"), "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clock / throttle-reason sampling during the timed region."""

    def __init__(self, index: int) -> None:
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self) -> None:
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self) -> dict:
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# --------------------------------------------------------------------------------------------------
# CPU oracle arm (cpu_baseline of our arm, and the whole of --impl reference)
# --------------------------------------------------------------------------------------------------
def cpu_oracle_step_rate(steps: int, warmup: int, budget_s: float) -> dict:
    """Times the CPU oracle port (oracle/, plain PyTorch fp32 -- a restatement of the reference's own PyTorch code
    path, which cannot travel to this box) on all host cores.  One sampler step = 4 sample-forwards at 1x4x32x688."""
    from oracle import unet_oracle as uo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = uo.default_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    g = torch.Generator().manual_seed(1)
    x1 = torch.randn(LATENT, generator=g)
    emb1 = uo.get_embeddings(sd, torch.randn(1, spec.in_channels_emb, generator=g), torch.tensor([True]))
    with torch.no_grad():
        t0 = time.perf_counter()
        uo.unet_forward(sd, spec, x1, torch.tensor([3.0]), emb1)        # warm-up + calibration
        t_fwd = time.perf_counter() - t0
        full_step = 4 * t_fwd * (steps + warmup) <= budget_s
        if full_step:
            x2 = x1.repeat(2, 1, 1, 1)
            emb2 = uo.get_embeddings(sd, torch.randn(1, spec.in_channels_emb, generator=g), torch.tensor([True, False]))
            def one():                                                   # 2 UNet calls at batch 2
                d = uo.unet_forward(sd, spec, x2, torch.tensor([3.0, 3.0]), emb2)
                cfg = d[1:].lerp(d[:1], 1.5)
                d = uo.unet_forward(sd, spec, torch.lerp(cfg, x1, 0.9).repeat(2, 1, 1, 1), torch.tensor([2.7, 2.7]), emb2)
                return d
            sample = "full sampler steps (2 UNet calls x batch 2, 1x4x32x688, fp32)"
            per_step = 1.0
        else:
            def one():
                return uo.unet_forward(sd, spec, x1, torch.tensor([3.0]), emb1)
            sample = "single sample-forwards (1x4x32x688, fp32) = 1/4 sampler step each, scaled x4"
            per_step = 0.25
        for _ in range(max(0, warmup - 1) if full_step else 0):
            one()
        n = max(1, steps if full_step else min(steps, max(1, int(budget_s / max(t_fwd, 1e-3)))))
        t0 = time.perf_counter()
        for _ in range(n):
            one()
        dt = time.perf_counter() - t0
    return {"value": per_step * n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} x {sample}; oracle/unet_oracle.py on {cores} host threads", "seconds": dt}


def _reference_unet(device, dtype, memory_format=None):
    """The UNMODIFIED reference UNet (baseline/_ref/src, staged by __graft_entry__.build) with the same seeded synthetic
    weights as our arm, through baseline/ref_loader.py (stubs for absent I/O-only imports)."""
    from baseline import ref_loader
    ref_loader.install()
    from oracle import unet_oracle as uo          # seeded synthetic weights only
    from modules.unets.unet_edm2_b4 import UNet as RefUNet, UNetConfig as RefUNetConfig
    from modules.formats.ms_mdct_dual import MS_MDCT_DualFormat, MS_MDCT_DualFormatConfig
    spec = uo.default_spec()
    cfg = RefUNetConfig(**{k: getattr(spec, k) for k in RefUNetConfig.__dataclass_fields__ if hasattr(spec, k)})
    net = RefUNet(cfg)
    net.load_state_dict(uo.synth_state_dict(spec, seed=0), strict=True)
    net = net.requires_grad_(False).train(False).to(device=device, dtype=dtype, memory_format=memory_format)
    fmt = MS_MDCT_DualFormat(MS_MDCT_DualFormatConfig()).to(device=device)
    return net, fmt, spec


def cpu_reference_step_rate(steps: int, warmup: int, budget_s: float) -> dict:
    """The reference's own CPU path: its UNet (fp32, all host threads) driven like one sampler step of
    pipeline.py:649-737 (2 UNet calls at batch 2 + CFG / Heun lerps)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net, fmt, spec = _reference_unet("cpu", torch.float32)
    g = torch.Generator().manual_seed(1)
    x1 = torch.randn(LATENT, generator=g)
    clap = torch.randn(1, spec.in_channels_emb, generator=g)
    with torch.no_grad():
        emb1 = net.get_embeddings(clap, torch.tensor([True]))
        t0 = time.perf_counter()
        net(x1, torch.tensor([3.0]), fmt, emb1)
        t_fwd = time.perf_counter() - t0
        full_step = 4 * t_fwd * (steps + warmup) <= budget_s
        if full_step:
            emb2 = net.get_embeddings(clap, torch.tensor([True, False]))
            x2 = x1.repeat(2, 1, 1, 1)

            def one():
                d = net(x2, torch.tensor([3.0, 3.0]), fmt, emb2).float()
                cfg = d[1:].lerp(d[:1], 1.5)
                return net(torch.lerp(cfg, x1, 0.9).repeat(2, 1, 1, 1), torch.tensor([2.7, 2.7]), fmt, emb2)
            sample, per_step = "full sampler steps (2 UNet calls x batch 2, 1x4x32x688, fp32)", 1.0
        else:
            def one():
                return net(x1, torch.tensor([3.0]), fmt, emb1)
            sample, per_step = "single sample-forwards (1x4x32x688, fp32) = 1/4 sampler step each, scaled x4", 0.25
        for _ in range(max(0, warmup - 1) if full_step else 0):
            one()
        n = max(1, steps if full_step else min(steps, max(1, int(budget_s / max(t_fwd, 1e-3)))))
        t0 = time.perf_counter()
        for _ in range(n):
            one()
        dt = time.perf_counter() - t0
    return {"value": per_step * n / dt, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": f"{n} x {sample}; the reference's own modules/unets/unet_edm2_b4.UNet on {cores} host threads",
            "seconds": dt}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    note = "the unmodified reference UNet (baseline/_ref/src) on the host cores"
    try:
        r = cpu_reference_step_rate(args.steps, args.warmup, budget_s=150.0)
    except Exception as exc:          # reference tree not staged on this box: the validated CPU port stands in
        note = f"reference tree unavailable ({exc!r}); the CPU oracle port (validated against it, tests/test_oracle.py) is timed"
        r = cpu_oracle_step_rate(args.steps, args.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / r["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": CONFIG, "config_details": {"note": note},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def build_sampler(device, steps_total: int):
    from oracle import unet_oracle as uo          # seeded synthetic weights only (no oracle compute here)
    from dualdiffusion_b200.modules.unets.unet_edm2_b4 import UNet, UNetConfig
    from dualdiffusion_b200.pipelines.dual_diffusion_pipeline import DualDiffusionPipeline, SampleParams
    spec = uo.default_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    cfg = UNetConfig(**{k: getattr(spec, k) for k in UNetConfig.__dataclass_fields__ if hasattr(spec, k)})
    net = UNet(cfg)
    net.load_state_dict(sd, strict=True)
    net = net.requires_grad_(False).train(False).to(device=device)
    pipe = DualDiffusionPipeline({"unet": net})
    params = SampleParams(seed=1234, num_steps=100, batch_size=1, cfg_scale=1.5, use_heun=True)
    g = torch.Generator().manual_seed(7)
    clap = torch.randn(1, spec.in_channels_emb, generator=g)
    state, gen = pipe.prepare_sampler(params, clap, LATENT)
    return net, pipe, state, gen


def bench_format(device) -> dict:
    """Second half of BASELINE.json's metric (configs[2]): mel-STFT encode + FGLA decode, batch 64 synthetic 45 s stereo
    waveforms, fp32, full 200 FGLA iterations.  HBM-bound by design: algorithmic bytes are
    raw-in + mel-out for the encoder and 28 B per STFT bin per iteration (+ the waveform) for FGLA."""
    from dualdiffusion_b200.modules.formats.spectrogram import SpectrogramFormat, SpectrogramFormatConfig
    pk = peaks()
    fmt = SpectrogramFormat(SpectrogramFormatConfig())
    B = 64
    Ls = fmt.sample_raw_crop_width(1408768)
    g = torch.Generator(device=device).manual_seed(0)
    raw = 0.1 * torch.randn(B, 2, Ls, device=device, generator=g)

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / n, out

    fmt.raw_to_sample(raw[:2])
    t_enc, mel = timed(lambda: fmt.raw_to_sample(raw), 3)
    fmt.sample_to_raw(mel[:1], n_fgla_iters=2)
    iters = fmt.config.num_fgla_iters
    t_dec, wave = timed(lambda: fmt.sample_to_raw(mel, n_fgla_iters=iters), 1)
    bins = fmt.config.num_stft_bins * mel.shape[-1]
    enc_bytes = (raw.numel() + mel.numel()) * 4
    dec_bytes = (28 * bins + 2 * 4 * Ls) * (B * 2) * iters
    mdct = None
    try:
        # SURVEY 8(f) N1, the live decode path: MCLT / inverse MCLT / mel -> MDCT-PSD of MS_MDCT_DualFormat at batch 16
        from dualdiffusion_b200.modules.formats.ms_mdct_dual import MS_MDCT_DualFormat, MS_MDCT_DualFormatConfig
        mfmt = MS_MDCT_DualFormat(MS_MDCT_DualFormatConfig())
        Bm = 16
        rawm = 0.1 * torch.randn(Bm, 2, mfmt.get_raw_crop_width(1408768), device=device, generator=g)
        coef = mfmt.raw_to_mdct(rawm[:1])
        t_f, coef = timed(lambda: mfmt.raw_to_mdct(rawm), 3)
        t_i, back = timed(lambda: mfmt.mdct_to_raw(coef), 3)
        melm = mfmt.raw_to_mel_spec(rawm)
        t_p, psd = timed(lambda: mfmt.mel_spec_to_mdct_psd(melm), 3)

        def leg(t, nbytes, flop):
            return {"value": Bm / t, "unit": "stereo samples/s", "ms": t * 1e3, "gflops": flop / t / 1e9,
                    "roofline": {"bound": "hbm", "achieved": nbytes / t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                 "frac": nbytes / t / 1e9 / pk["hbm_gbs"]}}
        S_, N2, T_ = Bm * 2, coef.shape[2] * 2, coef.shape[3]
        mdct = {"what": "MS_MDCT_DualFormat at batch 16 (45 s stereo): repo fp32 GEMM kernel (dd_gemm_f32), frames gathered on load",
                "raw_to_mdct": leg(t_f, (rawm.numel() + coef.numel()) * 4, 2.0 * S_ * N2 * N2 * T_),
                "mdct_to_raw": leg(t_i, (coef.numel() + back.numel()) * 4, 2.0 * S_ * coef.shape[1] // 2 * coef.shape[2] * N2 * T_),
                "mel_spec_to_mdct_psd": leg(t_p, (melm.numel() + psd.numel()) * 4,
                                            2.0 * S_ * psd.shape[2] * melm.shape[2] * melm.shape[3])}
    except Exception as exc:              # an auxiliary leg must never cost the bench line
        mdct = {"error": repr(exc)}
    return {"metric": "mel-STFT+FGLA samples/sec (batch 64 stereo 45 s @ 32 kHz, 200 FGLA iterations, fp32)",
            "mdct_side": mdct,
            "encode": {"value": B / t_enc, "unit": "stereo samples/s", "ms": t_enc * 1e3,
                       "roofline": {"bound": "hbm", "achieved": enc_bytes / t_enc / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                    "frac": enc_bytes / t_enc / 1e9 / pk["hbm_gbs"]}},
            "fgla_decode": {"value": B / t_dec, "unit": "stereo samples/s", "ms": t_dec * 1e3, "ms_per_iter": t_dec * 1e3 / iters,
                            "roofline": {"bound": "hbm", "achieved": dec_bytes / t_dec / 1e9, "peak": pk["hbm_gbs"],
                                         "unit": "GB/s", "frac": dec_bytes / t_dec / 1e9 / pk["hbm_gbs"]}},
            "encode_plus_decode": {"value": B / (t_enc + t_dec), "unit": "stereo samples/s"}}


def bench_train(device, dist, world: int, steps: int = 6, warmup: int = 3, fused_optimizer: bool = False,
                micro_steps: int = 1) -> dict:
    """BASELINE.json configs[3]: UNet train step forward + backward (bf16 tensor-core compute, fp32 master parameters and
    gradients), device batch 4 of the 45 s latent per GPU, gradients all-reduced (mean) over NCCL overlapped with the
    backward when world > 1.  Loss = unet_trainer.py:259-280.  With `fused_optimizer` the step also runs the optimizer-side
    sweep (clip_grad_norm_ + AdamW + 2 EMA copies + normalize_weights through FusedAdamW),
    i.e. everything trainer.py:1001-1108 does per optimizer step."""
    import torch.nn.functional as F
    from oracle import unet_oracle as uo          # seeded synthetic weights only
    from dualdiffusion_b200.ddp import GradAllReducer
    from dualdiffusion_b200.modules.unets.unet_edm2_b4 import UNet, UNetConfig
    from dualdiffusion_b200 import ops
    spec = uo.default_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    cfg = UNetConfig(**{k: getattr(spec, k) for k in UNetConfig.__dataclass_fields__ if hasattr(spec, k)})
    net = UNet(cfg)
    net.load_state_dict(sd, strict=True)
    net = net.to(device).train()
    net.grad_sync = GradAllReducer()
    B = 4
    g = torch.Generator(device=device).manual_seed(100 + (dist.get_rank() if dist is not None else 0))
    samples = torch.randn(B, *LATENT[1:], device=device, generator=g)
    noise = torch.randn(B, *LATENT[1:], device=device, generator=g)
    sigma = torch.exp(torch.randn(B, device=device, generator=g) * 1.2 - 0.4)
    clap = torch.randn(B, spec.in_channels_emb, device=device, generator=g)
    mask = torch.ones(B, device=device, dtype=torch.bool)
    sig = sigma.view(-1, 1, 1, 1)
    w = (sig ** 2 + spec.sigma_data ** 2) / (sig * spec.sigma_data) ** 2
    opt = None
    if fused_optimizer:
        from dualdiffusion_b200.training.optim import FusedAdamW
        plist = list(net.parameters())
        opt = FusedAdamW(plist, lr=1e-4, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.0)
        opt.attach_module(net)
        opt.attach_emas([[p.detach().clone() for p in plist] for _ in range(2)], [0.9999, 0.99999], [0.9999, None])

    import contextlib

    def step():
        net.zero_grad(set_to_none=True)
        # an optimizer step changes every weight: the per-step weight preparation (weight-norm inside the forward,
        # mp_tools.py:359-364) and the dgrad transposes are part of the train step
        if net._plan is not None and getattr(net._plan, "train_state", None) is not None:
            net._plan.train_state._prep_sig = None
        for m in range(micro_steps):       # gradient accumulation (trainer.py:1022-1044): exchange on the last micro-step only
            ctx = net.grad_sync.no_sync() if m + 1 < micro_steps else contextlib.nullcontext()
            with ctx:
                emb = net.get_embeddings(clap, mask)
                denoised = net(samples + noise * sig, sigma, None, emb)
                wl = (F.mse_loss(denoised, samples, reduction="none") * w).mean(dim=(1, 2, 3))
                logvar = net.get_sigma_loss_logvar(sigma)
                loss = (wl / logvar.exp() + logvar).mean() / micro_steps
                loss.backward()
        if opt is not None:
            opt.clip_grad_norm_(10.0)
            opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    l0 = ops.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ms = max_over_ranks(ms, dist, device)      # job time = slowest rank
    sps = world * B * micro_steps * steps / (ms * 1e-3)
    pk = peaks()
    flop = 3 * FLOP_PER_STEP / 4              # fwd + dgrad + wgrad of one sample-forward (0.489 TFLOP)
    gn = float(torch.sqrt(sum((p.grad.float() ** 2).sum() for p in net.parameters() if p.grad is not None)))
    what = "fwd+bwd+clip+AdamW+2 EMA+normalize_weights" if fused_optimizer else "fwd+bwd"
    if micro_steps > 1:
        what += f", {micro_steps} accumulation micro-steps per optimizer step"
    return {"metric": f"UNet train step {what} samples/sec (device batch 4 x 4x32x688, bf16 compute, fp32 grads)",
            "value": sps, "unit": "samples/s", "ms_per_step": ms / steps, "n_gpus": world,
            "global_batch": world * B * micro_steps,
            "allreduce_bytes_per_step": net.grad_sync.bytes_reduced // max(1, steps + warmup),
            "gpu_launches_per_step": (ops.launch_count - l0) // steps, "loss": float(loss.detach()), "grad_norm": gn,
            "roofline": {"bound": "tensor", "achieved": flop * sps / world / 1e12, "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": flop * sps / world / 1e12 / pk["tflops"]}}


def bench_dae(device, batch: int = 16, steps: int = 3) -> dict:
    """BASELINE.json configs[4]: DAE_D3 diffusion-decoder forward, batch 16 latents (16,8,32,688) -> mel-spectrograms
    (16,2,256,5504), bf16 tensor-core compute, default edm2_ddec_mclt_b1a decoder (random-init weights).  Tensor-bound
    (7.313 TFLOP per sample)."""
    from oracle import dae_oracle as do            # seeded synthetic weights only
    from dualdiffusion_b200.modules.daes.dae_edm2_d3 import DAE_D3, DAE_D3_Config
    from dualdiffusion_b200 import ops
    spec = do.DAESpec()
    sd = do.synth_dae_state_dict(spec, seed=0)
    net = DAE_D3(DAE_D3_Config(channel_mult_enc=spec.channel_mult_enc))
    net.load_state_dict(sd, strict=True)
    net = net.requires_grad_(False).train(False).to(device)
    g = torch.Generator(device=device).manual_seed(0)
    lat = torch.randn(batch, 8, LATENT[2], LATENT[3], device=device, generator=g)
    lat = lat / lat.square().mean(dim=(1, 2, 3), keepdim=True).sqrt()
    emb = net.get_embeddings(torch.randn(batch, spec.in_channels_emb, device=device, generator=g))
    # FLOPs from the layer shapes (2 * pixels * 2 stereo sides * Cout * fan_in per MPConv3D)
    flop = 0.0
    h, w = LATENT[2], LATENT[3]
    flop += 2.0 * h * w * 2 * spec.dec_channels[-1] * 5 * 18
    for name, cin, cout, up in do.dec_block_plan(spec):
        if up:
            h, w = 2 * h, 2 * w
        m = spec.mlp_multiplier
        flop += 2.0 * h * w * 2 * (cout * m * cin * 18 + cout * cout * m * 18 + (cout * cin if cin != cout else 0))
    flop += 2.0 * h * w * 2 * spec.dec_channels[0] * 25
    for _ in range(2):
        mel = net.decode(lat, emb)
    torch.cuda.synchronize()
    l0 = ops.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        mel = net.decode(lat, emb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    pk = peaks()
    sps = batch / (ms * 1e-3)
    return {"metric": "DAE_D3 decoder forward samples/sec (batch 16 latents 8x32x688 -> mel 2x256x5504, bf16)",
            "value": sps, "unit": "samples/s", "ms_per_batch": ms, "batch": batch, "out_shape": list(mel.shape),
            "gpu_launches_per_batch": (ops.launch_count - l0) // steps, "tflop_per_sample": flop / 1e12,
            "roofline": {"bound": "tensor", "achieved": flop * sps / 1e12, "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": flop * sps / 1e12 / pk["tflops"]}}


def bench_ddec(device, batch: int = 4, steps: int = 3) -> dict:
    """Row A17: one DDec_MCLT_UNet_B1 denoiser forward at the 45 s shape, x_in (B,2,256,5504) + PSD x_ref (B,2,4096,5504),
    default edm2_ddec_mclt_b1a configuration (random-init weights), bf16 tensor-core compute."""
    from oracle import ddec_oracle as dd            # seeded synthetic weights + block plan only
    from dualdiffusion_b200.modules.unets.unet_edm2_ddec_mclt_b1 import DDec_MCLT_UNet_B1, DDec_MCLT_UNet_B1_Config
    from dualdiffusion_b200 import ops
    spec = dd.DDecSpec()
    net = DDec_MCLT_UNet_B1(DDec_MCLT_UNet_B1_Config(mlp_multiplier=spec.mlp_multiplier))
    net.load_state_dict(dd.synth_ddec_state_dict(spec, seed=0), strict=True)
    net = net.requires_grad_(False).train(False).to(device)
    H, W = spec.in_num_freqs, LATENT[3] * 8
    g = torch.Generator(device=device).manual_seed(0)
    x = torch.randn(batch, 2, H, W, device=device, generator=g)
    xr = torch.rand(batch, 2, spec.in_psd_freqs, W, device=device, generator=g)
    sigma = torch.full((batch,), 1.5, device=device)
    enc, dec, _ = dd.ddec_block_plan(spec)
    flop, h, w, m = 0.0, H, W, spec.mlp_multiplier
    for name, kind, cin, cout, resample, _ in enc + dec:
        if resample == "down":
            h, w = h // 2, w // 2
        if resample == "up":
            h, w = h * 2, w * 2
        if kind == "conv":
            flop += 2.0 * h * w * 2 * cout * cin * 18
        else:
            c0 = cout if name.startswith("enc") else cin
            flop += 2.0 * h * w * 2 * (cout * m * c0 * 9 + cout * cout * m * 9 + cout * cin * 2)
    flop += 2.0 * H * W * 2 * spec.cblock[0] * 18
    for _ in range(2):
        d = net(x, sigma, None, None, xr)
    torch.cuda.synchronize()
    l0 = ops.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        d = net(x, sigma, None, None, xr)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    pk = peaks()
    sps = batch / (ms * 1e-3)
    return {"metric": "DDec_MCLT_UNet_B1 forward samples/sec (batch 4, x_in 2x256x5504 + PSD 2x4096x5504, bf16)",
            "value": sps, "unit": "samples/s", "ms_per_batch": ms, "batch": batch, "out_shape": list(d.shape),
            "gpu_launches_per_batch": (ops.launch_count - l0) // steps, "tflop_per_sample": flop / 1e12,
            "roofline": {"bound": "tensor", "achieved": flop * sps / 1e12, "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": flop * sps / 1e12 / pk["tflops"]}}


def bench_optim(device, steps: int = 10, warmup: int = 3) -> dict:
    """the optimizer-side sweep of one train step over the default UNet's 293 M fp32 parameters --
    clip_grad_norm_ + AdamW + 2 EMA copies (one feeding back, config/models/default/unet_train.json) + normalize_weights --
    as dd_grad_norm_clip + dd_optim_step_batched (3 launches).  HBM-bound; algorithmic bytes per parameter: 4 (norm) +
    20 read + 12 written (p, g, m, v) + 8 per EMA copy = 52.  `unfused_torch` times the reference's own sequence of library
    calls on the same GPU (clip_grad_norm_, torch AdamW fused=True, _foreach_lerp_ x3, per-tensor normalize + copy_)."""
    from oracle import unet_oracle as uo
    from dualdiffusion_b200.modules.unets.unet_edm2_b4 import UNet, UNetConfig
    from dualdiffusion_b200.modules.mp_tools import MPConv
    from dualdiffusion_b200.training.optim import FusedAdamW
    from dualdiffusion_b200 import ops
    spec = uo.default_spec()
    cfg = UNetConfig(**{k: getattr(spec, k) for k in UNetConfig.__dataclass_fields__ if hasattr(spec, k)})
    net = UNet(cfg).to(device).train()
    params = list(net.parameters())
    n_params = sum(p.numel() for p in params)
    g = torch.Generator(device=device).manual_seed(7)
    for p in params:
        p.grad = torch.randn(p.shape, device=device, generator=g)
    emas = [[p.detach().clone() for p in params] for _ in range(2)]
    betas, fbs = [0.9999, 0.99999], [0.9999, None]
    opt = FusedAdamW(params, lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.0)
    opt.attach_module(net)
    opt.attach_emas(emas, betas, fbs)

    def timed(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def ours():
        opt.clip_grad_norm_(10.0)
        opt.step()

    l0 = ops.launch_count
    ms = timed(ours)
    launches = (ops.launch_count - l0) // (steps + warmup)
    bytes_per_param = 4 + 32 + 8 * len(emas)
    gbs = n_params * bytes_per_param / (ms * 1e-3) / 1e9
    pk = peaks()
    out = {"metric": "optimizer-side sweep (clip + AdamW + 2 EMA + normalize_weights) over the default UNet's parameters",
           "value": 1e3 / ms, "unit": "sweeps/s", "ms": ms, "params": n_params, "gpu_launches_per_step": launches,
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                        "bytes_per_param": bytes_per_param}}
    try:
        convs = [m for m in net.modules() if isinstance(m, MPConv) and not m.disable_weight_norm]
        ref_opt = torch.optim.AdamW(params, lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.0, fused=True)

        def unfused():
            with torch.no_grad():
                torch.nn.utils.clip_grad_norm_(params, 10.0)
                ref_opt.step()
                for e, b, fb in zip(emas, betas, fbs):
                    torch._foreach_lerp_(e, params, 1 - b)
                    if fb is not None:
                        torch._foreach_lerp_(params, e, 1 - fb)
                for c in convs:                                   # mp_tools.py:42-49,375-378 in library calls
                    w = c.weight
                    n = torch.linalg.vector_norm(w, dim=list(range(1, w.ndim)), keepdim=True)
                    w.copy_(w / (1e-4 + n * (n.numel() / w.numel()) ** 0.5))
        ms_ref = timed(unfused)
        out["unfused_torch"] = {"ms": ms_ref, "speedup": ms_ref / ms}
    except Exception as exc:
        out["unfused_torch"] = {"error": repr(exc)}
    return out


def bench_gpu_eager(device, steps: int = 10, warmup: int = 3) -> dict:
    """SURVEY 8(d) / BASELINE.md section 4: the reference's OWN eager PyTorch path on the same B200 -- its UNet from
    baseline/_ref/src, bf16 parameters, channels_last, cudnn.benchmark and TF32 on (init_cuda,
    src/utils/dual_diffusion_utils.py:67-82) -- driven like one sampler step (2 UNet calls at batch 2 + the CFG / Heun
    lerps of pipeline.py:699-725).  cuDNN / cuBLAS / SDPA library kernels; none of this repository's code runs."""
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    net, fmt, spec = _reference_unet(device, torch.bfloat16, torch.channels_last)
    g = torch.Generator().manual_seed(1)
    x1 = torch.randn(LATENT, generator=g).to(device)
    clap = torch.randn(1, spec.in_channels_emb, generator=g).to(device)
    out = {"what": "reference UNet (unmodified, eager PyTorch: cuDNN/cuBLAS/SDPA), bf16 + channels_last + cudnn.benchmark + TF32"}
    with torch.inference_mode():
        emb2 = net.get_embeddings(clap, torch.tensor([True, False], device=device))
        s_a, s_b = torch.tensor([3.0, 3.0], device=device), torch.tensor([2.7, 2.7], device=device)

        def one():
            d = net(x1.repeat(2, 1, 1, 1), s_a, fmt, emb2).float()
            cfg = d[1:].lerp(d[:1], 1.5)
            d2 = net(torch.lerp(cfg, x1, 0.9).repeat(2, 1, 1, 1), s_b, fmt, emb2).float()
            return torch.lerp(cfg, d2[1:].lerp(d2[:1], 1.5), 0.5)
        for _ in range(warmup):
            one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    out["sampler_step"] = {"value": 1e3 / ms, "unit": UNIT, "ms_per_step": ms}
    del net
    torch.cuda.empty_cache()
    return out


def bench_gpu_eager_train(device, steps: int = 4, warmup: int = 2) -> dict:
    """Same for BASELINE config 4 at one GPU: the reference UNet's train step (fwd + bwd) in bf16 (module.half(), i.e.
    bf16 parameters -- the reference's MPConv casts its weights to the activation dtype itself, mp_tools.py:364, so
    autocast with fp32 parameters is not how its modules run), channels_last, device batch 4, loss of
    unet_trainer.py:259-280.  No fp32 master copy is kept here, which only favours the comparator."""
    import torch.nn.functional as F
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    net, fmt, spec = _reference_unet(device, torch.bfloat16, torch.channels_last)
    net = net.requires_grad_(True).train()
    B = 4
    g = torch.Generator(device=device).manual_seed(100)
    samples = torch.randn(B, *LATENT[1:], device=device, generator=g)
    noise = torch.randn(B, *LATENT[1:], device=device, generator=g)
    sigma = torch.exp(torch.randn(B, device=device, generator=g) * 1.2 - 0.4)
    clap = torch.randn(B, spec.in_channels_emb, device=device, generator=g)
    mask = torch.ones(B, device=device, dtype=torch.bool)
    sig = sigma.view(-1, 1, 1, 1)
    w = (sig ** 2 + spec.sigma_data ** 2) / (sig * spec.sigma_data) ** 2

    def step():
        net.zero_grad(set_to_none=True)
        emb = net.get_embeddings(clap, mask)
        denoised = net(samples + noise * sig, sigma, fmt, emb)
        wl = (F.mse_loss(denoised.float(), samples, reduction="none") * w).mean(dim=(1, 2, 3))
        logvar = net.get_sigma_loss_logvar(sigma).float()
        loss = (wl / logvar.exp() + logvar).mean()
        loss.backward()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del net
    torch.cuda.empty_cache()
    return {"value": B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "device_batch": B}


def run_ours(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    # stdout carries exactly ONE line (the JSON): native libraries write banners to fd 1 (NCCL prints its version there),
    # so fd 1 is pointed at stderr for the duration of the run and the JSON line goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=device)
        dist = dist_mod
    from dualdiffusion_b200 import ops, _lib

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:                                                   # the reference's own CPU path when its tree is staged,
            cpu_base = cpu_reference_step_rate(1, 1, budget_s=25.0)
        except Exception:                                      # else the validated CPU port
            cpu_base = cpu_oracle_step_rate(1, 1, budget_s=25.0)
        cpu_base = {k: cpu_base[k] for k in ("value", "unit", "cores", "kind", "sample")}

    K, W = args.steps, max(3, args.warmup)
    with torch.inference_mode():
        net, pipe, state, gen = build_sampler(device, K + W)
        shape = tuple(state.sample.shape)
        noise = [torch.randn(shape, generator=gen, device=device) for _ in range(4)]
        n_sched = state.params.num_steps - 1

        def step(i: int) -> None:                       # cycle through the 100-step schedule
            state.step(i % n_sched, noise[i % 4])

        for i in range(W):
            step(i)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = ops.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(K):
            step(W + i)
        e1.record()
        torch.cuda.synchronize()
        launches = ops.launch_count - launches0
        ms = e0.elapsed_time(e1)
        ms = max_over_ranks(ms, dist, device)          # job time = slowest rank
        if dist is not None:
            dist.barrier()
        value = world * K / (ms * 1e-3)

        # ---- e2e: BASELINE config 2 as a user runs it -- ONE 100-step generate through the public API
        # (DualDiffusionPipeline.diffusion_decode, the reference's sampler entry point, pipeline.py:589-752) with HOST
        # tensors in and out: the conditioning embedding is copied in from pinned memory, the finished sample is copied
        # back, the per-step noise is drawn from the seeded device generator inside the call as in the reference.
        from dualdiffusion_b200.pipelines.dual_diffusion_pipeline import SampleParams as _SP
        gen_params = _SP(seed=4321, num_steps=100, batch_size=1, cfg_scale=1.5, use_heun=True)
        clap_host = torch.randn(1, net.config.in_channels_emb).pin_memory()
        out_host = torch.empty(LATENT, dtype=torch.float32).pin_memory()
        pipe.diffusion_decode(gen_params, quiet=True, audio_embedding=clap_host.to(device, non_blocking=True), sample_shape=LATENT)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        n_gen = max(1, min(3, K // 10))
        t0 = time.perf_counter()
        for _ in range(n_gen):
            res = pipe.diffusion_decode(gen_params, quiet=True, audio_embedding=clap_host.to(device, non_blocking=True),
                                        sample_shape=LATENT)
            out_host.copy_(res, non_blocking=True)
            torch.cuda.synchronize()
        gen_s = (time.perf_counter() - t0) / n_gen
        gen_s = max_over_ranks(gen_s, dist, device)      # job time = slowest rank
        e2e_generate = {"value": world * gen_params.num_steps / gen_s, "unit": UNIT, "seconds_per_100_step_generate": gen_s,
                        "calls_timed": n_gen, "h2d_bytes_per_step": clap_host.numel() * 4 / gen_params.num_steps,
                        "d2h_bytes_per_step": out_host.numel() * 4 / gen_params.num_steps,
                        "api": "DualDiffusionPipeline.diffusion_decode(SampleParams(num_steps=100, use_heun=True, cfg_scale=1.5))"}

        # ---- and the single step driven with HOST buffers every step (pinned), copies inside the timed region
        host_in = torch.empty(shape, dtype=torch.float32).pin_memory()
        host_out = torch.empty(shape, dtype=torch.float32).pin_memory()
        host_in.copy_(state.sample.cpu())
        for i in range(2):
            state.set_sample(host_in); step(i); host_out.copy_(state.sample, non_blocking=True); torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(K):
            state.set_sample(host_in)                                   # H2D from pinned memory
            step(W + i)
            host_out.copy_(state.sample, non_blocking=True)             # D2H of the step's result
            torch.cuda.synchronize()
            host_in, host_out = host_out, host_in
        e2e_s = time.perf_counter() - t0
        e2e_s = max_over_ranks(e2e_s, dist, device)      # job time = slowest rank
        if rank == 0:
            sampler.stop_flag.set()
            sampler.join(timeout=2)
        nbytes = host_in.numel() * 4

        # ---- roofline of the dominant kernel.  The launch list (shapes + fused epilogues) comes from one eager UNet
        # evaluation; every distinct layer of the dominant kernel is then timed IN A CUDA GRAPH (the regime of the
        # captured UNet: no host launch cost between kernels, programmatic dependent launch) cycling through enough
        # buffer sets that its inputs never sit in the 126 MB L2 -- CUDA events around the replays.
        roof = None
        if rank == 0:
            pk = peaks()
            net.use_cuda_graphs = False
            x2 = state.sample2.clone()
            net(x2, state.table[3, 0], None, state.emb)
            torch.cuda.synchronize()
            ops.timing = []
            net(x2, state.table[3, 0], None, state.emb)
            torch.cuda.synchronize()
            rec, ops.timing = ops.timing, None
            net.use_cuda_graphs = True
            is_dx = lambda d: d[5] == 3 and d[6] > 1                        # every grouped 3x3 layer runs on conv3x3_dx_kernel
            eager = [(f, a.elapsed_time(b) * 1e-3, d) for f, a, b, d in rec]
            allc = [(f, t) for f, t, d in eager]
            fl_all, tt_all = sum(f for f, _ in allc), sum(t for _, t in allc)
            shapes = {}
            for f, t, d in eager:
                if is_dx(d):
                    shapes.setdefault(d, [f, 0])[1] += 1

            def layer_time(d) -> float:
                B_, H_, W_, Ci, Co, k_, g_, epi_, epi2_ = d
                per_set = 2 * B_ * H_ * W_ * (Ci + Co * (1 + (epi_ == 2) + (epi2_ != 0)))
                nset = max(2, min(8, -(-320_000_000 // per_set)))
                wp = ops.weight_prep(torch.randn(Co, Ci // g_, k_, k_, device=device))
                sets = []
                for _ in range(nset):
                    kw = {}
                    if epi_ == 1:
                        kw = dict(epi=1, scale=torch.ones(B_, Co, device=device))
                    elif epi_ == 2:
                        kw = dict(epi=2, alpha=0.7, beta=0.3, clip=256.0,
                                  residual=torch.randn(B_, H_, W_, Co, device=device).to(torch.bfloat16))
                    if epi2_ != 0:
                        kw.update(epi2=epi2_, out2=torch.empty(B_, H_, W_, Co, device=device, dtype=torch.bfloat16))
                        if epi2_ == 2:
                            kw["scale2"] = torch.ones(B_, Co, device=device)
                    sets.append((torch.randn(B_, H_, W_, Ci, device=device).to(torch.bfloat16),
                                 torch.empty(B_, H_, W_, Co, device=device, dtype=torch.bfloat16), kw))
                for xs, os_, kw in sets[:2]:
                    ops.mpconv(xs, wp, k_, g_, out=os_, **kw)
                side = torch.cuda.Stream(device=device)
                side.wait_stream(torch.cuda.current_stream(device))
                with torch.cuda.stream(side):
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=side):
                        for rep_ in range(2):
                            for xs, os_, kw in sets:
                                ops.mpconv(xs, wp, k_, g_, out=os_, **kw)
                torch.cuda.current_stream(device).wait_stream(side)
                graph.replay()
                torch.cuda.synchronize()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for _ in range(3):
                    graph.replay()
                g1.record()
                torch.cuda.synchronize()
                return g0.elapsed_time(g1) * 1e-3 / (3 * 2 * nset)

            fl = tt = fl_big = tt_big = 0.0
            n_dx = n_big = 0
            per_layer = []
            for d, (f, cnt) in shapes.items():
                t = layer_time(d)
                fl += f * cnt
                tt += t * cnt
                n_dx += cnt
                if d[1] >= 8:              # levels 0-2 (>= 8 image rows): the launches large enough to be tensor-bound
                    fl_big += f * cnt
                    tt_big += t * cnt
                    n_big += cnt
                per_layer.append({"shape": list(d), "launches_per_unet_call": cnt, "us": t * 1e6, "tflops": f / t / 1e12})
            torch.cuda.empty_cache()
            tt_dx_eager = sum(t for f, t, d in eager if is_dx(d))
            roof = {"bound": "tensor", "kernel": "conv3x3_dx_kernel (tcgen05 tap-stacked implicit-GEMM MPConv, every grouped 3x3 layer of the UNet, levels 0-4)",
                    "achieved": fl / tt / 1e12, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": fl / tt / 1e12 / pk["tflops"],
                    "how": "per distinct layer: CUDA events around replays of a CUDA graph of that layer's launches cycling "
                           "through >= 320 MB of input / output buffer sets (L2-cold inputs); weighted by launches per UNet call",
                    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel from the committed
                    # `ncu --set full` capture: no DRAM re-reads of the activations
                    "traffic": 75.76e6, "traffic_source": "profiles/r02_ncu_dx_l0_summary.csv (ncu --set full, second launch: the "
                                                          "2x32x688 512->256 residual layer, 70.77 MB read + 4.99 MB written "
                                                          "inside the capture window; algorithmic 45.1 MB in + 22.5 MB residual "
                                                          "+ 22.5 MB out = 90.2 MB: the output and part of the residual stay in "
                                                          "the 126 MB L2, no DRAM re-reads)",
                    "launches": n_dx, "avg_launch_us": tt / max(1, n_dx) * 1e6,
                    "flop_per_launch_avg": fl / max(1, n_dx), "peak_source": pk["source"] + " (bf16 sustained)",
                    "share_of_unet_call": tt / (ms * 1e-3 / K / 2),
                    "eager_event_timing": {"avg_launch_us": tt_dx_eager / max(1, n_dx) * 1e6,
                                           "note": "same launches timed one by one with events in an eager pass (host-bound launch gaps included)"},
                    "per_layer": per_layer,
                    # the same kernel on levels 0-2 only (what rounds 1 and 2a reported as `frac`: 0.28, 0.44); the launches of
                    # levels 3-4 (4x86 / 2x43 pixels, 7-20 us each) are latency-bound and pull the all-launch average down
                    "levels_0_2": {"launches": n_big, "achieved": fl_big / max(tt_big, 1e-12) / 1e12,
                                   "frac": fl_big / max(tt_big, 1e-12) / 1e12 / pk["tflops"],
                                   "avg_launch_us": tt_big / max(1, n_big) * 1e6},
                    "all_mpconv": {"achieved_eager": fl_all / tt_all / 1e12, "launches": len(allc)},
                    "step": {"achieved": FLOP_PER_STEP * value / world / 1e12, "frac": FLOP_PER_STEP * value / world / 1e12 / pk["tflops"]}}
            halo_rec = [(f, a, b, d) for f, a, b, d in rec if is_dx(d)]

            try:
                # The same launches against both rooflines: algorithmic HBM bytes of a launch = activations in + every
                # output tensor + the residual operand + the prepared weights (each exactly once).  At level 0 the
                # fused-epilogue layers move more bytes per FLOP than the tensor/HBM balance point, so the binding
                # roofline of a launch is max(flop / tensor peak, bytes / HBM peak).
                def conv_bytes(d):
                    B_, H_, W_, Ci, Co, k_, g_, epi_, epi2_ = d
                    outs = 1 + (1 if epi_ == 2 else 0) + (1 if epi2_ != 0 else 0)
                    return 2.0 * B_ * H_ * W_ * (Ci + Co * outs) + 2.0 * Co * (Ci // g_) * k_ * k_
                hb = [(f, conv_bytes(d), 0.0) for f, a, b, d in halo_rec]
                by = sum(x[1] for x in hb)
                t_bound = sum(max(f / (pk["tflops"] * 1e12), nb / (pk["hbm_gbs"] * 1e9)) for f, nb, _ in hb)
                roof["hbm_view"] = {"algorithmic_bytes_per_launch_avg": by / max(1, len(hb)),
                                    "achieved_gbs": by / tt / 1e9, "peak_gbs": pk["hbm_gbs"],
                                    "frac": by / tt / 1e9 / pk["hbm_gbs"],
                                    "frac_of_binding_roofline": t_bound / tt}
            except Exception as exc:      # an auxiliary view must never cost the bench line
                roof["hbm_view"] = {"error": repr(exc)}

    train = None
    if not args.no_train:
        del net, pipe, state
        torch.cuda.empty_cache()
        train = bench_train(device, dist, world)
        # The complete optimizer step (clip + AdamW + 2 EMA + normalize_weights through FusedAdamW) at EVERY N: the sweep is
        # replicated and deterministic, so it adds no exchange.  At N = 1 also the survey's strong-scaling base: global batch
        # 32 on one GPU as 8 accumulation micro-steps (no_sync) per optimizer step.
        try:
            torch.cuda.empty_cache()
            train["with_optimizer"] = bench_train(device, dist, world, steps=4, fused_optimizer=True)
        except Exception as exc:           # must never cost the bench line
            train["with_optimizer"] = {"error": repr(exc)}
        if world == 1:
            try:
                torch.cuda.empty_cache()
                train["global_batch_32_on_one_gpu"] = bench_train(device, dist, world, steps=2, warmup=1,
                                                                  fused_optimizer=True, micro_steps=8)
            except Exception as exc:
                train["global_batch_32_on_one_gpu"] = {"error": repr(exc)}

    dae = ddec = None
    if rank == 0 and world == 1 and not args.no_dae:
        torch.cuda.empty_cache()
        dae = bench_dae(device)
        torch.cuda.empty_cache()
        ddec = bench_ddec(device)
        torch.cuda.empty_cache()

    secondary = None
    if rank == 0 and world == 1 and not args.no_format:
        secondary = bench_format(device)

    gpu_eager = None
    if rank == 0 and world == 1 and not args.no_gpu_eager:
        try:
            torch.cuda.empty_cache()
            gpu_eager = bench_gpu_eager(device)
            gpu_eager["sampler_step"]["ours_over_eager"] = value / gpu_eager["sampler_step"]["value"]
            if train is not None:
                gpu_eager["train_step"] = bench_gpu_eager_train(device)
                gpu_eager["train_step"]["ours_over_eager"] = train["value"] / gpu_eager["train_step"]["value"]
        except Exception as exc:
            gpu_eager = {**(gpu_eager or {}), "unavailable": repr(exc)}

    optim = None
    if rank == 0 and world == 1 and not args.no_train:
        # last on purpose: every other number is already on the host when this leg runs, and it must never cost the line
        try:
            torch.cuda.empty_cache()
            optim = bench_optim(device)
        except Exception as exc:
            optim = {"error": repr(exc)}
        optim["train_step_with_optimizer"] = (train or {}).get("with_optimizer")      # measured with the train legs above

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": CONFIG,
                "config_details": {"sampler_batch": 1, "unet_batch": 2,
                                   "parallelism": f"replicas x{world} (sampler is batch-sharded, no collective)",
                                   "l2_policy": "per-step working set (585 MB bf16 weights + activations) exceeds the 126 MB L2; no flush",
                                   "library": os.path.relpath(_lib.lib_path(), ROOT)},
                "e2e": e2e_generate,
                "e2e_per_step_roundtrip": {"value": world * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": nbytes,
                                           "d2h_bytes_per_step": nbytes,
                                           "what": "EDMSamplerState.step with the sample copied in from / out to pinned host memory every step"},
                "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu_base,
                "gpu_eager": gpu_eager, "train_step": train, "dae_decode": dae, "ddec_forward": ddec, "secondary": secondary,
                "optim_step": optim}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-format", action="store_true", help="skip the mel-STFT/FGLA secondary measurement")
    ap.add_argument("--no-dae", action="store_true", help="skip the DAE_D3 decoder (BASELINE config 5) measurement")
    ap.add_argument("--no-train", action="store_true", help="skip the train-step (fwd+bwd+all-reduce) measurement")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the reference's eager PyTorch path on the same GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
