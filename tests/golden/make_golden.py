"""Regenerates tests/golden/*.pt from the UNMODIFIED reference, imported on CPU in the build container through
oracle/ref_shim.py.  Run:  python tests/golden/make_golden.py   (needs /root/reference; never runs on the GPU box).

Golden files hold inputs + reference outputs only.  Model weights are NOT stored: they are re-created from
`oracle.unet_oracle.synth_state_dict(spec, seed)` (a seeded CPU generator), and a checksum of them is stored so
a silent RNG change is detected.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim, unet_oracle as uo  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def weight_checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def build_reference_unet(spec, sd):
    from modules.unets.unet_edm2_b4 import UNet, UNetConfig
    cfg = UNetConfig(**{k: getattr(spec, k) for k in UNetConfig.__dataclass_fields__ if hasattr(spec, k)})
    net = UNet(cfg).eval()
    net.load_state_dict(sd, strict=True)
    return net


def main():
    ref_shim.install()
    from modules.formats.ms_mdct_dual import MS_MDCT_DualFormat, MS_MDCT_DualFormatConfig
    from pipelines.dual_diffusion_pipeline import DualDiffusionPipeline, SampleParams
    from sampling.schedule import SamplingSchedule
    fmt = MS_MDCT_DualFormat(MS_MDCT_DualFormatConfig())
    torch.set_grad_enabled(False)

    # ---- UNet forward, reduced config (every code path) and the default config at BASELINE config 1 ----
    for tag, spec, shape, sigmas, masks in (
            ("unet_small", uo.small_spec(), (2, 4, 32, 48), [2.0, 0.3], [True, False]),
            ("unet_default_c1", uo.default_spec(), (1, 4, 64, 64), [2.0], [True])):
        sd = uo.synth_state_dict(spec, seed=0)
        net = build_reference_unet(spec, sd)
        g = torch.Generator().manual_seed(1)
        x = torch.randn(shape, generator=g)
        sigma = torch.tensor(sigmas)
        clap = torch.randn(shape[0], spec.in_channels_emb, generator=g)
        mask = torch.tensor(masks)
        emb = net.get_embeddings(clap, mask)
        d = net(x, sigma, fmt, emb)
        x_ref = torch.rand(shape[0], shape[1] + 1, *shape[2:], generator=g)
        d_ref = net(x, sigma, fmt, emb, x_ref)
        logvar = net.get_sigma_loss_logvar(sigma)
        net.train()
        d_train = net(x, sigma, fmt, emb)
        torch.save(dict(x=x, sigma=sigma, clap=clap, mask=mask, emb=emb, d=d, x_ref=x_ref, d_xref=d_ref, logvar=logvar,
                        d_train=d_train, weight_checksum=weight_checksum(sd)), os.path.join(OUT, tag + ".pt"))
        print(tag, "std", d.std().item())

    # ---- train step: gradients of the reference's loss (unet_trainer.py:259-280) through the reference module ----
    torch.set_grad_enabled(True)
    spec = uo.small_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    net = build_reference_unet(spec, sd).train()
    g = torch.Generator().manual_seed(5)
    samples = torch.randn(2, 4, 32, 48, generator=g)
    noise = torch.randn(2, 4, 32, 48, generator=g)
    sigma = torch.tensor([1.5, 0.2])
    clap = torch.randn(2, spec.in_channels_emb, generator=g)
    mask = torch.tensor([True, False])
    emb = net.get_embeddings(clap, mask)
    sig = sigma.view(-1, 1, 1, 1)
    denoised = net(samples + noise * sig, sigma, fmt, emb)
    w = (sig ** 2 + spec.sigma_data ** 2) / (sig * spec.sigma_data) ** 2
    wl = (torch.nn.functional.mse_loss(denoised, samples, reduction="none") * w).mean(dim=(1, 2, 3))
    logvar = net.get_sigma_loss_logvar(sigma)
    loss = (wl / logvar.exp() + logvar).mean()
    loss.backward()
    stats = {}
    for name, p in net.named_parameters():
        stats[name] = (float(p.grad.norm()), float((p.grad * uo.grad_probe(name, p.shape)).sum()))
    small = {n: p.grad.clone() for n, p in net.named_parameters() if p.numel() <= 4096}
    torch.save(dict(samples=samples, noise=noise, sigma=sigma, clap=clap, mask=mask, loss=loss.detach(),
                    denoised=denoised.detach(), grad_stats=stats, small_grads=small,
                    weight_checksum=weight_checksum(sd)), os.path.join(OUT, "unet_small_train.pt"))
    print("train loss", float(loss), "params", len(stats))
    torch.set_grad_enabled(False)

    # ---- DAE_D3 decoder (dae_edm2_d3.py:356-369), reduced config, two latent sizes (coarsest level 4 and 8 rows) ----
    from oracle import dae_oracle as do
    from modules.daes.dae_edm2_d3 import DAE_D3, DAE_D3_Config
    dspec = do.small_dae_spec()
    dsd = do.synth_dae_state_dict(dspec, seed=0)
    dae = DAE_D3(DAE_D3_Config(in_channels_emb=dspec.in_channels_emb, model_channels=dspec.model_channels,
                               channel_mult_enc=dspec.channel_mult_enc, channel_mult_dec=tuple(dspec.channel_mult_dec),
                               channel_mult_emb=dspec.channel_mult_emb, num_enc_layers=dspec.num_enc_layers,
                               num_dec_layers_per_block=dspec.num_dec_layers_per_block,
                               mlp_multiplier=dspec.mlp_multiplier)).eval()
    dae.load_state_dict(dsd, strict=True)
    g = torch.Generator().manual_seed(6)
    cases = {}
    for tag, shape in (("h4", (2, 8, 4, 12)), ("h8", (1, 8, 8, 22))):
        lat = uo.normalize(torch.randn(shape, generator=g))
        emb_in = torch.randn(shape[0], dspec.in_channels_emb, generator=g)
        emb = dae.get_embeddings(emb_in)
        cases[tag] = dict(latents=lat, emb_in=emb_in, emb=emb, mel=dae.decode(lat, emb))
        print("dae", tag, cases[tag]["mel"].shape, float(cases[tag]["mel"].std()))
    mel_in = torch.randn(2, 2, 16, 24, generator=g).abs() * 20                 # encoder (:342-354)
    enc = dict(mel=mel_in, latents=dae.encode(mel_in, None), pre_norm=dae.encode(mel_in, None, training=True))
    torch.save(dict(cases=cases, encode=enc, weight_checksum=weight_checksum(dsd),
                    mel_shape=dae.get_mel_spec_shape((3, 8, 32, 688)),
                    latent_shape=dae.get_latent_shape((3, 2, 256, 5504))), os.path.join(OUT, "dae_small.pt"))

    # ---- diffusion-decoder UNet DDec_MCLT_UNet_B1 (unet_edm2_ddec_mclt_b1.py:278-326), reduced config ----
    from oracle import ddec_oracle as dd
    from modules.unets.unet_edm2_ddec_mclt_b1 import DDec_MCLT_UNet_B1, DDec_MCLT_UNet_B1_Config
    sspec = dd.small_ddec_spec()
    ssd = dd.synth_ddec_state_dict(sspec, seed=0)
    ddec = DDec_MCLT_UNet_B1(DDec_MCLT_UNet_B1_Config(
        in_num_freqs=sspec.in_num_freqs, in_psd_freqs=sspec.in_psd_freqs, model_channels=sspec.model_channels,
        logvar_channels=sspec.logvar_channels, channel_mult=tuple(sspec.channel_mult), double_midblock=sspec.double_midblock,
        channel_mult_noise=sspec.channel_mult_noise, channel_mult_emb=sspec.channel_mult_emb,
        num_layers_per_block=sspec.num_layers_per_block, mlp_multiplier=sspec.mlp_multiplier)).eval()
    ddec.load_state_dict(ssd, strict=True)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 2, sspec.in_num_freqs, 24, generator=g)
    x_ref = torch.rand(2, 2, sspec.in_psd_freqs, 24, generator=g)
    pert = x + 0.1 * torch.randn(x.shape, generator=g)
    sigma = torch.tensor([0.5, 3.0])
    torch.save(dict(x=x, x_ref=x_ref, sigma=sigma, perturbed=pert, d=ddec(x, sigma, None, None, x_ref),
                    d_perturbed=ddec(x, sigma, None, None, x_ref, pert), logvar=ddec.get_sigma_loss_logvar(sigma),
                    weight_checksum=weight_checksum(ssd)), os.path.join(OUT, "ddec_small.pt"))

    # ---- 2-D diffusion-decoder UNet unet_edm2_q4_ddec.UNet (:253-303), reduced config ----
    from modules.unets.unet_edm2_q4_ddec import UNet as Q4UNet, UNet_Config as Q4Config
    qspec = dd.small_q4_spec()
    qsd = dd.synth_q4_state_dict(qspec, seed=0)
    q4 = Q4UNet(Q4Config(in_num_freqs=qspec.in_num_freqs, in_psd_freqs=qspec.in_psd_freqs, model_channels=qspec.model_channels,
                         logvar_channels=qspec.logvar_channels, channel_mult=tuple(qspec.channel_mult),
                         double_midblock=qspec.double_midblock, channel_mult_noise=qspec.channel_mult_noise,
                         channel_mult_emb=qspec.channel_mult_emb, num_layers_per_block=qspec.num_layers_per_block,
                         mlp_multiplier=qspec.mlp_multiplier)).eval()
    q4.load_state_dict(qsd, strict=True)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 2, qspec.in_num_freqs, 24, generator=g)
    x_ref = torch.rand(2, 2, qspec.in_psd_freqs, 24, generator=g)
    sigma = torch.tensor([0.5, 3.0])
    torch.save(dict(x=x, x_ref=x_ref, sigma=sigma, d=q4(x, sigma, None, None, x_ref), weight_checksum=weight_checksum(qsd)),
               os.path.join(OUT, "q4_ddec_small.pt"))

    # ---- sampler: reference diffusion_decode on CPU, reduced config, 3 Heun+CFG steps ----
    spec = uo.small_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    net = build_reference_unet(spec, sd)
    pipe = DualDiffusionPipeline({"unet": net, "format": fmt})
    g = torch.Generator().manual_seed(2)
    clap = torch.randn(1, spec.in_channels_emb, generator=g)
    cases = {}
    for name, kw in (("heun3", dict(num_steps=3, use_heun=True)), ("euler4", dict(num_steps=4, use_heun=False, cfg_scale=2.0))):
        params = SampleParams(seed=1234, batch_size=1, **kw)
        out = pipe.diffusion_decode(params, quiet=True, audio_embedding=clap, sample_shape=(1, 4, 32, 48), module=net)
        cases[name] = dict(kwargs=kw, seed=1234, sample=out.clone())
        print("sampler", name, out.std().item())
    torch.save(dict(clap=clap, cases=cases, weight_checksum=weight_checksum(sd)), os.path.join(OUT, "sampler_small.pt"))

    # ---- mel-STFT encode + FGLA decode (old SpectrogramFormat; two config attributes the current base
    #      config no longer defines are restored) ----
    from modules.formats.old.spectrogram import SpectrogramFormat, SpectrogramFormatConfig
    SpectrogramFormatConfig.sample_raw_channels = 2
    SpectrogramFormatConfig.sample_raw_length = 1440000
    sfmt = SpectrogramFormat(SpectrogramFormatConfig())
    g = torch.Generator().manual_seed(3)
    raw = 0.1 * torch.randn(2, 2, 256 * 95, generator=g)              # 96 frames
    t = torch.arange(raw.shape[-1]) / 32000.0
    raw[0, 0] += 0.3 * torch.sin(2 * torch.pi * 440.0 * t)            # a tonal component in one channel
    raw[0, 1] += 0.2 * torch.sin(2 * torch.pi * 1320.0 * t + 0.5)
    mel = sfmt.raw_to_sample(raw)
    decoded = {n: sfmt.sample_to_raw(mel, n_fgla_iters=n, quiet=True) for n in (1, 2, 5, 12)}
    torch.save(dict(raw=raw, mel=mel, decoded=decoded, shape_1408768=sfmt.get_sample_shape(1, 1408768),
                    crop_1440000=sfmt.sample_raw_crop_width(1440000)), os.path.join(OUT, "format_small.pt"))
    print("format", mel.shape, {k: float(v.std()) for k, v in decoded.items()})

    # ---- live format: MS_MDCT_DualFormat.raw_to_mel_spec (two-window mel-STFT) ----
    g = torch.Generator().manual_seed(4)
    raw = 0.1 * torch.randn(2, 2, 256 * 79, generator=g)              # 80 frames
    raw[1, 0] += 0.25 * torch.sin(2 * torch.pi * 880.0 * torch.arange(raw.shape[-1]) / 32000.0)
    torch.save(dict(raw=raw, mel=fmt.raw_to_mel_spec(raw), mel_shape_default=fmt.get_mel_spec_shape(3),
                    crop_default=fmt.get_raw_crop_width(), unscaled_34=fmt.ms_freq_scale.get_unscaled(34)),
               os.path.join(OUT, "ms_dual_small.pt"))

    # ---- live format, MDCT side (ms_mdct_dual.py:259-318): default config and dual-channel / 4096-bin PSD variant ----
    g = torch.Generator().manual_seed(9)
    raw = 0.1 * torch.randn(1, 2, 256 * 19 + 100, generator=g)                # ragged length: remainder padding path
    raw[0, 1] += 0.2 * torch.sin(2 * torch.pi * 1000.0 * torch.arange(raw.shape[-1]) / 32000.0)
    mdct_cases = {}
    for tag, kw in (("default", {}), ("dual4096", dict(mdct_dual_channel=True, mdct_psd_num_bins=4096))):
        f2 = MS_MDCT_DualFormat(MS_MDCT_DualFormatConfig(**kw))
        mdct = f2.raw_to_mdct(raw)
        mel = f2.raw_to_mel_spec(raw)
        mdct_cases[tag] = dict(kwargs=kw, mdct=mdct, psd=f2.raw_to_mdct_psd(raw), raw_back=f2.mdct_to_raw(mdct),
                               mel=mel, mel_psd=f2.mel_spec_to_mdct_psd(mel), mdct_shape=f2.get_mdct_shape(3))
    torch.save(dict(raw=raw, cases=mdct_cases), os.path.join(OUT, "mdct_small.pt"))

    # ---- sigma schedules ----
    sched = {n: SamplingSchedule.get_schedule(n, 10, 1.0, sigma_max=200.0, sigma_min=0.03, rho=7.0)
             for n in ("edm2", "ln_linear", "linear", "cos", "scale_invariant")}
    sched["edm2_100"] = SamplingSchedule.get_schedule("edm2", 100, 1.0, sigma_max=200.0, sigma_min=0.03, rho=7.0)
    torch.save(sched, os.path.join(OUT, "schedules.pt"))


if __name__ == "__main__":
    main()
