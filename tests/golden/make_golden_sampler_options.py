"""Regenerates tests/golden/sampler_options_small.pt from the UNMODIFIED reference (CPU, build container only): the EDM
sampler loop `DualDiffusionPipeline.diffusion_decode` (pipelines/dual_diffusion_pipeline.py:589-752) with its two loop
options -- `seamless_loop` (random circular shift + 32-column circular padding around every step, :651-656, :729-732; the
reference needs an x_ref for it) and `stereo_fix` (:638-640; its un-seeded `torch.randn_like` draw is pinned by seeding
the global RNG right before the call and stored).

    python tests/golden/make_golden_sampler_options.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import ref_shim, unet_oracle as uo  # noqa: E402
from make_golden import build_reference_unet, weight_checksum  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
# 16 rows: with the 2 x 32 loop-padding columns the coarsest level is 8 x 56 = 448 tokens, inside the attention kernel's
# 640-token limit (the 45 s latent pads to 32 x 752 -> 4 x 94 tokens at the first attention level)
SHAPE = (1, 4, 16, 48)
GLOBAL_SEED = 77


def main():
    ref_shim.install()
    from modules.formats.ms_mdct_dual import MS_MDCT_DualFormat, MS_MDCT_DualFormatConfig
    from pipelines.dual_diffusion_pipeline import DualDiffusionPipeline, SampleParams
    fmt = MS_MDCT_DualFormat(MS_MDCT_DualFormatConfig())
    torch.set_grad_enabled(False)
    spec = uo.small_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    net = build_reference_unet(spec, sd)
    pipe = DualDiffusionPipeline({"unet": net, "format": fmt})
    g = torch.Generator().manual_seed(21)
    clap = torch.randn(1, spec.in_channels_emb, generator=g)
    x_ref = torch.rand(SHAPE[0], SHAPE[1] + 1, *SHAPE[2:], generator=g)
    cases = {}
    for name, kw, use_ref in (("loop_heun2", dict(num_steps=2, use_heun=True, seamless_loop=True), True),
                              ("loop_euler3", dict(num_steps=3, use_heun=False, seamless_loop=True, cfg_scale=2.0), True),
                              ("stereo_fix", dict(num_steps=2, use_heun=True, stereo_fix=0.3), False),
                              ("stereo_fix_hi", dict(num_steps=1, use_heun=False, stereo_fix=0.8), False)):
        torch.manual_seed(GLOBAL_SEED)
        params = SampleParams(seed=4321, batch_size=1, **kw)
        out = pipe.diffusion_decode(params, quiet=True, audio_embedding=clap, sample_shape=SHAPE,
                                    x_ref=x_ref if use_ref else None, module=net)
        torch.manual_seed(GLOBAL_SEED)
        fresh = torch.randn(SHAPE)                  # what randn_like(noise) drew from the global RNG
        cases[name] = dict(kwargs=kw, seed=4321, use_ref=use_ref, sample=out.clone(), stereo_noise=fresh)
        print("sampler", name, out.std().item())
    torch.save(dict(clap=clap, x_ref=x_ref, shape=SHAPE, cases=cases, weight_checksum=weight_checksum(sd)),
               os.path.join(OUT, "sampler_options_small.pt"))


if __name__ == "__main__":
    main()
