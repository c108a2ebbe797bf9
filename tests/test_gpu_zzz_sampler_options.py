"""GPU parity tests of the sampler-loop options (`seamless_loop`, `stereo_fix`; pipelines/dual_diffusion_pipeline.py:
638-656, :729-732) through the C ABI: the index kernels bit for bit against torch.roll / torch.cat, the noise mix against
the reference formula, and `diffusion_decode` with each option against the golden produced by the unmodified reference
(noise draws injected: a CUDA generator cannot reproduce a CPU generator's stream).  Tolerances: bit-exact for index
maps, 1e-6 for the fp32 noise mix, 3e-2 relative L2 for the whole bf16 network (as tests/test_gpu_parity.py).  The golden's
sample is 16 x 48 so that the loop-padded frame (16 x 112) stays inside the attention kernel's 640-token limit at the
reduced UNet's attention level (8 x 56 = 448); the 45 s latent pads to 32 x 752 -> 4 x 94 tokens.

The file name sorts last on purpose: written after the round's GPU budget was spent; not yet run on a GPU."""
import os

import pytest
import torch

from oracle import sampler_oracle, unet_oracle as uo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampler_options_small.pt")
BF16_NET = 3e-2


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("W,pad,shift", [(48, 32, 0), (48, 32, 17), (48, 32, 47), (688, 32, 300), (33, 33, 5), (40, 0, 39)])
def test_roll_pad_and_crop_unroll_are_bit_exact(W, pad, shift):
    from dualdiffusion_b200 import ops
    dev = _dev()
    gen = torch.Generator().manual_seed(W + shift)
    x = torch.randn(2, 3, 4, W, generator=gen)
    r = torch.roll(x, shifts=shift, dims=-1)
    ref = torch.cat((r[..., W - pad:], r, r[..., :pad]), dim=-1).repeat(2, 1, 1, 1)
    out = ops.roll_pad_w(x.to(dev), shift, pad, copies=2)
    assert out.shape == ref.shape and torch.equal(out.cpu(), ref)
    back = ops.crop_unroll_w(out[:2].contiguous(), shift, pad)
    assert torch.equal(back.cpu(), x)
    y = torch.randn(2, 3, 4, W + 2 * pad, generator=gen)
    assert torch.equal(ops.crop_unroll_w(y.to(dev), shift, pad).cpu(), torch.roll(y[..., pad:pad + W], shifts=-shift, dims=-1))


@pytest.mark.parametrize("t", [0.3, 0.5, 0.8])
def test_stereo_fix_noise_vs_reference_formula(t):
    from dualdiffusion_b200 import ops
    dev = _dev()
    gen = torch.Generator().manual_seed(11)
    noise, fresh = torch.randn(2, 4, 32, 48, generator=gen), torch.randn(2, 4, 32, 48, generator=gen)
    ref = noise.clone()
    ref[:, ::2] = ref[:, 1::2]
    ref = uo.mp_sum(fresh, ref, t)
    out = ops.stereo_fix_noise(noise.to(dev), fresh.to(dev), t)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-6, atol=1e-6)


def test_sampler_options_vs_golden_reference():
    from dualdiffusion_b200.modules.unets.unet_edm2_b4 import UNet, UNetConfig
    from dualdiffusion_b200.pipelines.dual_diffusion_pipeline import DualDiffusionPipeline, SampleParams
    dev = _dev()
    spec = uo.small_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    g = torch.load(GOLD, weights_only=False)
    cfg = UNetConfig(**{k: getattr(spec, k) for k in UNetConfig.__dataclass_fields__ if hasattr(spec, k)})
    net = UNet(cfg)
    net.load_state_dict(sd, strict=True)
    net = net.requires_grad_(False).train(False).to(device=dev)
    pipe = DualDiffusionPipeline({"unet": net})
    for name, case in g["cases"].items():
        rec = {}
        x_ref = g["x_ref"] if case["use_ref"] else None
        ref = sampler_oracle.diffusion_decode(sd, spec, g["clap"], tuple(g["shape"]), seed=case["seed"], record=rec,
                                              x_ref=x_ref, stereo_noise=case["stereo_noise"], **case["kwargs"])
        assert rel_err(ref, case["sample"]) < 1e-4
        params = SampleParams(seed=case["seed"], batch_size=1, **case["kwargs"])
        out = pipe.diffusion_decode(params, quiet=True, audio_embedding=g["clap"], sample_shape=tuple(g["shape"]),
                                    x_ref=None if x_ref is None else x_ref.to(dev), initial_noise=rec["initial_noise"],
                                    step_noise=lambda i: rec["step_noise"][i], stereo_noise=case["stereo_noise"])
        assert out.shape == case["sample"].shape
        assert rel_err(out, case["sample"]) < BF16_NET, name


def test_seamless_loop_without_x_ref_raises_like_the_reference():
    from dualdiffusion_b200.pipelines.dual_diffusion_pipeline import DualDiffusionPipeline, SampleParams
    _dev()

    class Stub(torch.nn.Module):
        device = torch.device("cuda:0")
        config = type("C", (), dict(sigma_max=200.0, sigma_min=0.03, sigma_data=1.0))()
    pipe = DualDiffusionPipeline({"unet": Stub()})
    with pytest.raises(ValueError, match="x_ref"):
        pipe.diffusion_decode(SampleParams(seed=1, num_steps=2, seamless_loop=True), audio_embedding=torch.zeros(1, 8),
                              sample_shape=(1, 4, 32, 48))
