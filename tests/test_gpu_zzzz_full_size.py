"""Parity at BASELINE.json's FULL sizes (moved here from tools/dev_check_*.py): the assembled default UNet at the
45 s latent against the CPU oracle, and -- where the oracle would take hours -- size-independent properties: a batch of
16 DAE decodes / 64 FGLA reconstructions must reproduce the single-item runs item by item (the work units are
independent), which the small-size goldens pin against the reference."""
import pytest
import torch

from conftest import rel_err
from oracle import unet_oracle as uo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def test_default_unet_at_45s_latent_vs_cpu_oracle(dev):
    """BASELINE configs[1] shape: default 293 M UNet, batch 2 (cond | uncond) x 4 x 32 x 688, bf16 tensor-core body against
    the fp32 CPU oracle (itself pinned at 1e-6 against the reference).  Whole-network tolerance 3e-2 (measured 1.0e-2)."""
    from dualdiffusion_b200.modules.unets.unet_edm2_b4 import UNet, UNetConfig
    spec = uo.default_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    cfg = UNetConfig(**{k: getattr(spec, k) for k in UNetConfig.__dataclass_fields__ if hasattr(spec, k)})
    net = UNet(cfg)
    net.load_state_dict(sd, strict=True)
    net = net.requires_grad_(False).train(False).to(device=dev)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 4, 32, 688, generator=g)
    sigma = torch.tensor([3.0, 3.0])
    clap = torch.randn(1, spec.in_channels_emb, generator=g)
    mask = torch.tensor([True, False])
    with torch.inference_mode():
        d_eager = net(x.to(dev), sigma.to(dev), None, net.get_embeddings(clap, mask))
        d_graph = net(x.to(dev), sigma.to(dev), None, net.get_embeddings(clap, mask))      # second call: CUDA-graph replay
        ref = uo.unet_forward(sd, spec, x, sigma, uo.get_embeddings(sd, clap, mask))
    assert torch.equal(d_eager, d_graph)
    assert rel_err(d_graph, ref) < 3e-2
    c_skip = spec.sigma_data ** 2 / (sigma.view(-1, 1, 1, 1) ** 2 + spec.sigma_data ** 2)
    body, body_ref = d_graph.cpu() - c_skip * x, ref - c_skip * x                          # without the skip term
    assert rel_err(body, body_ref) < 3e-2


def test_dae_decode_batch16_reproduces_single_items(dev):
    """BASELINE configs[4]: 16 latents (8 x 32 x 688) -> mel (2 x 256 x 5504).  Items are independent, so the batch result
    must equal the single-item decodes bit for bit (same kernels, same per-output summation order)."""
    from oracle import dae_oracle as do
    from dualdiffusion_b200.modules.daes.dae_edm2_d3 import DAE_D3, DAE_D3_Config
    spec = do.DAESpec()
    net = DAE_D3(DAE_D3_Config(channel_mult_enc=spec.channel_mult_enc))
    net.load_state_dict(do.synth_dae_state_dict(spec, seed=0), strict=True)
    net = net.requires_grad_(False).train(False).to(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    lat = torch.randn(16, 8, 32, 688, device=dev, generator=g)
    lat = lat / lat.square().mean(dim=(1, 2, 3), keepdim=True).sqrt()
    emb = net.get_embeddings(torch.randn(16, spec.in_channels_emb, device=dev, generator=g))
    mel = net.decode(lat, emb)
    assert tuple(mel.shape) == (16, 2, 256, 5504) and bool(torch.isfinite(mel).all())
    for i in (0, 7, 15):
        one = net.decode(lat[i:i + 1].contiguous(), emb[i:i + 1].contiguous())
        assert torch.equal(one[0], mel[i]), i


def test_fgla_batch64_reproduces_single_items(dev):
    """BASELINE configs[2]: 64 stereo 45 s waveforms.  mel-STFT encode and 3 FGLA iterations at the full batch against the
    same calls on single items (stereo pairs are the independent unit, phase_recovery.py:63-64)."""
    from dualdiffusion_b200.modules.formats.spectrogram import SpectrogramFormat, SpectrogramFormatConfig
    fmt = SpectrogramFormat(SpectrogramFormatConfig())
    Ls = fmt.sample_raw_crop_width(1408768)
    g = torch.Generator(device=dev).manual_seed(0)
    raw = 0.1 * torch.randn(64, 2, Ls, device=dev, generator=g)
    mel = fmt.raw_to_sample(raw)
    assert tuple(mel.shape) == (64, 2, 256, 5504)
    wave = fmt.sample_to_raw(mel, n_fgla_iters=3)
    for i in (0, 33, 63):
        mel_i = fmt.raw_to_sample(raw[i:i + 1].contiguous())
        assert rel_err(mel_i[0], mel[i]) < 1e-6, i
        wave_i = fmt.sample_to_raw(mel[i:i + 1].contiguous(), n_fgla_iters=3)
        assert rel_err(wave_i[0], wave[i]) < 1e-5, i
