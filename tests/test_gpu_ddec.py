"""GPU parity tests (-m gpu) of the diffusion-decoder UNet DDec_MCLT_UNet_B1: the input
assembly (x_ref PSD view/permute -- an index map, bit-exact), the padded-layout glue kernels, and the assembled forward
against the reference golden (tests/golden/ddec_small.pt; the reference's own body runs in bf16) and the fp32 oracle."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_err
from oracle import ddec_oracle as dd, unet_oracle as uo
from test_gpu_dae import PW, bf16_round, fold, unfold

pytestmark = pytest.mark.gpu
BF16_OP, BF16_NET = 4e-3, 3e-2


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def test_ddec_stem_xref_permute_is_bit_exact(dev):
    """unet_edm2_ddec_mclt_b1.py:294-309: x_ref.view(B, 2, F, k, W).permute(0, 3, 1, 2, 4) joins c_in*x and the constant
    channel.  Pure data movement (+ one bf16 rounding): compared bit for bit with the reference's own statement."""
    from dualdiffusion_b200 import ops
    gen = torch.Generator().manual_seed(1)
    B, Fq, k, W = 2, 8, 4, 9
    x = torch.randn(B, 2, Fq, W, generator=gen)
    xr = torch.rand(B, 2, Fq * k, W, generator=gen)
    sigma = torch.tensor([0.5, 3.0])
    got = ops.ddec_stem(x.to(dev), xr.to(dev), sigma.to(dev), 1.0, k, PW, 64).float().cpu()
    c_in = 1 / (1 + sigma.view(-1, 1, 1, 1, 1) ** 2).sqrt()
    xr5 = xr.view(B, 2, Fq, k, W).permute(0, 3, 1, 2, 4).to(torch.bfloat16)
    x5 = (c_in * x.reshape(B, 1, 2, Fq, W)).to(torch.bfloat16)
    ref5 = torch.cat((x5, xr5, torch.ones_like(x5[:, :1])), dim=1).float()              # (B, k+2, 2, F, W)
    ref = fold(ref5, "cpu").float()
    ct = k + 2
    for z in range(2):
        assert torch.equal(got[..., z * ct + 1:(z + 1) * ct], ref[..., z * ct + 1:(z + 1) * ct])     # PSD bins + constant: exact
        assert rel_err(got[..., z * ct], ref[..., z * ct]) < 4e-3       # c_in*x: rsqrt vs 1/sqrt may differ by one bf16 ulp
    assert got[..., 2 * (k + 2):].abs().max().item() == 0


def test_padded_avgpool_and_per_side_pixel_norm(dev):
    from dualdiffusion_b200 import ops
    gen = torch.Generator().manual_seed(2)
    x5 = bf16_round(torch.randn(2, 32, 2, 8, 12, generator=gen) * 2)
    y = ops.avgpool2_pad(fold(x5, dev), PW)
    ref = dd.resample_3d(x5, "down")
    assert rel_err(unfold(y, 32), ref) < BF16_OP
    yl = y.float().cpu()
    assert torch.equal(yl[:, :, PW - 1], yl[:, :, PW + 1]) and torch.equal(yl[:, :, -1], yl[:, :, -PW - 3])
    xf = fold(x5, dev)
    xn, s = ops.pixnorm_silu(xf.view(xf.shape[0], xf.shape[1], xf.shape[2] * 2, 32))
    refn = uo.normalize(x5, dim=1)
    assert rel_err(unfold(xn.view(xf.shape), 32), refn) < BF16_OP
    assert rel_err(unfold(s.view(xf.shape), 32), uo.mp_silu(refn)) < BF16_OP
    # (2,1,1) conv_skip: dense 1x1 on the folded layout with the block-circulant weight
    w = torch.randn(64, 32, 2, 1, 1, generator=gen)
    fan = 64.0
    refc = dd.mp_conv3d(x5, bf16_round(w / fan ** 0.5) * fan ** 0.5)
    got = ops.mpconv(xf, ops.weight_prep_z2(w.to(dev)), 1)
    assert rel_err(unfold(got, 64), refc) < BF16_OP
    # folded mp_cat: per-side channel concat through the [.., 2*Wp, C] view
    b5 = bf16_round(torch.randn(2, 64, 2, 8, 12, generator=gen))
    wa, wb = uo.mp_cat_weights(32, 64, 0.5)
    bf = fold(b5, dev)
    xc, _ = ops.cat_silu(xf.view(2, 8, -1, 32), bf.view(2, 8, -1, 64), wa, wb, False)
    refcat = torch.cat([wa * x5, wb * b5], dim=1)
    assert rel_err(unfold(xc.view(2, 8, xf.shape[2], 192), 96), refcat) < BF16_OP


def make_ddec(spec, sd, dev):
    from dualdiffusion_b200.modules.unets.unet_edm2_ddec_mclt_b1 import DDec_MCLT_UNet_B1, DDec_MCLT_UNet_B1_Config
    cfg = DDec_MCLT_UNet_B1_Config(in_num_freqs=spec.in_num_freqs, in_psd_freqs=spec.in_psd_freqs,
                                   model_channels=spec.model_channels, logvar_channels=spec.logvar_channels,
                                   channel_mult=tuple(spec.channel_mult), double_midblock=spec.double_midblock,
                                   channel_mult_noise=spec.channel_mult_noise, channel_mult_emb=spec.channel_mult_emb,
                                   num_layers_per_block=spec.num_layers_per_block, mlp_multiplier=spec.mlp_multiplier)
    net = DDec_MCLT_UNet_B1(cfg)
    net.load_state_dict(sd, strict=True)
    return net.requires_grad_(False).train(False).to(dev)


@pytest.mark.parametrize("graphs", [False, True])
def test_ddec_forward_vs_golden_reference_and_oracle(dev, graphs):
    spec = dd.small_ddec_spec()
    sd = dd.synth_ddec_state_dict(spec, seed=0)
    g = load_golden("ddec_small.pt")
    net = make_ddec(spec, sd, dev)
    net.use_cuda_graphs = graphs
    c_skip = 1.0 / (1.0 + g["sigma"].view(-1, 1, 1, 1) ** 2)
    ref32 = dd.ddec_forward(sd, spec, g["x"], g["sigma"], g["x_ref"], torch.float32)
    for _ in range(2):
        d = net(g["x"].to(dev), g["sigma"].to(dev), None, None, g["x_ref"].to(dev))
    assert d.shape == g["d"].shape and d.dtype == torch.float32
    assert rel_err(d, ref32) < BF16_NET and rel_err(d, g["d"]) < BF16_NET
    # the network body alone (D - c_skip*x), which c_skip*x would otherwise mask
    assert rel_err(d.cpu() - c_skip * g["x"], ref32 - c_skip * g["x"]) < 2 * BF16_NET
    dp = net(g["x"].to(dev), g["sigma"].to(dev), None, None, g["x_ref"].to(dev), g["perturbed"].to(dev))
    assert rel_err(dp, g["d_perturbed"]) < BF16_NET
    assert rel_err(net.get_sigma_loss_logvar(g["sigma"].to(dev)), g["logvar"]) < 1e-5
    assert net.get_embeddings(None, None) is None


def test_ddec_default_config_runs_and_matches_oracle_band(dev):
    """Default 15 M-parameter ddec (256 mel rows, 4096 PSD bins) at a short width; the fp32 CPU oracle of the full
    network at this size takes ~10 s."""
    spec = dd.DDecSpec()
    sd = dd.synth_ddec_state_dict(spec, seed=0)
    net = make_ddec(spec, sd, dev)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(1, 2, 256, 32, generator=gen)
    xr = torch.rand(1, 2, 4096, 32, generator=gen)
    sigma = torch.tensor([1.5])
    d = net(x.to(dev), sigma.to(dev), None, None, xr.to(dev))
    ref = dd.ddec_forward(sd, spec, x, sigma, xr, torch.float32)
    c_skip = 1.0 / (1.0 + sigma.view(-1, 1, 1, 1) ** 2)
    assert rel_err(d, ref) < BF16_NET
    assert rel_err(d.cpu() - c_skip * x, ref - c_skip * x) < 2 * BF16_NET


# ------------------------------------------------------------------------------------------
# unet_edm2_q4_ddec.UNet (the 2-D variant of row A17)
# ------------------------------------------------------------------------------------------
def test_q4_stem_permute_and_bias_channel_are_exact(dev):
    """unet_edm2_q4_ddec.py:268-277: x_ref.view(B,C,F,k,W).permute(0,3,1,2,4).reshape(B,k*C,F,W) joined to c_in*x by mp_cat.
    With unit concat weights the PSD channels are pure data movement (+ one bf16 rounding): bit-exact."""
    from dualdiffusion_b200 import ops
    gen = torch.Generator().manual_seed(11)
    B, C, Fq, k, W = 2, 2, 8, 4, 9
    x = torch.randn(B, C, Fq, W, generator=gen)
    xr = torch.rand(B, C, Fq * k, W, generator=gen)
    sigma = torch.tensor([0.5, 3.0])
    got = ops.q4_stem(x.to(dev), xr.to(dev), sigma.to(dev), 1.0, 1.0, 1.0, k, 32).float().cpu()      # [B][F][W][32]
    ref = xr.view(B, C, Fq, k, W).permute(0, 3, 1, 2, 4).reshape(B, k * C, Fq, W).to(torch.bfloat16).float()
    assert torch.equal(got[..., C:C + k * C], ref.permute(0, 2, 3, 1))
    c_in = 1 / (1 + sigma.view(-1, 1, 1, 1) ** 2).sqrt()
    assert rel_err(got[..., :C], (c_in * x).to(torch.bfloat16).float().permute(0, 2, 3, 1)) < 4e-3
    assert torch.equal(got[..., C + k * C], torch.ones(B, Fq, W)) and got[..., C + k * C + 1:].abs().max().item() == 0


def make_q4(spec, sd, dev):
    from dualdiffusion_b200.modules.unets.unet_edm2_q4_ddec import UNet, UNet_Config
    cfg = UNet_Config(in_num_freqs=spec.in_num_freqs, in_psd_freqs=spec.in_psd_freqs, model_channels=spec.model_channels,
                      logvar_channels=spec.logvar_channels, channel_mult=tuple(spec.channel_mult),
                      double_midblock=spec.double_midblock, channel_mult_noise=spec.channel_mult_noise,
                      channel_mult_emb=spec.channel_mult_emb, num_layers_per_block=spec.num_layers_per_block,
                      mlp_multiplier=spec.mlp_multiplier)
    net = UNet(cfg)
    net.load_state_dict(sd, strict=True)
    return net.requires_grad_(False).train(False).to(dev)


@pytest.mark.parametrize("graphs", [False, True])
def test_q4_ddec_forward_vs_golden_reference_and_oracle(dev, graphs):
    spec = dd.small_q4_spec()
    sd = dd.synth_q4_state_dict(spec, seed=0)
    g = load_golden("q4_ddec_small.pt")
    net = make_q4(spec, sd, dev)
    net.use_cuda_graphs = graphs
    ref32 = dd.q4_forward(sd, spec, g["x"], g["sigma"], g["x_ref"], torch.float32)
    for _ in range(2):
        d = net(g["x"].to(dev), g["sigma"].to(dev), None, None, g["x_ref"].to(dev))
    assert d.shape == g["d"].shape and d.dtype == torch.float32
    assert rel_err(d, ref32) < BF16_NET and rel_err(d, g["d"]) < BF16_NET
    c_skip = 1.0 / (1.0 + g["sigma"].view(-1, 1, 1, 1) ** 2)
    assert rel_err(d.cpu() - c_skip * g["x"], ref32 - c_skip * g["x"]) < 2 * BF16_NET
    # the conv_in bias matters: dropping it changes the output far beyond the tolerance (the centre-tap column is live)
    sd0 = dict(sd)
    sd0["enc.conv_in.bias"] = torch.zeros_like(sd["enc.conv_in.bias"])
    assert rel_err(dd.q4_forward(sd0, spec, g["x"], g["sigma"], g["x_ref"]) - c_skip * g["x"], ref32 - c_skip * g["x"]) > 4 * BF16_NET


def test_q4_ddec_default_config_vs_oracle(dev):
    spec = dd.Q4Spec()
    sd = dd.synth_q4_state_dict(spec, seed=0)
    net = make_q4(spec, sd, dev)
    gen = torch.Generator().manual_seed(13)
    x = torch.randn(1, 2, 256, 32, generator=gen)
    xr = torch.rand(1, 2, 2048, 32, generator=gen)
    sigma = torch.tensor([1.5])
    d = net(x.to(dev), sigma.to(dev), None, None, xr.to(dev))
    ref = dd.q4_forward(sd, spec, x, sigma, xr, torch.float32)
    c_skip = 1.0 / (1.0 + sigma.view(-1, 1, 1, 1) ** 2)
    assert rel_err(d, ref) < BF16_NET
    assert rel_err(d.cpu() - c_skip * x, ref - c_skip * x) < 2 * BF16_NET


def test_ddec_sampler_unconditional_vs_oracle(dev):
    """The EDM sampler loop over an unconditional module (pipeline.py:598-752 with unet_class_embeddings = None): one copy of
    the batch, no CFG, the PSD reference passed through -- against the oracle sampler driving the fp32 ddec oracle with the
    same noise draws (3 Heun steps, batch 2)."""
    from oracle import sampler_oracle
    from dualdiffusion_b200.pipelines.dual_diffusion_pipeline import DualDiffusionPipeline, SampleParams
    spec = dd.small_ddec_spec()
    sd = dd.synth_ddec_state_dict(spec, seed=0)
    net = make_ddec(spec, sd, dev)
    gen = torch.Generator().manual_seed(17)
    B = 2
    x_ref = torch.rand(B, 2, spec.in_psd_freqs, 24, generator=gen)
    shape = (B, 2, spec.in_num_freqs, 24)
    rec = {}
    model = lambda x, s: dd.ddec_forward(sd, spec, x, s.expand(x.shape[0]), x_ref, torch.float32)
    ref = sampler_oracle.diffusion_decode_unconditional(model, shape, seed=5, sigma_max=20.0, sigma_min=0.03, num_steps=3,
                                                        record=rec)
    pipe = DualDiffusionPipeline({"ddec": net})
    params = SampleParams(seed=5, num_steps=3, batch_size=B, sigma_max=20.0, sigma_min=0.03)
    out = pipe.diffusion_decode(params, audio_embedding=None, sample_shape=shape, x_ref=x_ref, module=net,
                                initial_noise=rec["initial_noise"], step_noise=lambda i: rec["step_noise"][i])
    assert out.shape == ref.shape
    assert rel_err(out, ref) < BF16_NET, rel_err(out, ref)


def test_generation_chain_latents_to_audio(dev):
    """The live decode path end to end at the default configurations (random weights): latents -> DAE_D3.decode -> mel
    spectrogram -> mel_spec_to_mdct_psd -> unconditional EDM sampler over the q4 ddec UNet (x_ref = PSD) -> mdct_to_raw.
    Every stage is parity-tested on its own above; this checks that the shapes and layouts of the stages fit together."""
    from oracle import dae_oracle as do
    from test_gpu_dae import make_dae
    from dualdiffusion_b200.modules.formats.ms_mdct_dual import MS_MDCT_DualFormat, MS_MDCT_DualFormatConfig
    from dualdiffusion_b200.pipelines.dual_diffusion_pipeline import DualDiffusionPipeline, SampleParams
    dspec = do.DAESpec()
    dae = make_dae(dspec, do.synth_dae_state_dict(dspec, seed=0), dev)
    qspec = dd.Q4Spec()
    ddec = make_q4(qspec, dd.synth_q4_state_dict(qspec, seed=0), dev)
    fmt = MS_MDCT_DualFormat(MS_MDCT_DualFormatConfig())
    gen = torch.Generator().manual_seed(21)
    latents = uo.normalize(torch.randn(1, 8, 32, 32, generator=gen)).to(dev)
    mel = dae.decode(latents, dae.get_embeddings(torch.randn(1, dspec.in_channels_emb, generator=gen)))
    assert mel.shape == (1, 2, 256, 256)
    psd = fmt.mel_spec_to_mdct_psd(mel)
    assert psd.shape == (1, 2, 2048, 256)
    pipe = DualDiffusionPipeline({"ddec": ddec, "format": fmt})
    mdct = pipe.diffusion_decode(SampleParams(seed=3, num_steps=2, batch_size=1, sigma_max=20.0, sigma_min=0.03),
                                 audio_embedding=None, sample_shape=(1, 2, 256, 256), x_ref=psd, module=ddec)
    assert mdct.shape == (1, 2, 256, 256) and torch.isfinite(mdct).all()
    raw = fmt.mdct_to_raw(mdct)
    assert raw.shape == (1, 2, 255 * 256) and torch.isfinite(raw).all() and float(raw.std()) > 0
