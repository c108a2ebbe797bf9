"""CPU tests: the oracle restatement (oracle/) against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py), and against the live reference when /root/reference is present."""
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import ref_shim, sampler_oracle, unet_oracle as uo


def _checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


@pytest.fixture(scope="module")
def small():
    spec = uo.small_spec()
    return spec, uo.synth_state_dict(spec, seed=0)


def test_synth_weights_reproducible(small):
    spec, sd = small
    g = load_golden("unet_small.pt")
    assert _checksum(sd) == pytest.approx(g["weight_checksum"], rel=1e-12)


def test_block_plan_matches_reference_key_set(small):
    spec, sd = small
    shapes = uo.state_dict_shapes(uo.default_spec())
    assert len(shapes) == 291                      # default UNet has 291 tensors
    assert shapes["enc.conv_in.weight"] == (256, 6, 3, 3)
    assert shapes["conv_out.weight"] == (4, 256, 3, 3)
    assert shapes["emb_label.weight"] == (768, 512)


def test_unet_oracle_vs_golden_small(small):
    spec, sd = small
    g = load_golden("unet_small.pt")
    emb = uo.get_embeddings(sd, g["clap"], g["mask"])
    assert rel_err(emb, g["emb"]) < 1e-6
    d = uo.unet_forward(sd, spec, g["x"], g["sigma"], emb)
    assert rel_err(d, g["d"]) < 1e-5
    d = uo.unet_forward(sd, spec, g["x"], g["sigma"], emb, x_ref=g["x_ref"])
    assert rel_err(d, g["d_xref"]) < 1e-5
    d = uo.unet_forward(sd, spec, g["x"], g["sigma"], emb, training=True)
    assert rel_err(d, g["d_train"]) < 1e-5
    assert rel_err(uo.sigma_loss_logvar(sd, g["sigma"]), g["logvar"]) < 1e-6


def test_unet_oracle_vs_golden_default_config1():
    """BASELINE config 1: single EDM2 UNet forward, 1x4x64x64 latent, fp32 on CPU."""
    spec = uo.default_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    g = load_golden("unet_default_c1.pt")
    assert _checksum(sd) == pytest.approx(g["weight_checksum"], rel=1e-12)
    emb = uo.get_embeddings(sd, g["clap"], g["mask"])
    d = uo.unet_forward(sd, spec, g["x"], g["sigma"], emb)
    assert rel_err(d, g["d"]) < 1e-5
    # the body must actually contribute (gains are non-zero): D != c_skip * x
    c_skip = 1.0 / (1.0 + g["sigma"].view(-1, 1, 1, 1) ** 2)
    assert (d - c_skip * g["x"]).std() > 0.1


def test_sampler_oracle_vs_golden(small):
    spec, sd = small
    g = load_golden("sampler_small.pt")
    for name, case in g["cases"].items():
        out = sampler_oracle.diffusion_decode(sd, spec, g["clap"], (1, 4, 32, 48), seed=case["seed"], **case["kwargs"])
        assert rel_err(out, case["sample"]) < 1e-4, name


def test_schedule_oracle_vs_golden():
    g = load_golden("schedules.pt")
    s = sampler_oracle.schedule_edm2(100, 200.0, 0.03, 7.0)
    assert torch.equal(s, g["edm2_100"])
    assert s[0].item() == pytest.approx(200.0, rel=1e-6) and s[-1].item() == pytest.approx(0.03, rel=1e-5)


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")
def test_unet_oracle_vs_live_reference(small):
    spec, sd = small
    ref_shim.install()
    from modules.unets.unet_edm2_b4 import UNet, UNetConfig
    from modules.formats.ms_mdct_dual import MS_MDCT_DualFormat, MS_MDCT_DualFormatConfig
    cfg = UNetConfig(**{k: getattr(spec, k) for k in UNetConfig.__dataclass_fields__ if hasattr(spec, k)})
    net = UNet(cfg).eval()
    net.load_state_dict(sd, strict=True)
    fmt = MS_MDCT_DualFormat(MS_MDCT_DualFormatConfig())
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(1, 4, 16, 32, generator=gen)
    sigma = torch.tensor([7.0])
    clap = torch.randn(1, spec.in_channels_emb, generator=gen)
    with torch.no_grad():
        emb = net.get_embeddings(clap, torch.tensor([True]))
        ref = net(x, sigma, fmt, emb)
    got = uo.unet_forward(sd, spec, x, sigma, uo.get_embeddings(sd, clap, torch.tensor([True])))
    assert rel_err(got, ref) < 1e-5


def test_format_oracle_vs_golden():
    """mel-STFT encode + FGLA decode restatement against the reference's own output (bit-exact: both delegate the
    transforms to torch.stft / istft / lstsq; everything else must match exactly, incl. the momentum aliasing)."""
    from oracle import format_oracle as fo
    g = load_golden("format_small.pt")
    spec = fo.SpectrogramSpec()
    mel = fo.raw_to_sample(g["raw"], spec)
    assert rel_err(mel, g["mel"]) < 1e-6
    for n, ref in g["decoded"].items():
        out = fo.sample_to_raw(g["mel"], spec, n)
        assert out.shape == g["raw"].shape
        assert rel_err(out, ref) < 1e-5, n


def test_format_host_logic_vs_golden():
    """Shape bookkeeping of the drop-in format (no GPU needed)."""
    from dualdiffusion_b200.modules.formats.spectrogram import SpectrogramFormat, SpectrogramFormatConfig, mel_filterbank
    from oracle import format_oracle as fo
    g = load_golden("format_small.pt")
    fmt = SpectrogramFormat(SpectrogramFormatConfig())
    assert tuple(fmt.get_sample_shape(1, 1408768)) == tuple(g["shape_1408768"])
    assert fmt.sample_raw_crop_width(1440000) == g["crop_1440000"]
    assert torch.equal(mel_filterbank(fmt.config), fo.mel_filterbank(fo.SpectrogramSpec()))


def test_ms_dual_oracle_and_host_logic_vs_golden():
    """Live format (MS_MDCT_DualFormat.raw_to_mel_spec): oracle restatement and the drop-in's shape helpers."""
    from oracle import format_oracle as fo
    from dualdiffusion_b200.modules.formats.ms_mdct_dual import MS_MDCT_DualFormat, MS_MDCT_DualFormatConfig
    g = load_golden("ms_dual_small.pt")
    assert rel_err(fo.raw_to_mel_spec(g["raw"], fo.MSDualSpec()), g["mel"]) < 1e-6
    fmt = MS_MDCT_DualFormat(MS_MDCT_DualFormatConfig())
    assert tuple(fmt.get_mel_spec_shape(3)) == tuple(g["mel_shape_default"])
    assert fmt.get_raw_crop_width() == g["crop_default"]
    assert torch.equal(fmt.ms_freq_scale.get_unscaled(34), g["unscaled_34"])


def test_oracle_train_step_gradients_vs_golden_reference():
    """Row A12: autograd through the oracle's train loss (unet_trainer.py:259-280 restated) reproduces the
    reference module's loss and its parameter gradients (norm and seeded projection per tensor, plus every
    small tensor in full)."""
    g = load_golden("unet_small_train.pt")
    spec = uo.small_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    sdg = {k: (v.clone().requires_grad_(True) if "fourier" not in k else v) for k, v in sd.items()}
    loss = uo.train_loss(sdg, spec, g["samples"], g["noise"], g["sigma"], g["clap"], g["mask"])
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    loss.backward()
    assert set(g["grad_stats"]) == {k for k in sdg if "fourier" not in k}
    for name, (n_ref, dot_ref) in g["grad_stats"].items():
        gr = sdg[name].grad
        assert abs(float(gr.norm()) - n_ref) <= 1e-4 * n_ref + 1e-9, name
        dot = float((gr * uo.grad_probe(name, gr.shape)).sum())
        assert abs(dot - dot_ref) <= 1e-3 * n_ref * 3 + 1e-9, name
    for name, gref in g["small_grads"].items():
        assert rel_err(sdg[name].grad, gref) < 1e-4 or (sdg[name].grad - gref).abs().max() < 1e-8, name


def test_dae_oracle_decode_vs_golden_reference():
    """Row A16: the DAE_D3 decoder restatement (oracle/dae_oracle.py) against the reference module's output."""
    from oracle import dae_oracle as do
    g = load_golden("dae_small.pt")
    spec = do.small_dae_spec()
    sd = do.synth_dae_state_dict(spec, seed=0)
    assert abs(float(sum(v.double().abs().sum() for v in sd.values())) - g["weight_checksum"]) < 1e-6 * g["weight_checksum"]
    for tag, c in g["cases"].items():
        emb = do.dae_get_embeddings(sd, c["emb_in"])
        assert rel_err(emb, c["emb"]) < 1e-6
        mel = do.dae_decode(sd, spec, c["latents"], emb)
        assert mel.shape == c["mel"].shape and rel_err(mel, c["mel"]) < 1e-5, tag
    e = g["encode"]
    assert rel_err(do.dae_encode(sd, spec, e["mel"]), e["latents"]) < 1e-5
    assert rel_err(do.dae_encode(sd, spec, e["mel"], training=True), e["pre_norm"]) < 1e-5
    r = 2 ** (len(spec.channel_mult_dec) - 1)      # get_mel_spec_shape / get_latent_shape (:323-342)
    assert tuple(g["mel_shape"]) == (3, 2, 32 * r, 688 * r) and tuple(g["latent_shape"]) == (3, 8, 256 // r, 5504 // r)


def test_ddec_oracle_vs_golden_reference():
    """Row A17: the DDec_MCLT_UNet_B1 restatement against the reference module.  With the reference's hard-coded bf16
    body dtype the restatement is bit-exact; the fp32 statement of the same arithmetic differs by the reference's own
    bf16 rounding."""
    from oracle import ddec_oracle as dd
    g = load_golden("ddec_small.pt")
    spec = dd.small_ddec_spec()
    sd = dd.synth_ddec_state_dict(spec, seed=0)
    d = dd.ddec_forward(sd, spec, g["x"], g["sigma"], g["x_ref"], torch.bfloat16)
    assert torch.equal(d, g["d"])
    d32 = dd.ddec_forward(sd, spec, g["x"], g["sigma"], g["x_ref"], torch.float32)
    assert rel_err(d32, g["d"]) < 1e-2
    assert rel_err(dd.ddec_sigma_loss_logvar(sd, g["sigma"]), g["logvar"]) < 1e-6


def test_q4_ddec_oracle_vs_golden_reference():
    """Row A17 (2-D variant): unet_edm2_q4_ddec.UNet restatement against the reference module (bf16 body in the
    reference; channels_last CPU kernels order the bf16 accumulation differently, hence bf16-level agreement)."""
    from oracle import ddec_oracle as dd
    g = load_golden("q4_ddec_small.pt")
    spec = dd.small_q4_spec()
    sd = dd.synth_q4_state_dict(spec, seed=0)
    c_skip = 1.0 / (1.0 + g["sigma"].view(-1, 1, 1, 1) ** 2)
    for dt, tol in ((torch.bfloat16, 8e-3), (torch.float32, 1e-2)):
        d = dd.q4_forward(sd, spec, g["x"], g["sigma"], g["x_ref"], dt)
        assert rel_err(d, g["d"]) < tol
        assert rel_err(d - c_skip * g["x"], g["d"] - c_skip * g["x"]) < 3 * tol


def test_dae_q4_oracle_matches_reference_golden():
    """dae_edm2_q4.DAE restatement (encode and decode, with and without add_pixel_norm) against the
    unmodified reference module run in fp32 on CPU (tests/golden/make_golden_dae_q4.py)."""
    from oracle import dae_q4_oracle as qo
    g = load_golden("dae_q4_small.pt")
    for tag, case in g.items():
        spec = qo.small_dae_q4_spec()
        spec.add_pixel_norm = tag == "pixel_norm"
        sd = qo.synth_dae_q4_state_dict(spec, seed=0)
        assert abs(float(sum(v.double().abs().sum() for v in sd.values())) - case["weight_checksum"]) < 1e-6 * case["weight_checksum"]
        assert rel_err(qo.dae_q4_encode(sd, spec, case["mel"]), case["latents"]) < 1e-5, tag
        assert rel_err(qo.dae_q4_decode(sd, spec, case["lat_in"]), case["decoded"]) < 1e-5, tag


def test_dae_q4_module_state_dict_matches_reference_layout():
    """The product module's parameters / buffers carry the reference's names and shapes (the golden's weights, which the
    unmodified reference loaded with strict=True, load here with strict=True too) and it refuses to run without CUDA."""
    from oracle import dae_q4_oracle as qo
    from dualdiffusion_b200.modules.daes.dae_edm2_q4 import DAE, DAE_Config
    spec = qo.small_dae_q4_spec()
    sd = qo.synth_dae_q4_state_dict(spec, seed=0)
    dae = DAE(DAE_Config(model_channels=spec.model_channels, channel_mult_enc=tuple(spec.channel_mult_enc),
                         channel_mult_dec=tuple(spec.channel_mult_dec), num_enc_layers_per_block=spec.num_enc_layers_per_block,
                         num_dec_layers_per_block=spec.num_dec_layers_per_block)).eval()
    dae.load_state_dict(sd, strict=True)
    assert {k: tuple(v.shape) for k, v in dae.state_dict().items()} == qo.dae_q4_state_dict_shapes(spec)
    assert tuple(dae.get_latent_shape((2, 2, 32, 72))) == load_golden("dae_q4_small.pt")["plain"]["latent_shape"]
    with torch.no_grad(), pytest.raises(RuntimeError):
        dae.decode(torch.zeros(1, spec.latent_channels, 4, 4), None)


def test_ms_dual2_oracle_vs_golden_reference():
    """the second MS_MDCT_DualFormat lineage (three-window mel-STFT blended per filter, utils/mdct MDCT)
    restated in oracle/format_oracle.py against the unmodified reference's outputs (tests/golden/make_golden_ms_dual2.py)."""
    from oracle import format_oracle as fo
    g = load_golden("ms_dual2_small.pt")
    spec = fo.MSDual2Spec()
    assert torch.equal(fo.ms2_filter_window_weights(spec), g["window_weights"])
    assert rel_err(fo.ms2_raw_to_mel_spec(g["raw"], spec), g["mel"]) < 1e-6
    assert rel_err(fo.ms2_mel_spec_to_linear(g["mel"], spec), g["mel_linear"]) < 1e-5
    assert rel_err(fo.ms2_raw_to_mdct(g["raw"], spec), g["mdct"]) < 1e-6
    assert rel_err(fo.ms2_raw_to_mdct(g["raw_odd"], spec), g["mdct_odd"]) < 1e-6
    assert rel_err(fo.ms2_mdct_to_raw(g["mdct"], spec), g["raw_back"]) < 1e-6
    phase, psd = fo.ms2_raw_to_mdct_phase_psd(g["raw"], spec)
    assert rel_err(phase, g["phase"]) < 1e-5 and rel_err(psd, g["psd"]) < 1e-6


def test_ms_dual2_host_tables_vs_reference_golden():
    """The product format's host-built tables (fp64 MDCT / inverse MDCT matrices with density and scales folded in, the
    min-norm inverse mel bank, the constructor's window weights) applied with plain torch on CPU reproduce the unmodified
    reference's outputs -- the GPU path only has to multiply by them.  Shape helpers as the reference states them."""
    from dualdiffusion_b200.modules.formats.ms_mdct_dual_2 import MS_MDCT_DualFormat, MS_MDCT_DualFormatConfig
    g = load_golden("ms_dual2_small.pt")
    f = MS_MDCT_DualFormat(MS_MDCT_DualFormatConfig())
    t = f._tables(torch.device("cpu"))
    assert torch.equal(f.ms_filter_window_weights, g["window_weights"])
    raw, N = g["raw"], 256
    n = raw.shape[-1]
    T = -(-n // N) + 1
    frames = torch.nn.functional.pad(raw.reshape(4, n), (N, 2 * N), mode="reflect").unfold(-1, 2 * N, N)[:, :T]
    y = torch.einsum("rk,stk->srt", t["fwd"], frames)
    assert rel_err(y[:, :N].reshape(2, 2, N, T), g["mdct"]) < 1e-4
    inv = torch.einsum("skt,kj->stj", g["mdct"].reshape(4, N, T), t["inv"])
    out = torch.zeros(4, (T + 1) * N)
    for tt in range(T):
        out[:, tt * N:tt * N + 2 * N] += inv[:, tt]
    assert rel_err(out[:, N:-N].reshape(2, 2, -1), g["raw_back"]) < 1e-4
    c = f.config
    lin = (g["mel"] - c.raw_to_mel_spec_offset / c.raw_to_mel_spec_scale).clip(min=0) ** (1 / c.ms_abs_exponent)
    assert rel_err(torch.einsum("bf,scft->scbt", t["pinv"], lin), g["mel_linear"]) < 1e-5
    assert tuple(f.get_mel_spec_shape(3, 1408768)) == g["mel_spec_shape"] and tuple(f.get_mdct_shape(3, 1408768)) == g["mdct_shape"]
    assert f.get_raw_crop_width(1408768) == g["raw_crop_width"]
    with pytest.raises(RuntimeError):
        f.raw_to_mdct(raw)                                                  # no CPU path


def test_b4_3_oracle_matches_reference_golden():
    """SURVEY 8(f) N4, the lineage without a CUDA path yet: unet_edm2_b4_3.UNet (time-axis blocks, (1,3) grouped convs,
    partial RoPE, in-place input skip) restated in oracle/unet_b4_3_oracle.py.  With the reference's hard-coded bf16 body the
    restatement is bit-identical to the unmodified reference module on CPU (tests/golden/make_golden_b4_3.py); the fp32
    body differs from it by the bf16 rounding only."""
    from oracle import unet_b4_3_oracle as bo
    g = load_golden("unet_b4_3_small.pt")
    spec = bo.small_b4_3_spec()
    sd = bo.synth_state_dict(spec, seed=0)
    assert abs(float(sum(v.double().abs().sum() for v in sd.values())) - g["weight_checksum"]) < 1e-6 * g["weight_checksum"]
    assert torch.equal(bo.get_embeddings(sd, g["clap"], g["mask"]), g["emb"])
    assert torch.equal(bo.forward(sd, spec, g["x"], g["sigma"], g["emb"], None, torch.bfloat16), g["d"])
    assert torch.equal(bo.forward(sd, spec, g["x"], g["sigma"], g["emb"], g["x_ref"], torch.bfloat16), g["d_xref"])
    assert torch.equal(bo.sigma_loss_logvar(sd, spec, g["sigma"]), g["logvar"])
    c_skip = 1.0 / (1.0 + g["sigma"].view(-1, 1, 1, 1) ** 2)
    d32 = bo.forward(sd, spec, g["x"], g["sigma"], g["emb"], None, torch.float32)
    assert rel_err(d32 - c_skip * g["x"], g["d"] - c_skip * g["x"]) < 3e-2
    # RoPE's even | odd | tail re-ordering is applied to q and k alike: the scores, hence the output, do not depend on it
    q = torch.randn(1, 2, 5, 32)
    k = torch.randn(1, 2, 5, 32)
    tables = bo.rope_tables(5, 24, 10000.0, torch.float32)
    scores = bo.rope_rotate(q, tables) @ bo.rope_rotate(k, tables).transpose(-1, -2)
    cos, sin = tables
    def rot_in_place(x):
        y = x.clone()
        y[..., 0:24:2] = x[..., 0:24:2] * cos - x[..., 1:24:2] * sin
        y[..., 1:24:2] = x[..., 1:24:2] * cos + x[..., 0:24:2] * sin
        return y
    assert rel_err(scores, rot_in_place(q) @ rot_in_place(k).transpose(-1, -2)) < 1e-6


def test_dae_q4_tiled_encode_chunk_placement_host_logic(monkeypatch):
    """dae_edm2_q4.DAE.tiled_encode (:314-372) is host logic around `encode`: with a purely local stand-in encoder (8x8
    average pooling, no receptive field beyond the cell) the stitched result must equal the whole-input result exactly,
    for widths that leave a short last chunk (extended to the left) and for one that fits a single chunk."""
    from oracle import dae_q4_oracle as qo
    from dualdiffusion_b200.modules.daes.dae_edm2_q4 import DAE, DAE_Config
    spec = qo.small_dae_q4_spec()
    dae = DAE(DAE_Config(model_channels=spec.model_channels, channel_mult_enc=tuple(spec.channel_mult_enc),
                         channel_mult_dec=tuple(spec.channel_mult_dec), num_enc_layers_per_block=1,
                         num_dec_layers_per_block=1)).eval()
    ds, Lc = dae.downsample_ratio, dae.config.latent_channels
    calls = []

    def fake_encode(x, embeddings=None, training=False):
        calls.append(x.shape[-1])
        return torch.nn.functional.avg_pool2d(x.mean(dim=1, keepdim=True), ds).expand(-1, Lc, -1, -1).contiguous()

    monkeypatch.setattr(dae, "encode", fake_encode)
    g = torch.Generator().manual_seed(5)
    for width, max_chunk, overlap in ((ds * 40, ds * 16, ds * 2), (ds * 37, ds * 16, ds * 4), (ds * 12, ds * 16, ds * 2)):
        calls.clear()
        x = torch.randn(2, 2, ds * 4, width, generator=g)
        whole = fake_encode(x)
        calls.clear()
        tiled = dae.tiled_encode(x, None, max_chunk=max_chunk, overlap=overlap)
        assert tiled.shape == whole.shape and torch.equal(tiled, whole), (width, max_chunk, overlap)
        assert all(c <= max_chunk for c in calls) and (len(calls) == 1) == (width <= max_chunk)


def test_mdct_oracle_vs_golden_reference():
    """MCLT / inverse MCLT / PSD / mel -> PSD restatements against the reference's own outputs."""
    from oracle import format_oracle as fo
    g = load_golden("mdct_small.pt")
    for tag, c in g["cases"].items():
        spec = fo.MDCTSpec(**c["kwargs"])
        assert rel_err(fo.raw_to_mdct(g["raw"], spec), c["mdct"]) < 1e-6, tag
        assert rel_err(fo.raw_to_mdct_psd(g["raw"], spec), c["psd"]) < 1e-6, tag
        assert rel_err(fo.mdct_to_raw(c["mdct"], spec), c["raw_back"]) < 1e-6, tag
        assert rel_err(fo.mel_spec_to_mdct_psd(c["mel"], fo.MSDualSpec(), spec), c["mel_psd"]) < 1e-5, tag


def test_axis_attention_oracle_vs_reference_golden():
    """Row A9: the row/col <-> batch reshape attention of the legacy ddec UNets, pinned against the unmodified reference
    Block (tests/golden/make_golden_axis_attention.py hooks attn_qkv / attn_proj of
    src/modules/unets/old/unet_edm2_ddec_mdct_b3.py)."""
    g = load_golden("axis_attention_b3.pt")
    for tag, case in g.items():
        got = uo.mp_silu(uo.axis_attention_b3(case["qkv"], case["heads"]))
        assert got.shape == case["y_silu"].shape
        assert rel_err(got, case["y_silu"]) < 1e-5, tag


def test_b4_2_oracle_vs_reference_golden():
    """Row N4: the b4_2 lineage restatement against the unmodified reference (tests/golden/make_golden_b4_2.py)."""
    from oracle import unet_b4_2_oracle as bo
    g = load_golden("unet_b4_2_small.pt")
    spec = bo.small_spec()
    sd = bo.synth_state_dict(spec, seed=0)
    assert _checksum(sd) == pytest.approx(g["weight_checksum"], rel=1e-12)
    emb = bo.get_embeddings(sd, g["clap"], g["mask"])
    assert rel_err(emb, g["emb"]) < 1e-6
    assert rel_err(bo.unet_forward(sd, spec, g["x"], g["sigma"], emb), g["d"]) < 1e-5
    assert rel_err(bo.unet_forward(sd, spec, g["x"], g["sigma"], emb, g["x_ref"]), g["d_xref"]) < 1e-5
    assert rel_err(bo.sigma_loss_logvar(sd, spec, g["sigma"]), g["logvar"]) < 1e-6
