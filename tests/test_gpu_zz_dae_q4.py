"""GPU parity tests (-m gpu) of the 2-D diffusion autoencoder lineage dae_edm2_q4.DAE:
the four entry points of its two ends (index maps bit-exact, the direct 5x5 convolution against torch fp32 on the same
bf16-rounded operands), and the assembled encode / decode against the unmodified reference's golden
(tests/golden/dae_q4_small.pt) and the fp32 oracle, reduced and default configurations."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_err
from oracle import dae_q4_oracle as qo

pytestmark = pytest.mark.gpu
BF16_OP, BF16_NET = 4e-3, 3e-2


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def make_dae_q4(spec, sd, dev):
    from dualdiffusion_b200.modules.daes.dae_edm2_q4 import DAE, DAE_Config
    cfg = DAE_Config(model_channels=spec.model_channels, channel_mult_enc=tuple(spec.channel_mult_enc),
                     channel_mult_dec=tuple(spec.channel_mult_dec), num_enc_layers_per_block=spec.num_enc_layers_per_block,
                     num_dec_layers_per_block=spec.num_dec_layers_per_block, latent_channels=spec.latent_channels,
                     mlp_multiplier=spec.mlp_multiplier, res_balance=spec.res_balance, add_pixel_norm=spec.add_pixel_norm)
    dae = DAE(cfg)
    dae.load_state_dict(sd, strict=True)
    return dae.requires_grad_(False).train(False).to(dev)


def test_pack_unpack_and_patches_are_exact(dev):
    """dd_pack_nhwc / dd_unpack_nchw / dd_patches5x5 are index maps plus one bf16 rounding: bit-exact against the statement
    in torch (F.unfold orders a 5x5 patch channel-major, the GEMM operand tap-major; ragged sizes, odd widths)."""
    from dualdiffusion_b200 import ops
    gen = torch.Generator().manual_seed(41)
    for B, C, H, W in ((2, 8, 5, 7), (1, 3, 1, 33), (3, 2, 9, 1)):
        x = torch.randn(B, C, H, W, generator=gen)
        got = ops.pack_nhwc(x.to(dev), 32, ones_channel=C).float().cpu()
        assert torch.equal(got[..., :C], x.to(torch.bfloat16).float().permute(0, 2, 3, 1))
        assert torch.equal(got[..., C], torch.ones(B, H, W)) and got[..., C + 1:].abs().max().item() == 0
        assert ops.pack_nhwc(x.to(dev), 16 if C <= 8 else 32).float().cpu()[..., C:].abs().max().item() == 0
        back = ops.unpack_nchw(ops.pack_nhwc(x.to(dev), 32, ones_channel=C), C).cpu()
        assert torch.equal(back, x.to(torch.bfloat16).float())
    for B, C, H, W in ((2, 2, 6, 11), (1, 2, 3, 2), (1, 5, 4, 9)):
        x = torch.randn(B, C, H, W, generator=gen)
        cols = 64 if 25 * C + 1 <= 64 else 128
        got = ops.patches5x5(x.to(dev), cols).float().cpu()                                          # [B][H][W][cols]
        ref = F.unfold(x.to(torch.bfloat16).float(), 5, padding=2).view(B, C, 25, H, W)              # [B][C][tap][H][W]
        ref = ref.permute(0, 3, 4, 2, 1).reshape(B, H, W, 25 * C)
        assert torch.equal(got[..., :25 * C], ref)
        assert torch.equal(got[..., 25 * C], torch.ones(B, H, W)) and got[..., 25 * C + 1:].abs().max().item() == 0


@pytest.mark.parametrize("shape", [(2, 7, 45, 64, 2), (1, 3, 32, 32, 1), (1, 12, 70, 96, 4), (1, 1, 1, 64, 3)])
def test_conv5x5_dense_vs_torch(dev, shape):
    """dd_conv5x5_dense against F.conv2d in fp32 on the same bf16-rounded activation (fp32 weights and accumulation in
    both: only the summation order differs)."""
    from dualdiffusion_b200 import ops
    B, H, W, C, Cout = shape
    gen = torch.Generator().manual_seed(43)
    x = torch.randn(B, C, H, W, generator=gen).to(torch.bfloat16)
    w = torch.randn(Cout, C, 5, 5, generator=gen)
    gain = torch.tensor([0.7])
    wq = (w / math.sqrt(C * 25)).permute(0, 2, 3, 1).reshape(Cout, 25, C).contiguous()
    got = ops.conv5x5_dense(x.permute(0, 2, 3, 1).contiguous().to(dev), wq.to(dev), gain.to(dev)).cpu()
    ref = F.conv2d(x.float(), w * (0.7 / math.sqrt(C * 25)), padding=2)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 1e-5


@pytest.mark.parametrize("graphs", [False, True])
@pytest.mark.parametrize("tag", ["plain", "pixel_norm"])
def test_dae_q4_vs_golden_reference_and_oracle(dev, graphs, tag):
    g = load_golden("dae_q4_small.pt")[tag]
    spec = qo.small_dae_q4_spec()
    spec.add_pixel_norm = tag == "pixel_norm"
    sd = qo.synth_dae_q4_state_dict(spec, seed=0)
    dae = make_dae_q4(spec, sd, dev)
    dae.use_cuda_graphs = graphs
    for _ in range(2):
        lat = dae.encode(g["mel"].to(dev), None)
        rec = dae.decode(g["lat_in"].to(dev), None)
    assert lat.shape == g["latents"].shape and rec.shape == g["decoded"].shape
    assert rel_err(lat, g["latents"]) < BF16_NET and rel_err(lat, qo.dae_q4_encode(sd, spec, g["mel"])) < BF16_NET
    assert rel_err(rec, g["decoded"]) < BF16_NET and rel_err(rec, qo.dae_q4_decode(sd, spec, g["lat_in"])) < BF16_NET
    assert tuple(dae.get_latent_shape((2, 2, 32, 72))) == g["latent_shape"]
    assert tuple(dae.get_mel_spec_shape((2, 16, 8, 18))) == g["mel_spec_shape"]
    assert dae.get_embeddings(torch.zeros(2, 4)) is None
    # both biases matter: dropping either moves the result far beyond the tolerance (the bias columns are live)
    for key, fn, inp, ref in (("enc.conv_in.bias", qo.dae_q4_encode, g["mel"], g["latents"]),
                              ("conv_latents_in.bias", qo.dae_q4_decode, g["lat_in"], g["decoded"])):
        sd0 = dict(sd)
        sd0[key] = torch.zeros_like(sd[key])
        assert rel_err(fn(sd0, spec, inp), ref) > 4 * BF16_NET


def test_dae_q4_weight_update_is_seen_and_round_trip_shapes(dev):
    """A parameter changed in place (an optimizer step, load_state_dict) is picked up by the next call -- the captured graphs
    read the prepared weights through stable pointers -- and the state_dict round-trips through the module."""
    spec = qo.small_dae_q4_spec()
    sd = qo.synth_dae_q4_state_dict(spec, seed=0)
    dae = make_dae_q4(spec, sd, dev)
    assert set(dae.state_dict().keys()) == set(sd.keys())
    gen = torch.Generator().manual_seed(47)
    lat = torch.randn(1, spec.latent_channels, 4, 8, generator=gen)
    a = dae.decode(lat.to(dev), None)
    sd2 = qo.synth_dae_q4_state_dict(spec, seed=1)
    dae.load_state_dict(sd2, strict=True)
    b = dae.decode(lat.to(dev), None)
    assert rel_err(b, qo.dae_q4_decode(sd2, spec, lat)) < BF16_NET
    assert rel_err(a, qo.dae_q4_decode(sd, spec, lat)) < BF16_NET and rel_err(a, b) > 0.5


def test_dae_q4_default_config_and_tiled_encode(dev):
    """The dataclass-default configuration (64 x (1,2,4,8), 3 layers per block, 256 mel bins) against the fp32 oracle, and
    tiled_encode against the whole-input encode away from the chunk borders' receptive field."""
    spec = qo.DAEQ4Spec()
    sd = qo.synth_dae_q4_state_dict(spec, seed=0)
    dae = make_dae_q4(spec, sd, dev)
    gen = torch.Generator().manual_seed(53)
    mel = torch.randn(1, 2, 256, 64, generator=gen)
    lat = dae.encode(mel.to(dev), None)
    ref = qo.dae_q4_encode(sd, spec, mel)
    assert lat.shape == (1, 8, 32, 8) and rel_err(lat, ref) < BF16_NET
    lat_in = torch.randn(1, 8, 32, 8, generator=gen)
    rec = dae.decode(lat_in.to(dev), None)
    assert rec.shape == (1, 2, 256, 64) and rel_err(rec, qo.dae_q4_decode(sd, spec, lat_in)) < BF16_NET
    # tiled: 3 chunks of 128 columns with 32 columns of overlap on a 256-column input
    mel2 = torch.randn(1, 2, 256, 256, generator=gen).to(dev)
    whole = dae.encode(mel2, None)
    tiled = dae.tiled_encode(mel2, None, max_chunk=128, overlap=32)
    assert tiled.shape == whole.shape
    # chunks see a truncated context (that is the approximation tiling makes); misplaced chunks would be uncorrelated (~1.4)
    print("tiled vs whole:", rel_err(tiled, whole), "first columns:", rel_err(tiled[..., :4], whole[..., :4]))
    assert rel_err(tiled, whole) < 0.1 and rel_err(tiled[..., :4], whole[..., :4]) < BF16_NET      # measured 0.019 / 0.002


def test_dae_q4_refuses_what_is_not_built(dev):
    from dualdiffusion_b200.modules.daes.dae_edm2_q4 import DAE, DAE_Config
    with pytest.raises(NotImplementedError):
        DAE(DAE_Config(in_channels_emb=512))
    with pytest.raises(NotImplementedError):
        DAE(DAE_Config(attn_levels=(3,)))
    spec = qo.small_dae_q4_spec()
    dae = make_dae_q4(spec, qo.synth_dae_q4_state_dict(spec, seed=0), dev)
    with pytest.raises(ValueError):
        dae.encode(torch.zeros(1, 2, 30, 72, device=dev), None)            # H not a multiple of the downsample ratio
    with pytest.raises(ValueError):
        dae.decode(torch.zeros(1, 4, 8, 8, device=dev), None)              # wrong latent channel count
    with pytest.raises(NotImplementedError):
        dae.encode(torch.zeros(1, 2, 32, 72, device=dev), None, training=True)
    with pytest.raises(RuntimeError):
        make_dae_q4(spec, qo.synth_dae_q4_state_dict(spec, seed=0), torch.device("cpu")).decode(torch.zeros(1, 8, 4, 4), None)
