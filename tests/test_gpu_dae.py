"""GPU parity tests (-m gpu) of the DAE_D3 decoder (BASELINE config 5): the folded-stereo /
halo-column kernels of csrc/dae.cu and the assembled `decode` against the CPU oracle (oracle/dae_oracle.py) and the
golden output of the unmodified reference (tests/golden/dae_small.pt).  Index-only ops are bit-exact; bf16 tensor-core
ops 4e-3 relative L2 per op, the whole decoder 3e-2."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_err
from oracle import dae_oracle as do, unet_oracle as uo

pytestmark = pytest.mark.gpu
BF16_OP, BF16_NET = 4e-3, 3e-2
PW = 2


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def bf16_round(x):
    return x.to(torch.bfloat16).float()


def fold(x5, dev):
    """(B, C, 2, H, W) fp32 -> folded NHWC bf16 [B][H][W+2PW][2C] with mirrored halo columns."""
    b, c, z, h, w = x5.shape
    x = x5.permute(0, 3, 4, 2, 1).reshape(b, h, w, z * c)                       # channel = z*C + c
    x = F.pad(x.permute(0, 3, 1, 2), (PW, PW, 0, 0), mode="reflect").permute(0, 2, 3, 1)
    return x.contiguous().to(device=dev, dtype=torch.bfloat16)


def unfold(y, c):
    """folded [B][H][Wp][2C] -> (B, C, 2, H, W) fp32 on the CPU (halo columns dropped)."""
    b, h, wp, _ = y.shape
    return y.float().cpu()[:, :, PW:wp - PW].reshape(b, h, wp - 2 * PW, 2, c).permute(0, 4, 3, 1, 2)


@pytest.mark.parametrize("B,C,Co,H,W,kz,k", [(1, 32, 64, 16, 24, 2, 3), (2, 64, 32, 4, 11, 2, 3), (1, 128, 64, 8, 40, 2, 3),
                                             (2, 64, 32, 8, 12, 1, 1), (1, 32, 32, 16, 9, 1, 3)])
def test_mpconv3d_as_folded_conv_vs_oracle(dev, B, C, Co, H, W, kz, k):
    """MPConv3D (kz,k,k) with W reflection / one-sided Z reflection / H zero padding (dae_edm2_d3.py:60-86) ==
    2-D tensor-core convolution on the stereo-folded, halo-padded layout with the block-circulant weight."""
    from dualdiffusion_b200 import ops
    gen = torch.Generator().manual_seed(C + Co + H + W + kz)
    x5 = bf16_round(torch.randn(B, C, 2, H, W, generator=gen))
    w = torch.randn(Co, C, kz, k, k, generator=gen)
    ref = do.mp_conv3d(x5, bf16_round(w / (C * kz * k * k) ** 0.5) * (C * kz * k * k) ** 0.5)
    wz = ops.weight_prep_z2(w.to(dev))
    y = ops.mpconv(fold(x5, dev), wz, k, 1 if kz == 2 else 2)
    assert rel_err(unfold(y, Co), ref) < BF16_OP
    ops.reflect_fill_w(y, PW)
    yl = y.float().cpu()
    assert torch.equal(yl[:, :, PW - 1], yl[:, :, PW + 1]) and torch.equal(yl[:, :, PW - 2], yl[:, :, PW + 2])
    assert torch.equal(yl[:, :, -PW], yl[:, :, -PW - 2]) and torch.equal(yl[:, :, -1], yl[:, :, -PW - 3])


def test_weight_prep_z2_layout_is_exact(dev):
    from dualdiffusion_b200 import ops
    O, I, taps = 16, 8, 9
    w = (torch.arange(O * I * 2 * taps, dtype=torch.float32) % 251).view(O, I, 2, 3, 3)
    s = (I * 2 * taps) ** 0.5
    got = ops.weight_prep_z2(w.to(dev), gain_host=s, i_stride=32).float().cpu()          # [2O][9][32]
    for zp in range(2):
        for z in range(2):
            blk = got[zp * O:(zp + 1) * O, :, z * I:(z + 1) * I]                        # [O][tap][I]
            assert torch.equal(blk, w[:, :, z ^ zp].reshape(O, I, taps).permute(0, 2, 1))
    assert got[:, :, 2 * I:].abs().max().item() == 0
    w1 = (torch.arange(O * I * taps, dtype=torch.float32) % 251).view(O, I, 1, 3, 3)
    got = ops.weight_prep_z2(w1.to(dev), gain_host=(I * taps) ** 0.5).float().cpu()     # [2O][9][I]
    ref = w1.reshape(O, I, taps).permute(0, 2, 1)
    assert torch.equal(got[:O], ref) and torch.equal(got[O:], ref)


def test_dae_stem_and_upsample_are_exact(dev):
    from dualdiffusion_b200 import ops
    gen = torch.Generator().manual_seed(3)
    B, Lc, H, W = 2, 4, 4, 7
    lat = torch.randn(B, 2 * Lc, H, W, generator=gen)
    x = ops.dae_stem(lat.to(dev), Lc, PW, 32).float().cpu()
    x5 = lat.reshape(B, Lc, 2, H, W)                                                     # tensor_4d_to_5d
    x5 = torch.cat((x5, torch.ones_like(x5[:, :1])), dim=1)
    ref = fold(x5, "cpu").float()
    assert torch.equal(x[..., :2 * (Lc + 1)], ref) and x[..., 2 * (Lc + 1):].abs().max().item() == 0
    a5 = bf16_round(torch.randn(B, 16, 2, H, W, generator=gen))
    xc, s = ops.up2_silu_pad(fold(a5, dev), PW)
    up = a5.repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2)                    # resample_3d "up"
    assert torch.equal(xc.float().cpu(), fold(up, "cpu").float())
    assert rel_err(unfold(s, 16), uo.mp_silu(up)) < BF16_OP
    sl = s.float().cpu()
    assert torch.equal(sl[:, :, PW - 1], sl[:, :, PW + 1]) and torch.equal(sl[:, :, -1], sl[:, :, -PW - 3])


@pytest.mark.parametrize("C", [32, 64])
def test_conv5x5_out_vs_oracle(dev, C):
    from dualdiffusion_b200 import ops, _lib as L
    gen = torch.Generator().manual_seed(5 + C)
    B, H, W = 2, 9, 14
    x5 = bf16_round(torch.randn(B, C, 2, H, W, generator=gen))
    w = torch.randn(1, C, 1, 5, 5, generator=gen)
    gain = torch.tensor(0.7)
    ref = do.mp_conv3d(x5, w, gain)                                                      # (B, 1, 2, H, W)
    w25 = ops.weight_prep(w.reshape(1, C, 5, 5).to(dev), fmt=L.WFMT_F32_OIT)
    out = ops.conv5x5_out(fold(x5, dev), w25, gain.to(dev).reshape(1), PW)
    assert out.shape == (B, 2, H, W) and out.dtype == torch.float32
    assert rel_err(out, ref.reshape(B, 2, H, W)) < 1e-5


def make_dae(spec, sd, dev):
    from dualdiffusion_b200.modules.daes.dae_edm2_d3 import DAE_D3, DAE_D3_Config
    cfg = DAE_D3_Config(in_channels_emb=spec.in_channels_emb, model_channels=spec.model_channels,
                        channel_mult_enc=spec.channel_mult_enc, channel_mult_dec=tuple(spec.channel_mult_dec),
                        channel_mult_emb=spec.channel_mult_emb, num_enc_layers=spec.num_enc_layers,
                        num_dec_layers_per_block=spec.num_dec_layers_per_block, mlp_multiplier=spec.mlp_multiplier)
    net = DAE_D3(cfg)
    net.load_state_dict(sd, strict=True)
    return net.requires_grad_(False).train(False).to(dev)


@pytest.mark.parametrize("graphs", [False, True])
def test_dae_decode_vs_golden_reference(dev, graphs):
    spec = do.small_dae_spec()
    sd = do.synth_dae_state_dict(spec, seed=0)
    g = load_golden("dae_small.pt")
    net = make_dae(spec, sd, dev)
    net.use_cuda_graphs = graphs
    for tag, c in g["cases"].items():
        emb = net.get_embeddings(c["emb_in"])
        assert rel_err(emb, c["emb"]) < 1e-5
        for _ in range(2):                               # second call replays the captured graph
            mel = net.decode(c["latents"].to(dev), emb)
        assert mel.shape == c["mel"].shape and mel.dtype == torch.float32
        assert rel_err(mel, c["mel"]) < BF16_NET, (tag, rel_err(mel, c["mel"]))
    assert tuple(net.get_mel_spec_shape((3, 8, 32, 688))) == tuple(g["mel_shape"])
    assert tuple(net.get_latent_shape((3, 2, 256, 5504))) == tuple(g["latent_shape"])


def test_dae_encode_vs_golden_reference(dev):
    """DAE_D3.encode (:342-354): conv_in (1,5,5) as a patch GEMM, 2-group encoder blocks, conv_latents_out, 2x2 pooling,
    latent normalisation; plus the autoencoder round trip forward() = (latents, reconstruction, pre-norm latents)."""
    spec = do.small_dae_spec()
    sd = do.synth_dae_state_dict(spec, seed=0)
    g = load_golden("dae_small.pt")
    net = make_dae(spec, sd, dev)
    e = g["encode"]
    lat = net.encode(e["mel"].to(dev), None)
    assert lat.shape == e["latents"].shape and lat.dtype == torch.float32
    assert rel_err(lat, e["latents"]) < BF16_NET, rel_err(lat, e["latents"])
    assert rel_err(net.encode(e["mel"].to(dev), None, training=True), e["pre_norm"]) < BF16_NET
    # stem patches: an index map (+ one bf16 rounding) -- bit-exact against an unfold of the reflect / zero padded input
    from dualdiffusion_b200 import ops
    mel = e["mel"]
    patches = ops.dae_enc_patches(mel.to(dev), PW).float().cpu()                           # [B][H][Wp][128]
    x = torch.stack((mel, torch.ones_like(mel)), dim=2)                                     # (B, z, c, H, W)
    xp = F.pad(F.pad(x.flatten(0, 1), (2, 2, 0, 0), mode="reflect"), (0, 0, 2, 2))          # reflect W, zero H
    un = F.unfold(xp, 5).view(mel.shape[0], 2, 2, 25, mel.shape[2], mel.shape[3])           # (B, z, c, tap, H, W)
    ref = un.permute(0, 4, 5, 1, 3, 2).reshape(mel.shape[0], mel.shape[2], mel.shape[3], 2, 50).to(torch.bfloat16).float()
    got = patches[:, :, PW:-PW].reshape(mel.shape[0], mel.shape[2], mel.shape[3], 2, 64)
    assert torch.equal(got[..., :50], ref) and got[..., 50:].abs().max().item() == 0
    emb = net.get_embeddings(g["cases"]["h8"]["emb_in"].repeat(2, 1))
    latents, recon, pre = net(e["mel"].to(dev), emb)
    assert rel_err(latents, e["latents"]) < BF16_NET and recon.shape == e["mel"].shape and pre.shape == lat.shape


def test_dae_decode_default_config_properties(dev):
    """Default 27 M-parameter decoder at a reduced width (full height: 32 latent rows -> 256 mel rows).  The CPU oracle
    checks a slice; stereo symmetry is a size-independent property: swapping the stereo sides of the latents swaps the
    sides of the output (the folded weights are block-circulant)."""
    spec = do.DAESpec()
    sd = do.synth_dae_state_dict(spec, seed=0)
    net = make_dae(spec, sd, dev)
    gen = torch.Generator().manual_seed(9)
    lat = uo.normalize(torch.randn(1, 8, 32, 24, generator=gen))
    emb_in = torch.randn(1, spec.in_channels_emb, generator=gen)
    emb = net.get_embeddings(emb_in)
    mel = net.decode(lat.to(dev), emb)
    assert mel.shape == (1, 2, 256, 192)
    ref = do.dae_decode(sd, spec, lat, do.dae_get_embeddings(sd, emb_in))
    assert rel_err(mel, ref) < BF16_NET, rel_err(mel, ref)
    swapped = lat.view(1, 4, 2, 32, 24).flip(2).reshape(1, 8, 32, 24)
    mel_s = net.decode(swapped.to(dev), emb)
    assert rel_err(mel_s, mel.flip(1)) < BF16_NET      # same arithmetic, different accumulation order over z (measured 1.4e-2)
