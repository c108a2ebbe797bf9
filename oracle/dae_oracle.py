"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch, fp32) of the reference's DAE_D3 diffusion-autoencoder
*decoder* (/root/reference/src/modules/daes/dae_edm2_d3.py).  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this file; the product path never does.

Parity status: PINNED against the reference itself (tests/golden/make_golden.py -> dae_small.pt; checked in
tests/test_oracle.py).  Functional restatement over a reference-layout state_dict (same keys / shapes as
`modules.daes.dae_edm2_d3.DAE_D3`), eval mode (no weight normalisation inside the forward, mp_tools/MPConv3D :75-80).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

from .unet_oracle import mp_silu, mp_sum, normalize

Tensor = torch.Tensor


@dataclass
class DAESpec:
    """dae_edm2_d3.py:98-121 (decoder-relevant fields; defaults of config/models/edm2_ddec_mclt_b1a/dae.json)."""
    in_channels_emb: int = 1024
    latent_channels: int = 4
    model_channels: int = 32
    channel_mult_enc: int = 1
    channel_mult_dec: Sequence[int] = (1, 2, 4, 8)
    channel_mult_emb: int = 4
    num_enc_layers: int = 6
    num_dec_layers_per_block: int = 3
    res_balance: float = 0.3
    mlp_multiplier: int = 2
    add_constant_channel: bool = True

    @property
    def cemb(self) -> int:
        return self.model_channels * self.channel_mult_emb * self.mlp_multiplier

    @property
    def dec_channels(self) -> List[int]:
        return [self.model_channels * m for m in self.channel_mult_dec]


def small_dae_spec() -> DAESpec:
    """Reduced decoder (two levels, one layer per level) that still has an up block, a channel-changing conv_skip and
    channel-preserving blocks without one."""
    return DAESpec(in_channels_emb=64, model_channels=32, channel_mult_dec=(1, 2), channel_mult_emb=2,
                   num_enc_layers=1, num_dec_layers_per_block=1)


def dec_block_plan(spec: DAESpec) -> List[Tuple[str, int, int, bool]]:
    """(name, cin, cout, upsample) of the decoder blocks in execution order (dae_edm2_d3.py:292-310)."""
    plan = []
    ch = spec.dec_channels
    cin = ch[-1]
    for level in reversed(range(len(ch))):
        cout = ch[level]
        if level == len(ch) - 1:
            plan.append((f"dec.block{level}_in0", cin, cout, False))
        else:
            plan.append((f"dec.block{level}_up", cin, cout, True))
        for idx in range(spec.num_dec_layers_per_block):
            plan.append((f"dec.block{level}_layer{idx}", cout, cout, False))
        cin = cout
    return plan


def mp_conv3d(x: Tensor, w: Tensor, gain=1.0) -> Tensor:
    """MPConv3D.forward, eval mode (dae_edm2_d3.py:70-86): scale by gain/sqrt(fan_in); reflection padding along W
    (both sides) and along Z (one-sided, at the end), zero padding along H."""
    w = w.float() * (gain / math.sqrt(w[0].numel()))
    if w.ndim == 2:
        return x @ w.t()
    kz, kh, kw = w.shape[2:]
    if kz // 2 or kw // 2:
        x = F.pad(x, (kw // 2, kw // 2, 0, 0, 0, kz // 2), mode="reflect")
    return F.conv3d(x, w, padding=(0, kh // 2, 0))


def dae_get_embeddings(sd: Dict[str, Tensor], emb_in: Tensor) -> Tensor:
    """DAE_D3.get_embeddings (:314-318)."""
    return mp_conv3d(normalize(emb_in.float()), sd["emb_label.weight"])


def dae_block_forward(sd: Dict[str, Tensor], name: str, spec: DAESpec, x: Tensor, emb: Tensor, up: bool) -> Tensor:
    """Block.forward, flavor "dec", no attention, no dropout (:186-238)."""
    p = name + "."
    if up:
        x = x.repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2)        # resample_3d "up" (mp_tools.py:92-93)
    y = mp_conv3d(mp_silu(x), sd[p + "conv_res0.weight"])
    c = mp_conv3d(emb[:, :, None, None, None], sd[p + "emb_linear.weight"], gain=sd[p + "emb_gain"]) + 1.0
    y = mp_silu(y * c)
    y = mp_conv3d(y, sd[p + "conv_res1.weight"])
    if p + "conv_skip.weight" in sd:
        x = mp_conv3d(x, sd[p + "conv_skip.weight"])
    return mp_sum(x, y, spec.res_balance).clip(-256.0, 256.0)


def dae_decode(sd: Dict[str, Tensor], spec: DAESpec, latents: Tensor, embeddings: Tensor) -> Tensor:
    """DAE_D3.decode (:356-369): latents (B, 2*latent_channels, H, W) -> mel-spectrogram (B, 2, H*r, W*r)."""
    b, _, h, w = latents.shape
    x = latents.float().reshape(b, spec.latent_channels, -1, h, w)              # tensor_4d_to_5d
    x = torch.cat((x, torch.ones_like(x[:, :1])), dim=1)
    x = mp_conv3d(x, sd["conv_latents_in.weight"])
    for name, _, _, up in dec_block_plan(spec):
        x = dae_block_forward(sd, name, spec, x, embeddings.float(), up)
    out = mp_conv3d(x, sd["conv_out.weight"], gain=sd["out_gain"])
    return out.reshape(b, out.shape[1] * out.shape[2], out.shape[3], out.shape[4])   # tensor_5d_to_4d


def dae_state_dict_shapes(spec: DAESpec) -> Dict[str, Tuple[int, ...]]:
    """Every parameter of DAE_D3 (encoder included, so a strict load_state_dict of the product module works)."""
    shapes: Dict[str, Tuple[int, ...]] = {"out_gain": (), "recon_loss_logvar": ()}
    cemb = spec.cemb
    shapes["emb_label.weight"] = (cemb, spec.in_channels_emb)
    enc = spec.model_channels * spec.channel_mult_enc
    shapes["enc.conv_in.weight"] = (enc, 1 + int(spec.add_constant_channel), 1, 5, 5)
    m = spec.mlp_multiplier
    for i in range(spec.num_enc_layers):
        p = f"enc.block0_layer{i}."
        shapes[p + "conv_res0.weight"] = (enc * m, enc, 1, 3, 3)
        shapes[p + "conv_res1.weight"] = (enc, enc * m, 1, 3, 3)
        shapes[p + "emb_gain"] = ()
    shapes["conv_latents_out.weight"] = (spec.latent_channels, enc, 2, 3, 3)
    shapes["conv_latents_in.weight"] = (spec.dec_channels[-1], spec.latent_channels + int(spec.add_constant_channel), 2, 3, 3)
    for name, cin, cout, _ in dec_block_plan(spec):
        p = name + "."
        shapes[p + "conv_res0.weight"] = (cout * m, cin, 2, 3, 3)
        shapes[p + "conv_res1.weight"] = (cout, cout * m, 2, 3, 3)
        if cin != cout:
            shapes[p + "conv_skip.weight"] = (cout, cin, 1, 1, 1)
        shapes[p + "emb_gain"] = ()
        shapes[p + "emb_linear.weight"] = (cout * m, cemb, 1, 1, 1)
    shapes["conv_out.weight"] = (1, spec.dec_channels[0], 1, 5, 5)
    return shapes


def synth_dae_state_dict(spec: DAESpec, seed: int = 0, gain: float = 0.5) -> Dict[str, Tensor]:
    """Seeded CPU weights in reference layout: randn followed by the post-load normalize_weights() of
    MPConv3D (:88-94, norm_dim = 1); scalar gains non-zero so that the embedding path contributes."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in sorted(dae_state_dict_shapes(spec).items()):
        if shape == ():
            sd[name] = torch.tensor(1.0 if name == "out_gain" else (0.0 if name == "recon_loss_logvar" else gain))
        else:
            sd[name] = normalize(torch.randn(shape, generator=gen), dim=1)
    return sd


def dae_encode(sd: Dict[str, Tensor], spec: DAESpec, mel: Tensor, training: bool = False) -> Tensor:
    """DAE_D3.encode (:342-354) with Block.forward flavor "enc" (:186-238; encoder blocks have no embedding, no conv_skip,
    no pixel norm in the shipped configuration): mel (B, 2, H, W) -> latents (B, 2*latent_channels, H/r, W/r)."""
    b, _, h, w = mel.shape
    x = mel.float().reshape(b, 1, -1, h, w)                                       # tensor_4d_to_5d(x, 1)
    x = torch.cat((x, torch.ones_like(x[:, :1])), dim=1)
    x = mp_conv3d(x, sd["enc.conv_in.weight"])
    for i in range(spec.num_enc_layers):
        p = f"enc.block0_layer{i}."
        y = mp_conv3d(mp_silu(x), sd[p + "conv_res0.weight"])
        y = mp_silu(y)
        y = mp_conv3d(y, sd[p + "conv_res1.weight"])
        x = mp_sum(x, y, spec.res_balance).clip(-256.0, 256.0)
    lat = mp_conv3d(x, sd["conv_latents_out.weight"])
    lat = lat.reshape(b, lat.shape[1] * lat.shape[2], lat.shape[3], lat.shape[4])  # tensor_5d_to_4d
    lat = F.avg_pool2d(lat, 2 ** (len(spec.channel_mult_dec) - 1))
    return lat if training else normalize(lat)
