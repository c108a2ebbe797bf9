"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch fp32) of the newer latent UNet lineage
`modules/unets/unet_edm2_b4_2.py` of the reference: the b4 block recipe with
  * one fused q|k|v projection whose input is scaled by a single embedding gain, no activation on the attention
    output (Block.forward :121-163),
  * the noise level shifted before the Fourier embedding, c_noise = (ln sigma - offset) / 4, and a bandwidth factor
    in the embedding frequencies (:181, :258-259, :237-238),
  * 8-channel latents, 3 layers per level, mlp_multiplier 1, attention from level 2 (config :44-69).
Only tests/, smoke() and bench.py's CPU legs may import this module.  Parity is pinned against the unmodified reference
(tests/golden/make_golden_b4_2.py -> tests/golden/unet_b4_2_small.pt)."""
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import unet_oracle as uo


@dataclass
class UNetB42Spec(uo.UNetSpec):
    in_channels: int = 8
    out_channels: int = 8
    in_channels_emb: int = 1024
    sigma_max: float = 400.0
    sigma_min: float = 0.004
    sigma_data: float = 1.0
    mp_fourier_ln_sigma_offset: float = 0.5
    mp_fourier_bandwidth: float = 1.4
    logvar_channels: int = 192
    channel_mult: Sequence[int] = (2, 2, 3, 4, 5)
    num_layers_per_block: int = 3
    attn_levels: Sequence[int] = (2, 3, 4)
    mlp_multiplier: int = 1
    emb_linear_groups: int = 1


def small_spec() -> UNetB42Spec:
    """Reduced configuration that still has down / up sampling, unequal skip widths, grouped 3x3 convolutions and attention."""
    return UNetB42Spec(model_channels=128, channel_mult=(2, 4), channel_mult_noise=1, channel_mult_emb=2,
                       num_layers_per_block=1, attn_levels=(1,), in_channels_emb=64, logvar_channels=32)


def state_dict_shapes(spec: UNetB42Spec) -> Dict[str, Tuple[int, ...]]:
    shapes: Dict[str, Tuple[int, ...]] = {}
    cemb, g, m, eg = spec.cemb, spec.mlp_groups, spec.mlp_multiplier, spec.emb_linear_groups
    shapes["out_gain"] = ()
    for n, sh in (("emb_fourier.freqs", (spec.cnoise,)), ("emb_fourier.phases", (spec.cnoise,)),
                  ("emb_noise.weight", (cemb, spec.cnoise)), ("emb_label.weight", (cemb, spec.in_channels_emb)),
                  ("emb_label_unconditional.weight", (cemb, 1)), ("logvar_fourier.freqs", (spec.logvar_channels,)),
                  ("logvar_fourier.phases", (spec.logvar_channels,)), ("logvar_linear.weight", (1, spec.logvar_channels))):
        shapes[n] = sh
    enc, dec = uo.block_plan(spec)
    for bs in enc + dec:
        p = bs.name + "."
        if bs.kind == "conv":
            shapes[p + "weight"] = (bs.cout, bs.cin, 3, 3)
            continue
        res_in = bs.cout if bs.flavor == "enc" else bs.cin
        shapes[p + "emb_gain"] = ()
        shapes[p + "conv_res0.weight"] = (bs.cout * m, res_in // g, 3, 3)
        shapes[p + "conv_res1.weight"] = (bs.cout, bs.cout * m // g, 3, 3)
        shapes[p + "conv_skip.weight"] = (bs.cout, bs.cin, 1, 1)
        shapes[p + "emb_linear.weight"] = (bs.cout * m, cemb // eg, 1, 1)
        if bs.attention:
            shapes[p + "emb_gain_qkv"] = ()
            shapes[p + "emb_linear_qkv.weight"] = (bs.cout, cemb // eg, 1, 1)
            shapes[p + "attn_qkv.weight"] = (bs.cout * 3, bs.cout, 1, 1)
            shapes[p + "attn_proj.weight"] = (bs.cout, bs.cout, 1, 1)
    shapes["conv_out.weight"] = (spec.out_channels, spec.cblock[0], 3, 3)
    return shapes


def synth_state_dict(spec: UNetB42Spec, seed: int = 0, gain: float = 0.5) -> Dict[str, Tensor]:
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in sorted(state_dict_shapes(spec).items()):
        if name.endswith(".freqs") or name.endswith(".phases"):
            continue
        if shape == ():
            sd[name] = torch.tensor(gain)
        else:
            w = torch.randn(shape, generator=gen)
            sd[name] = w if name == "logvar_linear.weight" else uo.normalize(w)
    sd["emb_fourier.freqs"], sd["emb_fourier.phases"] = uo.mp_fourier_buffers(spec.cnoise, spec.mp_fourier_bandwidth)
    sd["logvar_fourier.freqs"], sd["logvar_fourier.phases"] = uo.mp_fourier_buffers(spec.logvar_channels)
    return sd


def attention_qkv(qkv: Tensor, heads: int) -> Tensor:
    """unet_edm2_b4_2.py:148-156: channel index (head, d, {q,k,v}); cosine-normalise over d; SDPA over the H*W tokens."""
    b, _, h, w = qkv.shape
    x = qkv.reshape(b, heads, -1, 3, h * w)
    q, k, v = uo.normalize(x, dim=2).unbind(3)
    d = q.shape[2]
    s = torch.einsum("bhdq,bhdk->bhqk", q.float(), k.float()) / math.sqrt(d)
    y = torch.einsum("bhqk,bhdk->bhdq", torch.softmax(s, dim=-1), v.float())
    return y.reshape(b, -1, h, w)


def block_forward(sd, bs, spec: UNetB42Spec, x: Tensor, emb: Tensor, training: bool = False) -> Tensor:
    p = bs.name + "."
    g, eg = spec.mlp_groups, spec.emb_linear_groups

    def conv(name, inp, gain=1.0, groups=1):
        return uo.mp_conv(inp, sd[p + name + ".weight"], gain, groups, training)

    x = uo.resample_2d(x, bs.resample)
    if bs.flavor == "enc":
        x = conv("conv_skip", x)
        x = uo.normalize(x, dim=1)
    y = conv("conv_res0", uo.mp_silu(x), groups=g)
    c = conv("emb_linear", emb, gain=sd[p + "emb_gain"], groups=eg) + 1.0
    y = uo.mp_silu(y * c)
    y = conv("conv_res1", y, groups=g)
    if bs.flavor == "dec":
        x = conv("conv_skip", x)
    x = uo.mp_sum(x, y, spec.res_balance)
    if bs.attention:
        heads = bs.cout // spec.channels_per_head
        c = conv("emb_linear_qkv", emb, gain=sd[p + "emb_gain_qkv"], groups=eg) + 1.0
        y = attention_qkv(conv("attn_qkv", x * c), heads)
        y = conv("attn_proj", y)
        x = uo.mp_sum(x, y, spec.attn_balance)
    return x.clip(-256.0, 256.0)


def get_embeddings(sd, emb_in: Tensor, conditioning_mask: Tensor) -> Tensor:
    return uo.get_embeddings(sd, emb_in, conditioning_mask)


def sigma_loss_logvar(sd, spec: UNetB42Spec, sigma: Tensor) -> Tensor:
    """unet_edm2_b4_2.py:237-238."""
    ln_sigma = sigma.flatten().float().log() - spec.mp_fourier_ln_sigma_offset
    f = uo.mp_fourier(ln_sigma / 4, sd["logvar_fourier.freqs"], sd["logvar_fourier.phases"])
    return uo.mp_conv(f, sd["logvar_linear.weight"]).view(-1, 1, 1, 1).float()


def unet_forward(sd, spec: UNetB42Spec, x_in: Tensor, sigma: Tensor, embeddings: Tensor, x_ref: Optional[Tensor] = None,
                 training: bool = False) -> Tensor:
    """unet_edm2_b4_2.py:250-305, all arithmetic in fp32."""
    x_in = x_in.float()
    sigma = sigma.float().view(-1, 1, 1, 1)
    sd2 = spec.sigma_data ** 2
    c_skip = sd2 / (sigma ** 2 + sd2)
    c_out = sigma * spec.sigma_data / (sigma ** 2 + sd2).sqrt()
    c_in = 1 / (sd2 + sigma ** 2).sqrt()
    c_noise = (sigma.flatten().log() - spec.mp_fourier_ln_sigma_offset) / 4
    emb = uo.mp_conv(uo.mp_fourier(c_noise, sd["emb_fourier.freqs"], sd["emb_fourier.phases"]), sd["emb_noise.weight"],
                     training=training)
    emb = uo.mp_silu(uo.mp_sum(emb, embeddings.float(), spec.label_balance))[:, :, None, None]
    x = c_in * x_in
    b, _, h, w = x.shape
    x = torch.cat((x, torch.ones_like(x[:, :1]), uo.ln_freqs_channel(spec, b, h, w)), dim=1)
    enc, dec = uo.block_plan(spec)
    skips: List[Tensor] = []
    for bs in enc:
        x = uo.mp_conv(x, sd[bs.name + ".weight"], training=training) if bs.kind == "conv" else \
            block_forward(sd, bs, spec, x, emb, training)
        skips.append(x)
    for bs in dec:
        if bs.takes_skip:
            x = uo.mp_cat(x, skips.pop(), spec.concat_balance)
        x = block_forward(sd, bs, spec, x, emb, training)
    x = uo.mp_conv(x, sd["conv_out.weight"], gain=sd["out_gain"], training=training)
    d = c_skip * x_in + c_out * x
    if x_ref is not None:
        d = uo.mp_sum(x_ref[:, :-1].float(), d, x_ref[:, -1:].float())
    return d
