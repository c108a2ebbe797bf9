"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch, fp32) of the reference's 2-D diffusion autoencoder
`modules.daes.dae_edm2_q4.DAE` (/root/reference/src/modules/daes/dae_edm2_q4.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file; the product path never does.

Parity status: PINNED against the unmodified reference (tests/golden/make_golden_dae_q4.py -> dae_q4_small.pt;
tests/test_oracle.py::test_dae_q4_oracle_matches_reference_golden).
Eval mode (no weight normalisation inside the forward, mp_tools.py:360), in_channels_emb = 0 (the dataclass default: no
embedding, emb_linear is None), no attention (the reference raises, :163), no dropout, add_pixel_norm either way.
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
class DAEQ4Spec:
    """dae_edm2_q4.py:92-113 (dataclass defaults)."""
    in_channels: int = 2
    out_channels: int = 2
    latent_channels: int = 8
    model_channels: int = 64
    channel_mult_enc: Sequence[int] = (1, 2, 4, 8)
    channel_mult_dec: Sequence[int] = (1, 2, 4, 8)
    num_enc_layers_per_block: int = 3
    num_dec_layers_per_block: int = 3
    res_balance: float = 0.3
    mlp_multiplier: int = 2
    add_pixel_norm: bool = False

    @property
    def num_levels(self) -> int:
        return len(self.channel_mult_dec)


def small_dae_q4_spec() -> DAEQ4Spec:
    return DAEQ4Spec(model_channels=32, channel_mult_enc=(1, 2, 3), channel_mult_dec=(1, 2, 2), num_enc_layers_per_block=1,
                     num_dec_layers_per_block=2)


def dae_q4_block_plan(spec: DAEQ4Spec) -> Tuple[List[tuple], List[tuple]]:
    """(enc, dec): (name, cin, cout, flavor, resample) in execution order (:225-262)."""
    enc_ch = [spec.model_channels * m for m in spec.channel_mult_enc]
    dec_ch = [spec.model_channels * m for m in spec.channel_mult_dec]
    enc, dec = [], []
    cin = enc_ch[0]
    for level in range(spec.num_levels):
        cout = enc_ch[level]
        if level > 0:
            enc.append((f"enc.block{level}_down", cin, cout, "enc", "down"))
        for idx in range(spec.num_enc_layers_per_block):
            enc.append((f"enc.block{level}_layer{idx}", cout, cout, "enc", "keep"))
        cin = cout
    cin = dec_ch[-1]
    for level in reversed(range(spec.num_levels)):
        cout = dec_ch[level]
        if level == spec.num_levels - 1:
            dec.append((f"dec.block{level}_in0", cin, cout, "dec", "keep"))
        else:
            dec.append((f"dec.block{level}_up", cin, cout, "dec", "up"))
        for idx in range(spec.num_dec_layers_per_block):
            dec.append((f"dec.block{level}_layer{idx}", cout, cout, "dec", "keep"))
        cin = cout
    return enc, dec


def mp_conv2d(x: Tensor, w: Tensor, gain=1.0, bias=None) -> Tensor:
    """MPConv.forward, eval mode (mp_tools.py:357-373)."""
    w = (w.float() * (gain / math.sqrt(w[0].numel()))).to(x.dtype)
    y = F.conv2d(x, w, padding=(w.shape[-2] // 2, w.shape[-1] // 2))
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1).to(x.dtype)
    return y


def dae_q4_block_forward(sd: Dict[str, Tensor], name: str, flavor: str, resample: str, spec: DAEQ4Spec, x: Tensor) -> Tensor:
    """Block.forward (:165-198) with emb_linear None, dropout 0, mlp_groups 1."""
    p = name + "."
    if resample == "down":
        x = F.avg_pool2d(x, 2)                                              # mp_tools.py:77
    elif resample == "up":
        x = F.interpolate(x, scale_factor=2, mode="nearest")                # mp_tools.py:79
    if flavor == "enc":
        if p + "conv_skip.weight" in sd:
            x = mp_conv2d(x, sd[p + "conv_skip.weight"])
        if spec.add_pixel_norm:
            x = normalize(x, dim=1)                                         # normalize_groups, groups = 1 (mp_tools.py:53-54)
    y = mp_conv2d(x, sd[p + "conv_res0.weight"])
    y = mp_silu(normalize(y, dim=1))
    y = mp_conv2d(y, sd[p + "conv_res1.weight"])
    if flavor == "dec" and p + "conv_skip.weight" in sd:
        x = mp_conv2d(x, sd[p + "conv_skip.weight"])
    return mp_sum(x, y, spec.res_balance).clip(-256.0, 256.0)


def dae_q4_encode(sd: Dict[str, Tensor], spec: DAEQ4Spec, x: Tensor) -> Tensor:
    """DAE.encode (:270-283): mel-spectrogram (B, 2, H, W) -> latents (B, latent_channels, H/r, W/r)."""
    x = mp_conv2d(x.float(), sd["enc.conv_in.weight"], bias=sd["enc.conv_in.bias"])
    for name, cin, cout, flavor, resample in dae_q4_block_plan(spec)[0]:
        x = dae_q4_block_forward(sd, name, flavor, resample, spec, x)
    return mp_conv2d(x, sd["conv_latents_out.weight"])


def dae_q4_decode(sd: Dict[str, Tensor], spec: DAEQ4Spec, latents: Tensor) -> Tensor:
    """DAE.decode (:285-299): latents -> mel-spectrogram (B, 2, H*r, W*r)."""
    x = mp_conv2d(latents.float(), sd["conv_latents_in.weight"], bias=sd["conv_latents_in.bias"])
    for name, cin, cout, flavor, resample in dae_q4_block_plan(spec)[1]:
        x = dae_q4_block_forward(sd, name, flavor, resample, spec, x)
    return mp_conv2d(x, sd["conv_out.weight"], gain=sd["out_gain"])


def dae_q4_state_dict_shapes(spec: DAEQ4Spec) -> Dict[str, Tuple[int, ...]]:
    shapes: Dict[str, Tuple[int, ...]] = {"out_gain": (), "recon_loss_logvar": ()}
    L = spec.latent_channels
    shapes["latents_stats_tracker.mean"] = shapes["latents_stats_tracker.var"] = (L,)
    shapes["latents_stats_tracker.global_mean"] = shapes["latents_stats_tracker.global_var"] = (1,)
    enc_ch = [spec.model_channels * m for m in spec.channel_mult_enc]
    dec_ch = [spec.model_channels * m for m in spec.channel_mult_dec]
    shapes["enc.conv_in.weight"] = (enc_ch[0], spec.in_channels, 5, 5)
    shapes["enc.conv_in.bias"] = (enc_ch[0],)
    shapes["conv_latents_out.weight"] = (L, enc_ch[-1], 3, 3)
    shapes["conv_latents_in.weight"] = (dec_ch[-1], L, 3, 3)
    shapes["conv_latents_in.bias"] = (dec_ch[-1],)
    shapes["conv_out.weight"] = (spec.out_channels, dec_ch[0], 5, 5)
    m = spec.mlp_multiplier
    enc, dec = dae_q4_block_plan(spec)
    for name, cin, cout, flavor, resample in enc + dec:
        p = name + "."
        shapes[p + "conv_res0.weight"] = (cout * m, cout if flavor == "enc" else cin, 3, 3)
        shapes[p + "conv_res1.weight"] = (cout, cout * m, 3, 3)
        if cin != cout:
            shapes[p + "conv_skip.weight"] = (cout, cin, 1, 1)
    return shapes


def synth_dae_q4_state_dict(spec: DAEQ4Spec, seed: int = 0, gain: float = 0.8) -> Dict[str, Tensor]:
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in sorted(dae_q4_state_dict_shapes(spec).items()):
        if name.startswith("latents_stats_tracker."):
            sd[name] = torch.ones(shape) if name.endswith("var") else torch.zeros(shape)
        elif name == "out_gain":
            sd[name] = torch.tensor(gain)
        elif shape == ():
            sd[name] = torch.tensor(0.0)
        elif name.endswith(".bias"):
            sd[name] = 0.3 * torch.randn(shape, generator=gen)
        else:
            sd[name] = normalize(torch.randn(shape, generator=gen))
    return sd
