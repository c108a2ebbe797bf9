"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch) of the reference's diffusion-decoder UNet
`DDec_MCLT_UNet_B1.forward` (/root/reference/src/modules/unets/unet_edm2_ddec_mclt_b1.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file; the product path never does.

Parity status: PINNED against the reference itself (tests/golden/make_golden.py -> ddec_small.pt; tests/test_oracle.py).
The reference hard-codes bfloat16 casts of the network input, x_ref and the embedding (:295,298,306), so its body runs
in bf16 even on the CPU; `body_dtype=torch.bfloat16` reproduces that statement by statement (used for pinning),
`body_dtype=torch.float32` is the same arithmetic without the rounding (the yardstick for the CUDA path).
Eval mode (no weight normalisation inside the forward), no attention, no dropout, in_channels_emb = 0.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

from .unet_oracle import mp_cat, mp_fourier, mp_fourier_buffers, mp_silu, mp_sum, normalize

Tensor = torch.Tensor


@dataclass
class DDecSpec:
    """unet_edm2_ddec_mclt_b1.py:45-73 (defaults of config/models/edm2_ddec_mclt_b1a/ddec.json)."""
    in_num_freqs: int = 256
    in_psd_freqs: int = 4096
    model_channels: int = 32
    logvar_channels: int = 128
    channel_mult: Sequence[int] = (1, 2, 3, 4)
    double_midblock: bool = True
    channel_mult_noise: int = 4
    channel_mult_emb: int = 4
    num_layers_per_block: int = 3
    concat_balance: float = 0.5
    res_balance: float = 0.3
    mlp_multiplier: int = 2
    sigma_data: float = 1.0

    @property
    def cblock(self) -> List[int]:
        return [self.model_channels * m for m in self.channel_mult]

    @property
    def cnoise(self) -> int:
        return self.model_channels * self.channel_mult_noise

    @property
    def cemb(self) -> int:
        return self.model_channels * self.channel_mult_emb * self.mlp_multiplier

    @property
    def psd_per_freq(self) -> int:
        return self.in_psd_freqs // self.in_num_freqs


def small_ddec_spec() -> DDecSpec:
    """Two levels, one layer per level, 32 mel rows x 4 PSD bins per row: every block flavour (enc, enc+down, mid x2,
    dec+up, dec with skip concat of unequal widths) at CPU-friendly size."""
    return DDecSpec(in_num_freqs=32, in_psd_freqs=128, channel_mult=(1, 2), num_layers_per_block=1, logvar_channels=32)


def ddec_block_plan(spec: DDecSpec):
    """(enc, dec) lists of (name, kind, cin, cout, resample, takes_skip) in execution order (:217-261)."""
    cblock = spec.cblock
    enc, dec = [], []
    cout = 1 + spec.psd_per_freq + 1
    for level, ch in enumerate(cblock):
        if level == 0:
            enc.append(("enc.conv_in", "conv", cout, ch, "keep", False))
            cout = ch
        else:
            enc.append((f"enc.block{level}_down", "block", cout, cout, "down", False))
        for idx in range(spec.num_layers_per_block):
            enc.append((f"enc.block{level}_layer{idx}", "block", cout, ch, "keep", False))
            cout = ch
    skips = [e[3] for e in enc]
    for level, ch in reversed(list(enumerate(cblock))):
        if level == len(cblock) - 1:
            dec.append((f"dec.block{level}_in0", "block", cout, cout, "keep", False))
            if spec.double_midblock:
                dec.append((f"dec.block{level}_in1", "block", cout, cout, "keep", False))
        else:
            dec.append((f"dec.block{level}_up", "block", cout, cout, "up", False))
        for idx in range(spec.num_layers_per_block + 1):
            cin = cout + skips.pop()
            dec.append((f"dec.block{level}_layer{idx}", "block", cin, ch, "keep", True))
            cout = ch
    return enc, dec, cout


def mp_conv3d(x: Tensor, w: Tensor, gain=1.0) -> Tensor:
    """MPConv3D.forward, eval mode (dae_edm2_d3.py:70-86), weights cast to the activation dtype (:81)."""
    w = w.float() * (gain / math.sqrt(w[0].numel()))
    w = w.to(x.dtype)
    if w.ndim == 2:
        return x @ w.t()
    kz, kh, kw = w.shape[2:]
    if kz // 2 or kw // 2:
        x = F.pad(x, (kw // 2, kw // 2, 0, 0, 0, kz // 2), mode="reflect")
    return F.conv3d(x, w, padding=(0, kh // 2, 0))


def resample_3d(x: Tensor, mode: str) -> Tensor:
    """mp_tools.py:81-93."""
    if mode == "keep":
        return x
    if mode == "down":
        s = x.shape
        return F.avg_pool2d(x.reshape(s[0] * s[1], s[2], s[3], s[4]), 2).view(s[0], s[1], s[2], s[3] // 2, s[4] // 2)
    return x.repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2)


def ddec_block_forward(sd, name: str, flavor: str, resample: str, spec: DDecSpec, x: Tensor, emb: Tensor) -> Tensor:
    """Block.forward (:127-175), no attention."""
    p = name + "."
    x = resample_3d(x, resample)
    if flavor == "enc":
        x = mp_conv3d(x, sd[p + "conv_skip.weight"])
        x = normalize(x, dim=1)
    y = mp_conv3d(mp_silu(x), sd[p + "conv_res0.weight"])
    c = mp_conv3d(emb, sd[p + "emb_linear.weight"], gain=sd[p + "emb_gain"]) + 1.0
    y = mp_silu(y * c)
    y = mp_conv3d(y, sd[p + "conv_res1.weight"])
    if flavor == "dec":
        x = mp_conv3d(x, sd[p + "conv_skip.weight"])
    return mp_sum(x, y, spec.res_balance).clip(-256.0, 256.0)


def ddec_forward(sd: Dict[str, Tensor], spec: DDecSpec, x_in: Tensor, sigma: Tensor, x_ref: Tensor,
                 body_dtype=torch.float32) -> Tensor:
    """DDec_MCLT_UNet_B1.forward (:278-326): x_in (B, 2, F, W) MDCT image, x_ref (B, 2, F*k, W) PSD -> D_x (B, 2, F, W)."""
    sig = sigma.float().view(-1, 1, 1, 1, 1)
    sd2 = spec.sigma_data ** 2
    c_skip = sd2 / (sig ** 2 + sd2)
    c_out = sig * spec.sigma_data / (sig ** 2 + sd2).sqrt()
    c_in = 1 / (sd2 + sig ** 2).sqrt()
    c_noise = sig.flatten().log() / 4
    b, _, f, w = x_in.shape
    k = spec.psd_per_freq
    xr = x_ref.view(b, x_ref.shape[1], spec.in_num_freqs, k, x_ref.shape[3]).permute(0, 3, 1, 2, 4).to(body_dtype)
    x = (c_in * x_in.float().reshape(b, 1, -1, f, w)).to(body_dtype)                     # tensor_4d_to_5d(x_in, 1)
    emb = mp_conv3d(mp_fourier(c_noise, sd["emb_fourier.freqs"], sd["emb_fourier.phases"]), sd["emb_noise.weight"])
    emb = emb[:, :, None, None, None].to(body_dtype)
    x = torch.cat((x, xr, torch.ones_like(x[:, :1])), dim=1)
    enc, dec, _ = ddec_block_plan(spec)
    skips = []
    for name, kind, cin, cout, resample, _ in enc:
        if kind == "conv":
            x = mp_conv3d(x, sd[name + ".weight"])
        else:
            x = ddec_block_forward(sd, name, "enc", resample, spec, x, emb)
        skips.append(x)
    for name, kind, cin, cout, resample, takes_skip in dec:
        if takes_skip:
            x = mp_cat(x, skips.pop(), spec.concat_balance)
        x = ddec_block_forward(sd, name, "dec", resample, spec, x, emb)
    x = mp_conv3d(x, sd["conv_out.weight"], gain=sd["out_gain"])
    d = c_skip * x_in.float().unsqueeze(1) + c_out * x.float()
    return d.reshape(b, d.shape[1] * d.shape[2], d.shape[3], d.shape[4])                 # tensor_5d_to_4d


def ddec_sigma_loss_logvar(sd: Dict[str, Tensor], sigma: Tensor) -> Tensor:
    """:268-269."""
    fo = mp_fourier(sigma.flatten().float().log() / 4, sd["logvar_fourier.freqs"], sd["logvar_fourier.phases"])
    return mp_conv3d(fo, sd["logvar_linear.weight"]).view(-1, 1, 1, 1).float()


def ddec_state_dict_shapes(spec: DDecSpec) -> Dict[str, Tuple[int, ...]]:
    shapes: Dict[str, Tuple[int, ...]] = {"out_gain": ()}
    shapes["emb_fourier.freqs"] = shapes["emb_fourier.phases"] = (spec.cnoise,)
    shapes["emb_noise.weight"] = (spec.cemb, spec.cnoise)
    shapes["logvar_fourier.freqs"] = shapes["logvar_fourier.phases"] = (spec.logvar_channels,)
    shapes["logvar_linear.weight"] = (1, spec.logvar_channels)
    enc, dec, cout = ddec_block_plan(spec)
    m = spec.mlp_multiplier
    for name, kind, cin, co, resample, _ in enc + dec:
        if kind == "conv":
            shapes[name + ".weight"] = (co, cin, 2, 3, 3)
            continue
        p = name + "."
        flavor = "enc" if name.startswith("enc") else "dec"
        shapes[p + "conv_res0.weight"] = (co * m, co if flavor == "enc" else cin, 1, 3, 3)
        shapes[p + "conv_res1.weight"] = (co, co * m, 1, 3, 3)
        shapes[p + "conv_skip.weight"] = (co, cin, 2, 1, 1)
        shapes[p + "emb_gain"] = ()
        shapes[p + "emb_linear.weight"] = (co * m, spec.cemb, 1, 1, 1)
    shapes["conv_out.weight"] = (1, cout, 2, 3, 3)
    return shapes


def synth_ddec_state_dict(spec: DDecSpec, seed: int = 0, gain: float = 0.5) -> Dict[str, Tensor]:
    """Seeded CPU weights in reference layout: randn + the post-load normalize_weights() of MPConv3D (norm_dim = 1);
    scalar gains non-zero so that the embedding path and the output head contribute."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in sorted(ddec_state_dict_shapes(spec).items()):
        if name.endswith(".freqs") or name.endswith(".phases"):
            continue
        if shape == ():
            sd[name] = torch.tensor(gain)
        else:
            w = torch.randn(shape, generator=gen)
            sd[name] = w if name == "logvar_linear.weight" else normalize(w, dim=1)
    sd["emb_fourier.freqs"], sd["emb_fourier.phases"] = mp_fourier_buffers(spec.cnoise)
    sd["logvar_fourier.freqs"], sd["logvar_fourier.phases"] = mp_fourier_buffers(spec.logvar_channels)
    return sd


# ======================================================================================================================
# unet_edm2_q4_ddec.UNet (/root/reference/src/modules/unets/unet_edm2_q4_ddec.py): the 2-D sibling of the decoder above.
# Stereo is a channel pair, convolutions are plain zero-padded MPConv (mp_tools.py:332-378; conv_in carries a bias), the
# PSD reference enters as 2k extra channels through mp_cat (:268-277).
# ======================================================================================================================
@dataclass
class Q4Spec:
    """unet_edm2_q4_ddec.py:44-70 (dataclass defaults; no configuration file ships for this model)."""
    in_channels: int = 2
    out_channels: int = 2
    in_num_freqs: int = 256
    in_psd_freqs: int = 2048
    model_channels: int = 32
    logvar_channels: int = 192
    channel_mult: Sequence[int] = (1, 2, 3, 4, 5)
    double_midblock: bool = True
    channel_mult_noise: int = 4
    channel_mult_emb: int = 4
    num_layers_per_block: int = 3
    label_balance: float = 0.5
    concat_balance: float = 0.5
    res_balance: float = 0.3
    mlp_multiplier: int = 2
    sigma_data: float = 1.0

    @property
    def cblock(self) -> List[int]:
        return [self.model_channels * m for m in self.channel_mult]

    @property
    def cnoise(self) -> int:
        return self.model_channels * self.channel_mult_noise

    @property
    def cemb(self) -> int:
        return self.model_channels * self.channel_mult_emb * self.mlp_multiplier

    @property
    def psd_per_freq(self) -> int:
        return self.in_psd_freqs // self.in_num_freqs


def small_q4_spec() -> Q4Spec:
    return Q4Spec(in_num_freqs=32, in_psd_freqs=128, channel_mult=(1, 2), num_layers_per_block=1, logvar_channels=32)


def q4_block_plan(spec: Q4Spec):
    """(enc, dec, cout): (name, kind, cin, cout, resample, takes_skip) in execution order (:190-236)."""
    cblock = spec.cblock
    enc, dec = [], []
    cout = spec.in_channels + spec.psd_per_freq * 2
    for level, ch in enumerate(cblock):
        if level == 0:
            enc.append(("enc.conv_in", "conv", cout, ch, "keep", False))
            cout = ch
        else:
            enc.append((f"enc.block{level}_down", "block", cout, cout, "down", False))
        for idx in range(spec.num_layers_per_block):
            enc.append((f"enc.block{level}_layer{idx}", "block", cout, ch, "keep", False))
            cout = ch
    skips = [e[3] for e in enc]
    for level, ch in reversed(list(enumerate(cblock))):
        if level == len(cblock) - 1:
            dec.append((f"dec.block{level}_in0", "block", cout, cout, "keep", False))
            if spec.double_midblock:
                dec.append((f"dec.block{level}_in1", "block", cout, cout, "keep", False))
        else:
            dec.append((f"dec.block{level}_up", "block", cout, cout, "up", False))
        for idx in range(spec.num_layers_per_block + 1):
            cin = cout + skips.pop()
            dec.append((f"dec.block{level}_layer{idx}", "block", cin, ch, "keep", True))
            cout = ch
    return enc, dec, cout


def mp_conv2d(x: Tensor, w: Tensor, gain=1.0, bias=None) -> Tensor:
    """MPConv.forward, eval mode (mp_tools.py:357-373), weights / bias cast to the activation dtype."""
    w = (w.float() * (gain / math.sqrt(w[0].numel()))).to(x.dtype)
    if w.ndim == 2:
        return x @ w.t()
    y = F.conv2d(x, w, padding=(w.shape[-2] // 2, w.shape[-1] // 2))
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1).to(x.dtype)
    return y


def q4_block_forward(sd, name: str, flavor: str, resample: str, spec: Q4Spec, x: Tensor, emb: Tensor) -> Tensor:
    """Block.forward (:120-150)."""
    p = name + "."
    if resample == "down":
        x = F.avg_pool2d(x, 2)
    elif resample == "up":
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    if flavor == "enc":
        if p + "conv_skip.weight" in sd:
            x = mp_conv2d(x, sd[p + "conv_skip.weight"])
        x = normalize(x, dim=1)
    y = mp_conv2d(mp_silu(x), sd[p + "conv_res0.weight"])
    c = mp_conv2d(emb, sd[p + "emb_linear.weight"], gain=sd[p + "emb_gain"]) + 1.0
    y = mp_silu(y * c)
    y = mp_conv2d(y, sd[p + "conv_res1.weight"])
    if flavor == "dec" and p + "conv_skip.weight" in sd:
        x = mp_conv2d(x, sd[p + "conv_skip.weight"])
    return mp_sum(x, y, spec.res_balance).clip(-256.0, 256.0)


def q4_forward(sd: Dict[str, Tensor], spec: Q4Spec, x_in: Tensor, sigma: Tensor, x_ref: Tensor,
               body_dtype=torch.float32) -> Tensor:
    """UNet.forward (:253-303)."""
    sig = sigma.float().view(-1, 1, 1, 1)
    sd2 = spec.sigma_data ** 2
    c_skip = sd2 / (sig ** 2 + sd2)
    c_out = sig * spec.sigma_data / (sig ** 2 + sd2).sqrt()
    c_in = 1 / (sd2 + sig ** 2).sqrt()
    c_noise = sig.flatten().log() / 4
    b, c, _, w = x_ref.shape
    k = spec.psd_per_freq
    xr = x_ref.view(b, c, spec.in_num_freqs, k, w).permute(0, 3, 1, 2, 4).reshape(b, k * c, spec.in_num_freqs, w).to(body_dtype)
    x = (c_in * x_in.float()).to(body_dtype)
    x = mp_cat(x, xr, spec.label_balance)
    emb = mp_conv2d(mp_fourier(c_noise, sd["emb_fourier.freqs"], sd["emb_fourier.phases"]), sd["emb_noise.weight"])
    emb = emb[:, :, None, None].to(body_dtype)
    enc, dec, _ = q4_block_plan(spec)
    skips = []
    for name, kind, cin, cout, resample, _ in enc:
        if kind == "conv":
            x = mp_conv2d(x, sd[name + ".weight"], bias=sd[name + ".bias"])
        else:
            x = q4_block_forward(sd, name, "enc", resample, spec, x, emb)
        skips.append(x)
    for name, kind, cin, cout, resample, takes_skip in dec:
        if takes_skip:
            x = mp_cat(x, skips.pop(), spec.concat_balance)
        x = q4_block_forward(sd, name, "dec", resample, spec, x, emb)
    x = mp_conv2d(x, sd["conv_out.weight"], gain=sd["out_gain"])
    return c_skip * x_in.float() + c_out * x.float()


def q4_state_dict_shapes(spec: Q4Spec) -> Dict[str, Tuple[int, ...]]:
    shapes: Dict[str, Tuple[int, ...]] = {"out_gain": ()}
    shapes["emb_fourier.freqs"] = shapes["emb_fourier.phases"] = (spec.cnoise,)
    shapes["emb_noise.weight"] = (spec.cemb, spec.cnoise)
    shapes["logvar_fourier.freqs"] = shapes["logvar_fourier.phases"] = (spec.logvar_channels,)
    shapes["logvar_linear.weight"] = (1, spec.logvar_channels)
    enc, dec, cout = q4_block_plan(spec)
    m = spec.mlp_multiplier
    for name, kind, cin, co, resample, _ in enc + dec:
        if kind == "conv":
            shapes[name + ".weight"] = (co, cin, 3, 3)
            shapes[name + ".bias"] = (co,)
            continue
        p = name + "."
        flavor = "enc" if name.startswith("enc") else "dec"
        shapes[p + "conv_res0.weight"] = (co * m, co if flavor == "enc" else cin, 3, 3)
        shapes[p + "conv_res1.weight"] = (co, co * m, 3, 3)
        if cin != co:
            shapes[p + "conv_skip.weight"] = (co, cin, 1, 1)
        shapes[p + "emb_gain"] = ()
        shapes[p + "emb_linear.weight"] = (co * m, spec.cemb, 1, 1)
    shapes["conv_out.weight"] = (spec.out_channels, cout, 3, 3)
    return shapes


def synth_q4_state_dict(spec: Q4Spec, seed: int = 0, gain: float = 0.5) -> Dict[str, Tensor]:
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in sorted(q4_state_dict_shapes(spec).items()):
        if name.endswith(".freqs") or name.endswith(".phases"):
            continue
        if shape == ():
            sd[name] = torch.tensor(gain)
        elif name.endswith(".bias"):
            sd[name] = 0.2 * torch.randn(shape, generator=gen)
        else:
            w = torch.randn(shape, generator=gen)
            sd[name] = w if name == "logvar_linear.weight" else normalize(w)
    sd["emb_fourier.freqs"], sd["emb_fourier.phases"] = mp_fourier_buffers(spec.cnoise)
    sd["logvar_fourier.freqs"], sd["logvar_fourier.phases"] = mp_fourier_buffers(spec.logvar_channels)
    return sd
