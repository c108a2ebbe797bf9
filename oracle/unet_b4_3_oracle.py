"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch) of the reference's time-axis UNet lineage
`modules.unets.unet_edm2_b4_3.UNet` (the one lineage without a CUDA path yet;
/root/reference/src/modules/unets/unet_edm2_b4_3.py, modules/rope.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file; the product path never does.

Parity status: PINNED against the unmodified reference (tests/golden/make_golden_b4_3.py -> unet_b4_3_small.pt;
tests/test_oracle.py::test_b4_3_oracle_matches_reference_golden).  The reference hard-codes bfloat16 casts of the network
input and of the embedding (:272-279), so its body runs in bf16 even on the CPU: `body_dtype=torch.bfloat16` reproduces that
statement by statement (used for pinning), `body_dtype=torch.float32` is the same arithmetic without the rounding (the
yardstick a CUDA path would be held to).  Eval mode (no weight normalisation inside the forward), dropout 0.

What makes this lineage a different network rather than a rewiring of b4: the 2-D latent is folded
into channels (in_channels * in_freqs) and only time remains as a spatial axis; one level of `model_channels` channels;
(1,3) grouped convolutions; attention over the time axis with `channels_per_head`-wide heads whose first `rope_channels`
channels are rotated pairwise (and re-ordered even | odd | tail: q and k get the same permutation, so their dot products
are unchanged by it); U-shaped skips between the first and the second half of the layer stack through a 1x1 conv over the
concatenation; the input is mixed back into the first 2 * cdata channels before every block -- in place, so a tensor
already parked in `skips` is modified too (:303, reproduced below).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .unet_oracle import mp_fourier, mp_fourier_buffers, mp_silu, mp_sum, normalize

Tensor = torch.Tensor


@dataclass
class B43Spec:
    """unet_edm2_b4_3.py:44-80 (dataclass defaults)."""
    in_channels: int = 8
    out_channels: int = 8
    in_channels_emb: int = 1024
    in_freqs: int = 32
    sigma_data: float = 1.0
    mp_fourier_ln_sigma_offset: float = 0.5
    mp_fourier_bandwidth: float = 1.0
    model_channels: int = 2048
    logvar_channels: int = 192
    channel_mult_noise: int = 1
    channel_mult_emb: int = 1
    use_skips: bool = True
    use_conv_skip: bool = True
    channels_per_head: int = 128
    rope_channels: int = 112
    rope_base: float = 10000.0
    num_layers_per_block: int = 9
    label_balance: float = 0.5
    res_balance: float = 0.5
    attn_balance: float = 0.5
    mlp_multiplier: int = 4
    mlp_groups: int = 4
    emb_linear_groups: int = 4
    input_skip_t: float = 0.5

    @property
    def cdata(self) -> int:
        return self.in_channels * self.in_freqs

    @property
    def cnoise(self) -> int:
        return self.model_channels * self.channel_mult_noise

    @property
    def cemb(self) -> int:
        return self.model_channels * self.channel_mult_emb


def small_b4_3_spec() -> B43Spec:
    return B43Spec(in_channels=2, out_channels=2, in_channels_emb=16, in_freqs=8, model_channels=64, logvar_channels=32,
                   channels_per_head=32, rope_channels=24, num_layers_per_block=3, mlp_multiplier=2, mlp_groups=2,
                   emb_linear_groups=2)


def skip_channels(spec: B43Spec, idx: int) -> int:
    """:231-234."""
    return spec.model_channels if spec.use_skips and spec.use_conv_skip and idx >= spec.num_layers_per_block / 2 else 0


def mp_conv(x: Tensor, w: Tensor, gain=1.0, groups: int = 1) -> Tensor:
    """MPConv.forward, eval mode (mp_tools.py:357-373), weights cast to the activation dtype."""
    w = (w.float() * (gain / math.sqrt(w[0].numel()))).to(x.dtype)
    if w.ndim == 2:
        return x @ w.t()
    return F.conv2d(x, w, padding=(w.shape[-2] // 2, w.shape[-1] // 2), groups=groups)


def rope_tables(n: int, rope_ch: int, base: float, dtype) -> Tuple[Tensor, Tensor]:
    """rope.py:46-78: cos / sin of position * base^(-2i / rope_ch), shaped (1, 1, N, rope_ch / 2), cast to the body dtype."""
    inv_freq = 1.0 / (base ** (torch.arange(0, rope_ch, 2, dtype=torch.float32) / rope_ch))
    ang = torch.einsum("w,d->wd", torch.arange(n, dtype=torch.float32), inv_freq)
    return ang.cos().to(dtype).view(1, 1, n, rope_ch // 2), ang.sin().to(dtype).view(1, 1, n, rope_ch // 2)


def rope_rotate(x: Tensor, tables: Tuple[Tensor, Tensor]) -> Tensor:
    """rope.py:26-44: pairwise rotation of the first rope_ch channels of x [..., D]; output order even | odd | tail."""
    cos, sin = tables
    rope_ch = cos.shape[-1] * 2
    x_rot, x_tail = x[..., :rope_ch], x[..., rope_ch:]
    x_even, x_odd = x_rot[..., 0::2], x_rot[..., 1::2]
    return torch.cat([x_even * cos - x_odd * sin, x_odd * cos + x_even * sin, x_tail], dim=-1)


def block_forward(sd: Dict[str, Tensor], name: str, spec: B43Spec, x: Tensor, emb: Tensor, tables) -> Tensor:
    """Block.forward (:133-175), use_attention True (attn_levels = (0,) and there is one level)."""
    p = name + "."
    if p + "conv_skip.weight" in sd:
        x = mp_conv(x, sd[p + "conv_skip.weight"])
    heads = spec.model_channels // spec.channels_per_head
    c = mp_conv(emb, sd[p + "emb_linear_qkv.weight"], gain=sd[p + "emb_gain_qkv"]) + 1.0
    y = x * c
    n_tok = y.shape[2] * y.shape[3]
    q, k, v = (normalize(mp_conv(y, sd[p + f"attn_{t}.weight"]).reshape(y.shape[0], heads, -1, n_tok), dim=2) for t in "qkv")
    q_rot = rope_rotate(q.transpose(-1, -2), tables)
    k_rot = rope_rotate(k.transpose(-1, -2), tables)
    y = F.scaled_dot_product_attention(q_rot, k_rot, v.transpose(-1, -2)).transpose(-1, -2)
    y = mp_conv(y.reshape(*x.shape), sd[p + "attn_proj.weight"])
    x = mp_sum(x, y, spec.attn_balance)
    y = mp_conv(x, sd[p + "conv_res0.weight"], groups=spec.mlp_groups)
    c = mp_conv(emb, sd[p + "emb_linear.weight"], gain=sd[p + "emb_gain"], groups=spec.emb_linear_groups) + 1.0
    y = mp_silu(normalize(y * c, dim=1))
    y = mp_conv(y, sd[p + "conv_res1.weight"])
    return mp_sum(x, y, spec.res_balance).clip(-256.0, 256.0)


def get_embeddings(sd: Dict[str, Tensor], emb_in: Tensor, conditioning_mask: Tensor) -> Tensor:
    """:241-244."""
    u = mp_conv(torch.ones(1), sd["emb_label_unconditional.weight"])
    c = mp_conv(normalize(emb_in.float()), sd["emb_label.weight"])
    return mp_sum(u, c, conditioning_mask.unsqueeze(1).float())


def sigma_loss_logvar(sd: Dict[str, Tensor], spec: B43Spec, sigma: Tensor) -> Tensor:
    """:246-248."""
    ln_sigma = sigma.flatten().log() - spec.mp_fourier_ln_sigma_offset
    four = mp_fourier(ln_sigma / 4, sd["logvar_fourier.freqs"], sd["logvar_fourier.phases"])
    return mp_conv(four, sd["logvar_linear.weight"]).view(-1, 1, 1, 1).float()


def forward(sd: Dict[str, Tensor], spec: B43Spec, x_in: Tensor, sigma: Tensor, embeddings: Tensor,
            x_ref: Optional[Tensor] = None, body_dtype=torch.float32) -> Tensor:
    """UNet.forward (:254-319)."""
    sig = sigma.float().view(-1, 1, 1, 1)
    sd2 = spec.sigma_data ** 2
    c_skip = sd2 / (sig ** 2 + sd2)
    c_out = sig * spec.sigma_data / (sig ** 2 + sd2).sqrt()
    c_in = 1 / (sd2 + sig ** 2).sqrt()
    c_noise = (sig.flatten().log() - spec.mp_fourier_ln_sigma_offset) / 4
    x = (c_in * x_in.float()).to(body_dtype)
    emb = mp_conv(mp_fourier(c_noise, sd["emb_fourier.freqs"], sd["emb_fourier.phases"]), sd["emb_noise.weight"])
    emb = mp_sum(emb, embeddings.to(emb.dtype), spec.label_balance)
    emb = mp_silu(emb).unsqueeze(2).unsqueeze(3).to(body_dtype)
    tables = rope_tables(x.shape[3], spec.rope_channels, spec.rope_base, x.dtype)
    x = x.reshape(x.shape[0], spec.cdata, 1, x.shape[3])
    x_input = torch.cat((x, -x), dim=1)
    x = torch.cat((x, torch.ones_like(x[:, :1])), dim=1)
    x = mp_conv(x, sd["dec.conv_in.weight"])
    skips: List[Tensor] = []
    n = spec.num_layers_per_block
    for idx in range(n):
        if spec.use_skips and idx >= n / 2:
            x = torch.cat((x, skips.pop()), dim=1) if spec.use_conv_skip else mp_sum(x, skips.pop(), 0.5)
        if spec.input_skip_t > 0:
            # in place, as the reference (:303): when x is the tensor appended to `skips` one iteration earlier, the
            # parked skip receives the input mix as well
            x[:, :x_input.shape[1]] = mp_sum(x[:, :x_input.shape[1]], x_input, spec.input_skip_t)
        x = block_forward(sd, f"dec.block0_layer{idx}", spec, x, emb, tables)
        if spec.use_skips and idx < n / 2 - 0.5:
            skips.append(x)
    x = mp_conv(x, sd["conv_out.weight"], gain=sd["out_gain"])
    x = x.reshape(x.shape[0], spec.out_channels, spec.in_freqs, x.shape[3])
    d = c_skip * x_in.float() + c_out * x.float()
    if x_ref is not None:
        d = mp_sum(x_ref[:, :-1].float(), d, x_ref[:, -1:].float())
    return d


def state_dict_shapes(spec: B43Spec) -> Dict[str, Tuple[int, ...]]:
    C, m = spec.model_channels, spec.mlp_multiplier
    shapes: Dict[str, Tuple[int, ...]] = {"out_gain": ()}
    shapes["emb_fourier.freqs"] = shapes["emb_fourier.phases"] = (spec.cnoise,)
    shapes["emb_noise.weight"] = (spec.cemb, spec.cnoise)
    shapes["emb_label.weight"] = (spec.cemb, spec.in_channels_emb)
    shapes["emb_label_unconditional.weight"] = (spec.cemb, 1)
    shapes["logvar_fourier.freqs"] = shapes["logvar_fourier.phases"] = (spec.logvar_channels,)
    shapes["logvar_linear.weight"] = (1, spec.logvar_channels)
    shapes["dec.conv_in.weight"] = (C, spec.cdata + 1, 1, 3)
    for idx in range(spec.num_layers_per_block):
        p = f"dec.block0_layer{idx}."
        cs = skip_channels(spec, idx)
        if cs:
            shapes[p + "conv_skip.weight"] = (C, C + cs, 1, 1)
        shapes[p + "conv_res0.weight"] = (C * m, C // spec.mlp_groups, 1, 3)
        shapes[p + "conv_res1.weight"] = (C, C * m, 1, 1)
        shapes[p + "emb_gain"] = shapes[p + "emb_gain_qkv"] = ()
        shapes[p + "emb_linear.weight"] = (C * m, spec.cemb // spec.emb_linear_groups, 1, 1)
        shapes[p + "emb_linear_qkv.weight"] = (C, spec.cemb, 1, 1)
        for t in ("q", "k", "v", "proj"):
            shapes[p + f"attn_{t}.weight"] = (C, C, 1, 1)
    shapes["conv_out.weight"] = (spec.out_channels * spec.in_freqs, C, 1, 3)
    return shapes


def synth_state_dict(spec: B43Spec, seed: int = 0, gain: float = 0.5) -> Dict[str, Tensor]:
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in sorted(state_dict_shapes(spec).items()):
        if name.endswith(".freqs") or name.endswith(".phases"):
            continue
        if shape == ():
            sd[name] = torch.tensor(gain)
        else:
            w = torch.randn(shape, generator=gen)
            sd[name] = w if name == "logvar_linear.weight" else normalize(w)
    sd["emb_fourier.freqs"], sd["emb_fourier.phases"] = mp_fourier_buffers(spec.cnoise, spec.mp_fourier_bandwidth)
    sd["logvar_fourier.freqs"], sd["logvar_fourier.phases"] = mp_fourier_buffers(spec.logvar_channels)
    return sd
