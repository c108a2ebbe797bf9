"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain PyTorch, fp32) of the reference's
EDM2 UNet denoise forward.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this file; the product path
(dualdiffusion_b200/) never does.

Parity status: PINNED against the reference itself — `tests/golden/make_golden.py`
imports the unmodified reference (oracle/ref_shim.py) in the build container and
writes golden vectors; `tests/test_oracle.py` checks this restatement against them
(and against the live reference when /root/reference is present).  The reference
ships no golden vectors of its own.

Every function cites the reference lines it restates (paths relative to
/root/reference/src).  The restatement is functional: it walks a reference-layout
state_dict (same key names / OIHW fp32 shapes as `modules.unets.unet_edm2_b4.UNet`)
instead of building nn.Modules.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# configuration (modules/unets/unet.py:33-42, modules/unets/unet_edm2_b4.py:41-58)
# --------------------------------------------------------------------------------------
@dataclass
class UNetSpec:
    in_channels: int = 4
    out_channels: int = 4
    in_channels_emb: int = 512
    dropout: float = 0.0
    sigma_max: float = 200.0
    sigma_min: float = 0.03
    sigma_data: float = 1.0
    model_channels: int = 256
    logvar_channels: int = 128
    channel_mult: Sequence[int] = (1, 2, 3, 4, 5)
    channel_mult_noise: Optional[int] = None
    channel_mult_emb: Optional[int] = None
    channels_per_head: int = 64
    num_layers_per_block: int = 2
    label_balance: float = 0.5
    concat_balance: float = 0.5
    res_balance: float = 0.3
    attn_balance: float = 0.3
    attn_levels: Sequence[int] = (3, 4)
    mlp_multiplier: int = 2
    mlp_groups: int = 8
    # format side: number of mel filters spanning freq_min..sample_rate/2 that feed the
    # positional "ln_freqs" channel (modules/formats/ms_mdct_dual.py:144-153 defaults)
    ms_freq_min: float = 0.0
    sample_rate: int = 32000

    @property
    def cblock(self) -> List[int]:
        return [self.model_channels * m for m in self.channel_mult]

    @property
    def cnoise(self) -> int:
        return (self.model_channels * self.channel_mult_noise
                if self.channel_mult_noise is not None else max(self.cblock))

    @property
    def cemb(self) -> int:
        return (self.model_channels * self.channel_mult_emb
                if self.channel_mult_emb is not None else max(self.cblock))


def default_spec() -> UNetSpec:
    """config/models/default/unet.json of the reference."""
    return UNetSpec(channel_mult_noise=1, channel_mult_emb=3)


@dataclass
class BlockSpec:
    """One entry of the enc/dec ModuleDicts (unet_edm2_b4.py:189-227)."""
    name: str            # state_dict prefix, e.g. "enc.block1_down"
    kind: str            # "conv" (enc.conv_in) or "block"
    level: int
    cin: int
    cout: int
    flavor: str = "enc"
    resample: str = "keep"
    attention: bool = False
    takes_skip: bool = False   # dec "layer" blocks consume a skip via mp_cat


def block_plan(spec: UNetSpec) -> Tuple[List[BlockSpec], List[BlockSpec]]:
    """Enumerate encoder / decoder blocks exactly as UNet.__init__ wires them
    (unet_edm2_b4.py:189-227), including the channel bookkeeping of the skips."""
    cblock = spec.cblock
    enc: List[BlockSpec] = []
    cout = spec.in_channels + 2
    for level, ch in enumerate(cblock):
        attn = level in spec.attn_levels
        if level == 0:
            enc.append(BlockSpec("enc.conv_in", "conv", 0, cout, ch))
            cout = ch
        else:
            enc.append(BlockSpec(f"enc.block{level}_down", "block", level, cout, cout,
                                 "enc", "down", attn))
        for i in range(spec.num_layers_per_block):
            enc.append(BlockSpec(f"enc.block{level}_layer{i}", "block", level, cout, ch,
                                 "enc", "keep", attn))
            cout = ch
    skips = [b.cout for b in enc]
    dec: List[BlockSpec] = []
    for level in reversed(range(len(cblock))):
        ch = cblock[level]
        attn = level in spec.attn_levels
        if level == len(cblock) - 1:
            dec.append(BlockSpec(f"dec.block{level}_in0", "block", level, cout, cout, "dec", "keep", True))
            dec.append(BlockSpec(f"dec.block{level}_in1", "block", level, cout, cout, "dec", "keep", True))
        else:
            dec.append(BlockSpec(f"dec.block{level}_up", "block", level, cout, cout, "dec", "up", attn))
        for i in range(spec.num_layers_per_block + 1):
            cin = cout + skips.pop()
            dec.append(BlockSpec(f"dec.block{level}_layer{i}", "block", level, cin, ch,
                                 "dec", "keep", attn, takes_skip=True))
            cout = ch
    return enc, dec


# --------------------------------------------------------------------------------------
# magnitude-preserving operator set (modules/mp_tools.py)
# --------------------------------------------------------------------------------------
def normalize(x: Tensor, dim=None, eps: float = 1e-4) -> Tensor:
    """mp_tools.py:42-49 — x / (eps + ||x||_2 * sqrt(n_norms / n_elems)), fp32 math."""
    dims = list(range(1, x.ndim)) if dim is None else dim
    n = torch.linalg.vector_norm(x.float(), dim=dims, keepdim=True)
    n = eps + n * math.sqrt(n.numel() / x.numel())
    return (x.float() / n).to(x.dtype)


def mp_silu(x: Tensor) -> Tensor:
    """mp_tools.py:268-269."""
    return F.silu(x) / 0.596


def mp_sum(a: Tensor, b: Tensor, t=0.5) -> Tensor:
    """mp_tools.py:274-279 — lerp(a,b,t)/sqrt((1-t)^2+t^2); t float or tensor."""
    if isinstance(t, Tensor):
        return torch.lerp(a, b, t) / torch.sqrt((1 - t) ** 2 + t ** 2).to(a.dtype)
    return torch.lerp(a, b, t) / math.sqrt((1 - t) ** 2 + t ** 2)


def mp_cat_weights(na: int, nb: int, t: float) -> Tuple[float, float]:
    """mp_tools.py:294-301 — per-source scales of the magnitude-preserving concat."""
    c = math.sqrt((na + nb) / ((1 - t) ** 2 + t ** 2))
    return c / math.sqrt(na) * (1 - t), c / math.sqrt(nb) * t


def mp_cat(a: Tensor, b: Tensor, t: float = 0.5) -> Tensor:
    wa, wb = mp_cat_weights(a.shape[1], b.shape[1], t)
    return torch.cat([wa * a, wb * b], dim=1)


def mp_fourier(x: Tensor, freqs: Tensor, phases: Tensor) -> Tensor:
    """mp_tools.py:324-330 (1-D input branch)."""
    y = torch.outer(x.float(), freqs.float()) + phases.float()
    return (y.cos() * math.sqrt(2.0)).to(x.dtype)


def mp_fourier_buffers(n: int, bandwidth: float = 1.0, eps: float = 1e-3) -> Tuple[Tensor, Tensor]:
    """mp_tools.py:318-322."""
    freqs = math.pi * torch.linspace(0, 1 - eps, n).erfinv() * bandwidth
    phases = math.pi / 2 * (torch.arange(n) % 2 == 0).float()
    return freqs, phases


def mp_weight(w: Tensor, gain=1.0, training: bool = False, out_dtype=torch.float32) -> Tensor:
    """mp_tools.py:359-364 — weight prep: (train-only) per-output-channel normalise, then
    scale by gain / sqrt(fan_in), then cast to the activation dtype."""
    w = w.float()
    if training:
        w = normalize(w)
    w = w * (gain / math.sqrt(w[0].numel()))
    return w.to(out_dtype)


def mp_conv(x: Tensor, w: Tensor, gain=1.0, groups: int = 1, training: bool = False) -> Tensor:
    """mp_tools.py:357-373 (no-bias path)."""
    w = mp_weight(w, gain, training, x.dtype)
    if w.ndim == 2:
        return x @ w.t()
    return F.conv2d(x, w, padding=(w.shape[-2] // 2, w.shape[-1] // 2), groups=groups)


def resample_2d(x: Tensor, mode: str) -> Tensor:
    """mp_tools.py:71-79 — down = plain 2x2 mean (not rescaled), up = nearest x2."""
    if mode == "keep":
        return x
    if mode == "down":
        return F.avg_pool2d(x, 2)
    return F.interpolate(x, scale_factor=2, mode="nearest")


# --------------------------------------------------------------------------------------
# mel positional channel (modules/formats/frequency_scale.py:30-34,144-149;
# unet_edm2_b4.py:244-248)
# --------------------------------------------------------------------------------------
def mel_points(freq_min: float, freq_max: float, n: int) -> Tensor:
    lo = 2595.0 * math.log10(1.0 + freq_min / 700.0)
    hi = 2595.0 * math.log10(1.0 + freq_max / 700.0)
    mels = torch.linspace(lo, hi, n)
    return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)


def ln_freqs_channel(spec: UNetSpec, b: int, h: int, w: int) -> Tensor:
    f = mel_points(spec.ms_freq_min, spec.sample_rate / 2, h + 2)[1:-1].log2()
    f = f.view(1, 1, -1, 1).repeat(b, 1, 1, w)
    return (f - f.mean()) / f.std()


# --------------------------------------------------------------------------------------
# Block (unet_edm2_b4.py:110-158) and UNet.forward (:250-296)
# --------------------------------------------------------------------------------------
def attention_core(qk: Tensor, v: Tensor, heads: int) -> Tensor:
    """unet_edm2_b4.py:136-148 — channel-major q/k/v are split per head, cosine-normalised
    over the 64 head channels (eps 1e-4 of `normalize`), and run through softmax(QK^T/8)V."""
    b, _, h, w = v.shape
    qk = qk.reshape(b, heads, -1, 2, h * w)
    q, k = normalize(qk, dim=2).unbind(3)
    vv = normalize(v.reshape(b, heads, -1, h * w), dim=2)
    d = q.shape[2]
    s = torch.einsum("bhdq,bhdk->bhqk", q.float(), k.float()) / math.sqrt(d)
    p = torch.softmax(s, dim=-1)
    y = torch.einsum("bhqk,bhdk->bhdq", p, vv.float()).to(v.dtype)
    return y.reshape(b, -1, h, w)


def block_forward(sd: Dict[str, Tensor], bs: BlockSpec, spec: UNetSpec, x: Tensor, emb: Tensor,
                  training: bool = False, taps: Optional[dict] = None) -> Tensor:
    p = bs.name + "."
    g = spec.mlp_groups

    def conv(name, inp, gain=1.0, groups=1):
        return mp_conv(inp, sd[p + name + ".weight"], gain, groups, training)

    x = resample_2d(x, bs.resample)
    if bs.flavor == "enc":
        x = conv("conv_skip", x)
        x = normalize(x, dim=1)
    y = conv("conv_res0", mp_silu(x), groups=g)
    c = conv("emb_linear", emb, gain=sd[p + "emb_gain"], groups=g) + 1.0
    y = mp_silu(y * c)
    y = conv("conv_res1", y, groups=g)
    if bs.flavor == "dec":
        x = conv("conv_skip", x)
    x = mp_sum(x, y, spec.res_balance)
    if bs.attention:
        heads = bs.cout // spec.channels_per_head
        c = conv("emb_linear_qk", emb, gain=sd[p + "emb_gain_qk"]) + 1.0
        qk = conv("attn_qk", x * c)
        v = conv("attn_v", x)
        y = attention_core(qk, v, heads)
        c = conv("emb_linear_v", emb, gain=sd[p + "emb_gain_v"]) + 1.0
        y = mp_silu(y * c)
        y = conv("attn_proj", y)
        x = mp_sum(x, y, spec.attn_balance)
    x = x.clip(-256.0, 256.0)
    if taps is not None:
        taps[bs.name] = x
    return x


def get_embeddings(sd: Dict[str, Tensor], emb_in: Tensor, conditioning_mask: Tensor) -> Tensor:
    """unet_edm2_b4.py:232-235."""
    u = mp_conv(torch.ones(1), sd["emb_label_unconditional.weight"])
    c = mp_conv(normalize(emb_in.float()), sd["emb_label.weight"])
    return mp_sum(u, c, conditioning_mask.unsqueeze(1).float())


def sigma_loss_logvar(sd: Dict[str, Tensor], sigma: Tensor) -> Tensor:
    """unet_edm2_b4.py:237-238."""
    f = mp_fourier(sigma.flatten().float().log() / 4, sd["logvar_fourier.freqs"], sd["logvar_fourier.phases"])
    return mp_conv(f, sd["logvar_linear.weight"]).view(-1, 1, 1, 1).float()


def unet_forward(sd: Dict[str, Tensor], spec: UNetSpec, x_in: Tensor, sigma: Tensor, embeddings: Tensor,
                 x_ref: Optional[Tensor] = None, training: bool = False,
                 taps: Optional[dict] = None) -> Tensor:
    """unet_edm2_b4.py:250-296, all arithmetic in fp32."""
    x_in = x_in.float()
    sigma = sigma.float().view(-1, 1, 1, 1)
    sd2 = spec.sigma_data ** 2
    c_skip = sd2 / (sigma ** 2 + sd2)
    c_out = sigma * spec.sigma_data / (sigma ** 2 + sd2).sqrt()
    c_in = 1 / (sd2 + sigma ** 2).sqrt()
    c_noise = sigma.flatten().log() / 4

    emb = mp_conv(mp_fourier(c_noise, sd["emb_fourier.freqs"], sd["emb_fourier.phases"]),
                  sd["emb_noise.weight"], training=training)
    emb = mp_sum(emb, embeddings.float(), spec.label_balance)
    emb = mp_silu(emb)[:, :, None, None]
    if taps is not None:
        taps["emb"] = emb[:, :, 0, 0]

    x = c_in * x_in
    b, _, h, w = x.shape
    x = torch.cat((x, torch.ones_like(x[:, :1]), ln_freqs_channel(spec, b, h, w)), dim=1)

    enc, dec = block_plan(spec)
    skips: List[Tensor] = []
    for bs in enc:
        if bs.kind == "conv":
            x = mp_conv(x, sd[bs.name + ".weight"], training=training)
            if taps is not None:
                taps[bs.name] = x
        else:
            x = block_forward(sd, bs, spec, x, emb, training, taps)
        skips.append(x)
    for bs in dec:
        if bs.takes_skip:
            x = mp_cat(x, skips.pop(), spec.concat_balance)
        x = block_forward(sd, bs, spec, x, emb, training, taps)
    x = mp_conv(x, sd["conv_out.weight"], gain=sd["out_gain"], training=training)
    d = c_skip * x_in + c_out * x
    if x_ref is not None:
        d = mp_sum(x_ref[:, :-1].float(), d, x_ref[:, -1:].float())
    return d


# --------------------------------------------------------------------------------------
# deterministic synthetic weights (random init as MPConv.__init__ does: randn, then the
# post-load normalize_weights() of module.py:185-191; scalar gains made non-zero so the
# body contributes — SURVEY.md §8(d))
# --------------------------------------------------------------------------------------
def state_dict_shapes(spec: UNetSpec) -> Dict[str, Tuple[int, ...]]:
    shapes: Dict[str, Tuple[int, ...]] = {}
    cemb, g, m = spec.cemb, spec.mlp_groups, spec.mlp_multiplier
    shapes["out_gain"] = ()
    shapes["emb_fourier.freqs"] = (spec.cnoise,)
    shapes["emb_fourier.phases"] = (spec.cnoise,)
    shapes["emb_noise.weight"] = (cemb, spec.cnoise)
    shapes["emb_label.weight"] = (cemb, spec.in_channels_emb)
    shapes["emb_label_unconditional.weight"] = (cemb, 1)
    shapes["logvar_fourier.freqs"] = (spec.logvar_channels,)
    shapes["logvar_fourier.phases"] = (spec.logvar_channels,)
    shapes["logvar_linear.weight"] = (1, spec.logvar_channels)
    enc, dec = block_plan(spec)
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
        shapes[p + "emb_linear.weight"] = (bs.cout * m, cemb // g, 1, 1)
        if bs.attention:
            shapes[p + "emb_gain_qk"] = ()
            shapes[p + "emb_gain_v"] = ()
            shapes[p + "emb_linear_qk.weight"] = (bs.cout, cemb, 1, 1)
            shapes[p + "emb_linear_v.weight"] = (bs.cout, cemb, 1, 1)
            shapes[p + "attn_qk.weight"] = (bs.cout * 2, bs.cout, 1, 1)
            shapes[p + "attn_v.weight"] = (bs.cout, bs.cout, 1, 1)
            shapes[p + "attn_proj.weight"] = (bs.cout, bs.cout, 1, 1)
    shapes["conv_out.weight"] = (spec.out_channels, cblock_last(spec), 3, 3)
    return shapes


def cblock_last(spec: UNetSpec) -> int:
    return spec.cblock[0]


def synth_state_dict(spec: UNetSpec, seed: int = 0, gain: float = 0.5) -> Dict[str, Tensor]:
    """Seeded CPU weights in reference layout.  Keys are drawn in sorted-name order from one
    CPU generator so the result does not depend on module construction order."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in sorted(state_dict_shapes(spec).items()):
        if name.endswith(".freqs") or name.endswith(".phases"):
            continue
        if shape == ():
            sd[name] = torch.tensor(gain)
        else:
            w = torch.randn(shape, generator=gen)
            if name != "logvar_linear.weight":      # disable_weight_norm=True there
                w = normalize(w)
            sd[name] = w
    sd["emb_fourier.freqs"], sd["emb_fourier.phases"] = mp_fourier_buffers(spec.cnoise)
    sd["logvar_fourier.freqs"], sd["logvar_fourier.phases"] = mp_fourier_buffers(spec.logvar_channels)
    return sd


def small_spec() -> UNetSpec:
    """A reduced UNet that still exercises every code path (down/up, skip-concat with unequal
    channel counts, grouped 3x3 with 32- and 64-wide groups, attention at the coarsest level)
    but builds and runs on CPU in well under a second."""
    return UNetSpec(model_channels=256, channel_mult=(1, 2), channel_mult_noise=1, channel_mult_emb=2,
                    num_layers_per_block=1, attn_levels=(1,), in_channels_emb=64, logvar_channels=32)


# --------------------------------------------------------------------------------------
# train step (training/module_trainers/unet_trainer.py:236-280): EDM loss with the learned
# per-noise-level uncertainty, as a function of a reference-layout state_dict
# --------------------------------------------------------------------------------------
def train_loss(sd: Dict[str, Tensor], spec: UNetSpec, samples: Tensor, noise: Tensor, sigma: Tensor, clap: Tensor,
               conditioning_mask: Tensor) -> Tensor:
    """unet_trainer.py:239,259-280 (input_perturbation = 0, no ref_samples / loss_weight) followed by the
    `.mean()` of trainer.py:1016.  `noise` is unit-variance; the trainer scales it by sigma (:252).  Weight
    normalisation runs inside the forward (train mode, mp_tools.py:360-361).  Keeps the (B,1,1,B) broadcast of
    `batch_weighted_loss / error_logvar.exp() + error_logvar`."""
    u = mp_conv(torch.ones(1), sd["emb_label_unconditional.weight"], training=True)
    c = mp_conv(normalize(clap.float()), sd["emb_label.weight"], training=True)
    emb = mp_sum(u, c, conditioning_mask.unsqueeze(1).float())
    sig = sigma.float().view(-1, 1, 1, 1)
    denoised = unet_forward(sd, spec, samples + noise * sig, sigma, emb, training=True)
    w = (sig ** 2 + spec.sigma_data ** 2) / (sig * spec.sigma_data) ** 2
    wl = (F.mse_loss(denoised, samples, reduction="none") * w).mean(dim=(1, 2, 3))
    logvar = sigma_loss_logvar(sd, sigma)
    return (wl / logvar.exp() + logvar).mean()


def grad_probe(name: str, shape) -> Tensor:
    """Seeded random direction per parameter name: golden files store <grad, probe> and ||grad|| instead of
    the full gradients (tests/golden/make_golden.py)."""
    g = torch.Generator().manual_seed(sum(ord(ch) * (i + 1) for i, ch in enumerate(name)) % (2 ** 31))
    return torch.randn(tuple(shape), generator=g)


# --------------------------------------------------------------------------------------
# axis ("separable") attention of the legacy ddec UNets
# (modules/unets/old/unet_edm2_ddec_mdct_b3.py:144-163) -- the only row/col <-> batch reshape in the reference
# --------------------------------------------------------------------------------------
def axis_attention_b3(qkv: Tensor, heads: int) -> Tensor:
    """qkv (b, 3c, z, h, w) with channel index (head, d, j in {q,k,v}); attention over h with (b, z, w) folded into
    the batch; returns y (b, c, z, h, w) BEFORE the block's mp_silu / attn_proj.  Statement by statement as :146-160."""
    b, c3, z, h, w = qkv.shape
    x = qkv.permute(0, 2, 4, 1, 3)                                        # b z w c h   (:148)
    x = x.reshape(b * z * w, heads, -1, 3, h)                             # (:149)
    q, k, v = normalize(x, dim=2).unbind(3)                               # (:150)
    d = q.shape[2]
    s = torch.einsum("nhdq,nhdk->nhqk", q.float(), k.float()) / math.sqrt(d)
    y = torch.einsum("nhqk,nhdk->nhdq", torch.softmax(s, dim=-1), v.float())      # SDPA on transposed views (:152-154)
    y = y.reshape(b, z, w, c3 // 3, h)                                    # (:157)
    return y.permute(0, 3, 1, 4, 2)                                       # b c z h w   (:159)
