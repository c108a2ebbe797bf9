"""B200-native drop-in for the reference's 2-D diffusion-decoder UNet `modules.unets.unet_edm2_q4_ddec.UNet`
(/root/reference/src/modules/unets/unet_edm2_q4_ddec.py:152-303; SURVEY.md section 8 row A17, second variant).

Same constructor / config dataclass / state_dict keys and shapes (conv_in bias included; strict `load_state_dict` of a
reference checkpoint works), same `forward(x_in, sigma, format, embeddings, x_ref, perturbed_input=None) -> D_x` (fp32,
(B, 2, F, W)), `get_embeddings` (None), `get_sigma_loss_logvar`, `get_latent_shape`.  Register it with
    "ddec": {"package": "dualdiffusion_b200.modules.unets.unet_edm2_q4_ddec", "class": "UNet"}

The forward is the EDM2 block schedule of unet_edm2_b4 without attention (conv_skip only where the width changes) on
the tcgen05 implicit-GEMM convolutions; the PSD reference joins the input through `dd_q4_stem` (view/permute/mp_cat of
:268-277 in one pass) and conv_in's bias rides on a constant-one input channel with a centre-tap weight column.
Eval mode, no_grad, CUDA graph per input shape; no CPU fallback.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

from typing import Union

import torch

from ... import _lib as L
from ... import ops
from ..mp_tools import MPFourier, mp_cat_weights
from .unet import DualDiffusionUNet, DualDiffusionUNetConfig

Tensor = torch.Tensor


def _ver(t: Tensor) -> int:
    return 0 if t.is_inference() else t._version


class _MPConv(torch.nn.Module):
    """Parameter container of mp_tools.MPConv (:332-378) including the optional bias (:347-353)."""

    def __init__(self, in_channels: int, out_channels: int, kernel: Tuple[int, ...], groups: int = 1, bias: bool = False,
                 disable_weight_norm: bool = False) -> None:
        super().__init__()
        self.in_channels, self.out_channels, self.groups = in_channels, out_channels, groups
        self.disable_weight_norm = disable_weight_norm
        self.weight = torch.nn.Parameter(torch.randn(out_channels, in_channels // groups, *kernel))
        self.weight.conv_groups = groups
        if bias:
            self.bias = torch.nn.Parameter(torch.zeros(out_channels))
            gd = out_channels // groups
            self.bias.data[0::2].fill_(1.0 / gd ** 0.5)
            self.bias.data[1::2].fill_(-1.0 / gd ** 0.5)
        else:
            self.bias = None

    @torch.no_grad()
    def normalize_weights(self) -> None:
        if not self.disable_weight_norm:
            from ..mp_tools import normalize
            self.weight.copy_(normalize(self.weight))


@dataclass
class UNet_Config(DualDiffusionUNetConfig):
    """unet_edm2_q4_ddec.py:44-70 (field-for-field, same defaults)."""
    in_channels: int = 2
    out_channels: int = 2
    in_channels_emb: int = 0
    in_num_freqs: int = 256
    in_psd_freqs: int = 2048
    model_channels: int = 32
    logvar_channels: int = 192
    channel_mult: Sequence[int] = (1, 2, 3, 4, 5)
    double_midblock: bool = True
    midblock_attn: bool = False
    channel_mult_noise: Optional[int] = 4
    channel_mult_emb: Optional[int] = 4
    channels_per_head: int = 64
    num_layers_per_block: int = 3
    label_balance: float = 0.5
    concat_balance: float = 0.5
    res_balance: float = 0.3
    attn_balance: float = 0.3
    attn_levels: Sequence[int] = ()
    mlp_multiplier: int = 2
    mlp_groups: int = 1
    emb_linear_groups: int = 1


class Block(torch.nn.Module):
    """Parameter container with the reference Block's names / shapes (:72-118)."""

    def __init__(self, level: int, in_channels: int, out_channels: int, emb_channels: int, num_freqs: int,
                 flavor: str = "enc", resample_mode: str = "keep", res_balance: float = 0.3, clip_act: float = 256,
                 mlp_multiplier: int = 1, use_attention: bool = False) -> None:
        super().__init__()
        if use_attention:
            raise NotImplementedError("unet_edm2_q4_ddec: attention is not implemented (the reference raises too, :118)")
        self.level, self.in_channels, self.out_channels = level, in_channels, out_channels
        self.flavor, self.resample_mode = flavor, resample_mode
        self.res_balance, self.clip_act = res_balance, clip_act
        self.conv_res0 = _MPConv(out_channels if flavor == "enc" else in_channels, out_channels * mlp_multiplier, (3, 3))
        self.conv_res1 = _MPConv(out_channels * mlp_multiplier, out_channels, (3, 3))
        self.conv_skip = _MPConv(in_channels, out_channels, (1, 1)) if in_channels != out_channels else None
        self.emb_gain = torch.nn.Parameter(torch.zeros([]))
        self.emb_linear = _MPConv(emb_channels, out_channels * mlp_multiplier, (1, 1))


class UNet(DualDiffusionUNet):

    # Parameters stay in PyTorch's default (OIHW) layout: the kernels read them through raw pointers and keep their own
    # NHWC activation / repacked-weight layouts, so the base class must not re-stride them (module.py:118-122).
    supports_channels_last: Union[bool, str] = False

    # resolved by from_pretrained (module.py:72); explicit because this file's annotations are strings
    config_class = UNet_Config

    supports_compile = False

    def __init__(self, config: UNet_Config) -> None:
        super().__init__()
        self.config = config
        if (config.in_channels_emb != 0 or config.mlp_groups != 1 or config.emb_linear_groups != 1 or config.midblock_attn
                or len(config.attn_levels) or config.dropout != 0):
            raise NotImplementedError("unet_edm2_q4_ddec: only the unconditioned, dense-MLP, attention-free configuration "
                                      "(the dataclass defaults) is implemented")
        kw = dict(mlp_multiplier=config.mlp_multiplier, res_balance=config.res_balance)
        cblock = [config.model_channels * x for x in config.channel_mult]
        cnoise = config.model_channels * config.channel_mult_noise if config.channel_mult_noise is not None else max(cblock)
        cemb = config.model_channels * config.channel_mult_emb if config.channel_mult_emb is not None else max(cblock)
        cemb *= config.mlp_multiplier
        self.num_levels = len(config.channel_mult)
        assert config.in_psd_freqs % config.in_num_freqs == 0
        self.psd_freqs_per_freq = config.in_psd_freqs // config.in_num_freqs
        self.emb_fourier = MPFourier(cnoise)
        self.emb_noise = _MPConv(cnoise, cemb, ())
        self.emb_label = None
        self.emb_label_unconditional = None
        self.logvar_fourier = MPFourier(config.logvar_channels)
        self.logvar_linear = _MPConv(config.logvar_channels, 1, (), disable_weight_norm=True)
        self.logvar_linear.weight.data.fill_(0)

        self.enc = torch.nn.ModuleDict()
        cout = config.in_channels + self.psd_freqs_per_freq * 2
        for level, channels in enumerate(cblock):
            nf = config.in_num_freqs // 2 ** level
            if level == 0:
                cin, cout = cout, channels
                self.enc["conv_in"] = _MPConv(cin, cout, (3, 3), bias=True)
            else:
                self.enc[f"block{level}_down"] = Block(level, cout, cout, cemb, nf, flavor="enc", resample_mode="down", **kw)
            for idx in range(config.num_layers_per_block):
                cin, cout = cout, channels
                self.enc[f"block{level}_layer{idx}"] = Block(level, cin, cout, cemb, nf, flavor="enc", **kw)
        self.dec = torch.nn.ModuleDict()
        skips = [b.out_channels for b in self.enc.values()]
        for level, channels in reversed(list(enumerate(cblock))):
            nf = config.in_num_freqs // 2 ** level
            if level == len(cblock) - 1:
                self.dec[f"block{level}_in0"] = Block(level, cout, cout, cemb, nf, flavor="dec", **kw)
                if config.double_midblock:
                    self.dec[f"block{level}_in1"] = Block(level, cout, cout, cemb, nf, flavor="dec", **kw)
            else:
                self.dec[f"block{level}_up"] = Block(level, cout, cout, cemb, nf, flavor="dec", resample_mode="up", **kw)
            for idx in range(config.num_layers_per_block + 1):
                cin = cout + skips.pop()
                cout = channels
                self.dec[f"block{level}_layer{idx}"] = Block(level, cin, cout, cemb, nf, flavor="dec", **kw)
        self.out_gain = torch.nn.Parameter(torch.zeros([]))
        self.conv_out = _MPConv(cout, config.out_channels, (3, 3))
        self.use_cuda_graphs = True
        self._prep: Dict[str, Tuple[int, Tensor]] = {}
        self._graphs: Dict[tuple, dict] = {}
        self._affine: Dict[object, dict] = {}

    # ---- helpers mirrored from the reference (:238-251) ----
    def get_embeddings(self, emb_in: Tensor, conditioning_mask: Tensor) -> Optional[Tensor]:
        return None

    def get_sigma_loss_logvar(self, sigma: Optional[Tensor] = None) -> Tensor:
        dev = torch.device(self.device)
        s = sigma.detach().to(device=dev, dtype=torch.float32).contiguous().flatten()
        L.require_cuda(self.logvar_linear.weight)
        out = ops.sigma_logvar(s, self.logvar_fourier.freqs.detach().float().contiguous(),
                               self.logvar_fourier.phases.detach().float().contiguous(),
                               self.logvar_linear.weight.detach().contiguous())
        return out.view(-1, 1, 1, 1)

    def get_latent_shape(self, latent_shape) -> tuple:
        m = 2 ** (self.num_levels - 1)
        return tuple(latent_shape[0:2]) + ((latent_shape[2] // m) * m, (latent_shape[3] // m) * m)

    def _apply(self, fn, *args, **kwargs):
        self._prep, self._graphs, self._affine = {}, {}, {}
        return super()._apply(fn, *args, **kwargs)

    # ---- prepared weights (eval mode), refreshed on parameter version change ----
    def _w(self, key: str, conv: _MPConv, gain: Optional[Tensor] = None, pad_rows: int = 0) -> Tensor:
        w = conv.weight
        ver = _ver(w) + (_ver(gain) if gain is not None else 0)
        hit = self._prep.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        g = None if gain is None else gain.detach().float().reshape(1)
        out = ops.weight_prep(w.detach(), gain=g, pad_rows=pad_rows, out=None if hit is None else hit[1])
        self._prep[key] = (ver, out)
        return out

    def _w_in(self) -> Tensor:
        """conv_in (18 -> C, bias): bf16 [C][9][32]; columns 0..17 the scaled weight, column 18 of the centre tap the
        bias (multiplied by the constant-one input channel; the centre tap is never in the zero padding)."""
        conv = self.enc["conv_in"]
        ver = _ver(conv.weight) + _ver(conv.bias)
        hit = self._prep.get("enc.conv_in")
        if hit is not None and hit[0] == ver:
            return hit[1]
        w = conv.weight.detach()
        O, I = w.shape[0], w.shape[1]
        full = ops.weight_prep_z2(w.reshape(O, I, 1, 3, 3), i_stride=32)            # [2*O][9][32], rows O.. a duplicate
        out = full[:O].contiguous()
        out[:, 4, I] = conv.bias.detach().to(torch.bfloat16)
        self._prep["enc.conv_in"] = (ver, out)
        return out

    def _blocks(self):
        for prefix, blocks in (("enc", self.enc), ("dec", self.dec)):
            for name, blk in blocks.items():
                if isinstance(blk, Block):
                    yield f"{prefix}.{name}", blk

    def _emb_scales(self, sigma: Tensor) -> Dict[str, Tensor]:
        """emb = emb_noise(emb_fourier(ln(sigma)/4)) (:279-286, no label mix); c = emb_linear(emb, gain) + 1 per block."""
        B = sigma.numel()
        four = ops.mp_fourier(sigma.log() / 4, self.emb_fourier.freqs.detach().float().contiguous(),
                              self.emb_fourier.phases.detach().float().contiguous())
        st = self._affine.get(B)
        sig = (_ver(self.emb_noise.weight),) + tuple(_ver(b.emb_linear.weight) + _ver(b.emb_gain) for _, b in self._blocks())
        if st is None or st["sig"] != sig:
            dev = sigma.device
            wn = self.emb_noise.weight.detach()
            emb = torch.empty((B, wn.shape[0]), device=dev, dtype=torch.float32)
            d0, m0 = ops.make_affine_descs([dict(w=wn, gain=None, out=emb, groups=1, bias=0.0, normalize=False)], dev)
            entries, outs = [], {}
            for name, blk in self._blocks():
                w = blk.emb_linear.weight.detach().flatten(1)
                out = torch.empty((B, w.shape[0]), device=dev, dtype=torch.float32)
                outs[name] = out
                entries.append(dict(w=w, gain=blk.emb_gain.detach().float().reshape(1), out=out, groups=1, bias=1.0,
                                    normalize=False))
            d1, m1 = ops.make_affine_descs(entries, dev)
            st = dict(sig=sig, emb=emb, d0=d0, m0=m0, entries=entries, outs=outs, d1=d1, m1=m1)
            self._affine[B] = st
        ops.emb_affine(st["d0"], 1, st["m0"], four)
        ops.emb_affine(st["d1"], len(st["entries"]), st["m1"], st["emb"])
        return st["outs"]

    def _run(self, x_in: Tensor, net_in: Tensor, sigma: Tensor, x_ref: Tensor) -> Tensor:
        """The launch schedule of UNet.forward (:253-303) + Block.forward (:120-150)."""
        cfg = self.config
        t = cfg.res_balance
        n = math.sqrt((1 - t) ** 2 + t ** 2)
        ca, cb = (1 - t) / n, t / n
        cvec = self._emb_scales(sigma)
        k = self.psd_freqs_per_freq
        wa, wb = mp_cat_weights(cfg.in_channels, k * cfg.in_channels, cfg.label_balance)
        x = ops.q4_stem(net_in, x_ref, sigma, cfg.sigma_data, wa, wb, k, 32)
        x = ops.mpconv(x, self._w_in(), 3)
        skips = [x]

        def residual_branch(p: str, blk: Block, s: Tensor, resid: Tensor) -> Tensor:
            y0 = ops.mpconv(s, self._w(p + ".conv_res0", blk.conv_res0), 3, epi=L.EPI_SCALE_SILU, scale=cvec[p])
            return ops.mpconv(y0, self._w(p + ".conv_res1", blk.conv_res1), 3, epi=L.EPI_RESIDUAL, alpha=cb, beta=ca,
                              clip=blk.clip_act, residual=resid)

        for name, blk in self.enc.items():
            if not isinstance(blk, Block):
                continue
            p = "enc." + name
            if blk.resample_mode == "down":
                x = ops.avgpool2(x)
            if blk.conv_skip is not None:
                x = ops.mpconv(x, self._w(p + ".conv_skip", blk.conv_skip), 1)
            xn, s = ops.pixnorm_silu(x)
            x = residual_branch(p, blk, s, xn)
            skips.append(x)
        for name, blk in self.dec.items():
            p = "dec." + name
            if "layer" in name:
                skip = skips.pop()
                wa, wb = mp_cat_weights(x.shape[-1], skip.shape[-1], cfg.concat_balance)
                xc, s = ops.cat_silu(x, skip, wa, wb, False)
            elif blk.resample_mode == "up":
                xc, s = ops.cat_silu(x, None, 1.0, 0.0, True)
            else:
                xc = x
                _, s = ops.cat_silu(x, None, 1.0, 0.0, False, need_cat=False)
            resid = xc if blk.conv_skip is None else ops.mpconv(xc, self._w(p + ".conv_skip", blk.conv_skip), 1)
            x = residual_branch(p, blk, s, resid)
        return ops.conv_out(x, self._w("conv_out", self.conv_out, gain=self.out_gain, pad_rows=16), x_in, sigma,
                            cfg.sigma_data, None)

    def forward(self, x_in: Tensor, sigma: Tensor, format=None, embeddings: Optional[Tensor] = None,
                x_ref: Optional[Tensor] = None, perturbed_input: Optional[Tensor] = None) -> Tensor:
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("dualdiffusion_b200 unet_edm2_q4_ddec: backward is not implemented (inference only)")
        if self.training:
            raise NotImplementedError("dualdiffusion_b200 unet_edm2_q4_ddec: train-mode forward is not built")
        dev = torch.device(self.device)
        if dev.type != "cuda":
            raise RuntimeError("dualdiffusion_b200 unet_edm2_q4_ddec has no CPU path: move the module to a CUDA device (B200)")
        if x_ref is None:
            raise ValueError("x_ref (the PSD reference, (B, C, in_psd_freqs, W)) is required")
        cfg = self.config
        B, C, Fq, W = x_in.shape
        if C != cfg.in_channels or Fq != cfg.in_num_freqs or tuple(x_ref.shape) != (B, C, cfg.in_psd_freqs, W):
            raise ValueError(f"expected x_in (B, {cfg.in_channels}, {cfg.in_num_freqs}, W) and x_ref (B, {cfg.in_channels}, "
                             f"{cfg.in_psd_freqs}, W), got {tuple(x_in.shape)} and {tuple(x_ref.shape)}")
        x32 = x_in.detach().to(device=dev, dtype=torch.float32).contiguous()
        n32 = x32 if perturbed_input is None else perturbed_input.detach().to(device=dev, dtype=torch.float32).contiguous()
        xr = x_ref.detach().to(device=dev, dtype=torch.float32).contiguous()
        sg = sigma.detach().to(device=dev, dtype=torch.float32).flatten().contiguous()
        if sg.numel() == 1 and B > 1:
            sg = sg.expand(B).contiguous()
        with torch.no_grad():
            if not self.use_cuda_graphs:
                return self._run(x32, n32, sg, xr)
            sig = tuple(_ver(p) for p in self.parameters())
            key = (tuple(x32.shape), perturbed_input is not None)
            gs = self._graphs.get(key)
            if gs is None or gs["sig"] != sig:
                st = dict(x=x32.clone(), n=n32.clone() if perturbed_input is not None else None, s=sg.clone(), r=xr.clone())
                n_st = st["n"] if st["n"] is not None else st["x"]
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):              # warm-up outside capture (weight prep, descriptor tables)
                    self._run(st["x"], n_st, st["s"], st["r"])
                torch.cuda.current_stream(dev).wait_stream(side)
                before = ops.launch_count
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = self._run(st["x"], n_st, st["s"], st["r"])
                gs = dict(sig=sig, graph=graph, static=st, out=out, launches=ops.launch_count - before)
                self._graphs[key] = gs
            st = gs["static"]
            st["x"].copy_(x32)
            if st["n"] is not None:
                st["n"].copy_(n32)
            st["s"].copy_(sg)
            st["r"].copy_(xr)
            gs["graph"].replay()
            ops.launch_count += gs["launches"]
            return gs["out"].clone()
