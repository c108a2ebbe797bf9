"""Operator-level mirror of the reference's magnitude-preserving tool set (src/modules/mp_tools.py) for the
operators on the denoising hot path.  Same names, argument meaning and error behaviour; tensors keep the
reference's logical NCHW shape (any memory format / float dtype) and are routed through the C-ABI kernels as
NHWC bf16.  The fused UNet schedule (modules/unets/unet_edm2_b4.py) does not go through these wrappers — they
exist so reference code written against `modules.mp_tools` keeps working, and so the parity tests read like
tests of the reference's operators.  CPU tensors raise: there is no CPU path.
"""
from __future__ import annotations

import math
from typing import Literal, Optional, Tuple, Union

import torch

from .. import _lib as L
from .. import ops

Tensor = torch.Tensor


def _to_nhwc(x: Tensor) -> Tensor:
    L.require_cuda(x)
    if x.ndim != 4:
        raise ValueError(f"expected a 4-D (B,C,H,W) tensor, got shape {tuple(x.shape)}")
    return x.detach().permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()


def _from_nhwc(y: Tensor, like: Tensor) -> Tensor:
    return y.permute(0, 3, 1, 2).to(like.dtype)


def normalize(x: Tensor, dim: Optional[Union[tuple, list, int]] = None, eps: float = 1e-4) -> Tensor:
    """mp_tools.py:42-49.  Supported reductions: pixel norm (4-D, dim=1) and per-row over all remaining
    dims (dim=None: weights, embeddings)."""
    if eps != 1e-4:
        raise NotImplementedError("normalize: only the reference default eps=1e-4 is implemented")
    L.require_cuda(x)
    if x.ndim == 4 and dim in (1, [1], (1,)):
        xn, _ = ops.pixnorm_silu(_to_nhwc(x))
        return _from_nhwc(xn, x)
    if dim is None and x.ndim >= 2:
        fan = x[0].numel()
        src = x.detach().contiguous()
        if src.dtype not in (torch.float32, torch.bfloat16):
            src = src.float()
        out = ops.weight_prep(src.view(src.shape[0], fan, 1), gain_host=math.sqrt(fan), normalize=True,
                              fmt=L.WFMT_F32_OIT)
        return out.view(x.shape).to(x.dtype)
    raise NotImplementedError(f"normalize: reduction dim={dim} on a {x.ndim}-D tensor is not on the hot path")


def resample_2d(x: Tensor, mode: Literal["keep", "down", "up"] = "keep", ratio: int = 2,
                filtering: str = "nearest") -> Tensor:
    """mp_tools.py:71-79."""
    if mode == "keep":
        return x
    if ratio != 2 or filtering != "nearest":
        raise NotImplementedError("resample_2d: only ratio=2, nearest is implemented")
    if mode == "down":
        return _from_nhwc(ops.avgpool2(_to_nhwc(x)), x)
    if mode == "up":
        xc, _ = ops.cat_silu(_to_nhwc(x), None, 1.0, 0.0, True)
        return _from_nhwc(xc, x)
    raise ValueError(f"unknown resample mode {mode}")


def mp_silu(x: Tensor) -> Tensor:
    """mp_tools.py:268-269."""
    _, s = ops.cat_silu(_to_nhwc(x), None, 1.0, 0.0, False, need_cat=False)
    return _from_nhwc(s, x)


def mp_sum(a: Tensor, b: Tensor, t: Union[Tensor, float] = 0.5) -> Tensor:
    """mp_tools.py:274-279 (float t)."""
    if isinstance(t, Tensor):
        if t.numel() != 1:
            raise NotImplementedError("mp_sum: tensor-valued t is fused into dd_conv_out / dd_label_embedding")
        t = float(t)
    n = math.sqrt((1 - t) ** 2 + t ** 2)
    return _from_nhwc(ops.axpby(_to_nhwc(a), _to_nhwc(b), (1 - t) / n, t / n), a)


def mp_cat_weights(na: int, nb: int, t: float = 0.5) -> Tuple[float, float]:
    """Per-source scales of mp_cat (mp_tools.py:294-301)."""
    c = math.sqrt((na + nb) / ((1 - t) ** 2 + t ** 2))
    return c / math.sqrt(na) * (1 - t), c / math.sqrt(nb) * t


def mp_cat(a: Tensor, b: Tensor, dim: int = 1, t: float = 0.5) -> Tensor:
    """mp_tools.py:294-301 (channel concat)."""
    if dim != 1:
        raise NotImplementedError("mp_cat: only dim=1 is implemented")
    wa, wb = mp_cat_weights(a.shape[1], b.shape[1], t)
    xc, _ = ops.cat_silu(_to_nhwc(a), _to_nhwc(b), wa, wb, False)
    return _from_nhwc(xc, a)


class MPFourier(torch.nn.Module):
    """mp_tools.py:316-330."""

    def __init__(self, num_channels: int, bandwidth: float = 1.0, eps: float = 1e-3) -> None:
        super().__init__()
        self.register_buffer("freqs", torch.pi * torch.linspace(0, 1 - eps, num_channels).erfinv() * bandwidth)
        self.register_buffer("phases", torch.pi / 2 * (torch.arange(num_channels) % 2 == 0).float())

    def forward(self, x: Tensor) -> Tensor:
        if x.ndim != 1:
            raise NotImplementedError("MPFourier: only the 1-D (per-sample scalar) path is on the hot path")
        L.require_cuda(x, self.freqs)
        y = ops.mp_fourier(x.detach().float().contiguous(), self.freqs.float().contiguous(), self.phases.float().contiguous())
        return y.to(x.dtype)


class MPConv(torch.nn.Module):
    """mp_tools.py:332-378.  `kernel=()` is a linear layer, otherwise a stride-1 'same' convolution."""

    def __init__(self, in_channels: int, out_channels: int, kernel: Tuple[int, ...], groups: int = 1, stride: int = 1,
                 disable_weight_norm: bool = False, bias: bool = False) -> None:
        super().__init__()
        if stride != 1:
            raise NotImplementedError("MPConv: stride != 1 is not used on the hot path")
        if bias:
            raise NotImplementedError("MPConv: bias is not used by unet_edm2_b4")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.groups = groups
        self.stride = stride
        self.disable_weight_norm = disable_weight_norm
        self.weight = torch.nn.Parameter(torch.randn(out_channels, in_channels // groups, *kernel))
        self.weight.conv_groups = groups
        self.bias = None
        self._prepped = None
        self._prepped_key = None

    def _prep(self, gain: Union[float, Tensor]) -> Tensor:
        gain_t = gain.detach().float() if isinstance(gain, Tensor) else None
        gain_h = 1.0 if isinstance(gain, Tensor) else float(gain)
        normalize = self.training and not self.disable_weight_norm
        def ver(t: Tensor) -> int:
            return 0 if t.is_inference() else t._version        # inference tensor: immutable
        key = (ver(self.weight), self.weight.data_ptr(), normalize, gain_h,
               None if gain_t is None else (gain_t.data_ptr(), ver(gain)))
        if self._prepped is None or key != self._prepped_key:
            self._prepped = ops.weight_prep(self.weight.detach(), gain=gain_t, gain_host=gain_h, normalize=normalize)
            self._prepped_key = key
        return self._prepped

    def forward(self, x: Tensor, gain: Union[float, Tensor] = 1.0) -> Tensor:
        L.require_cuda(x, self.weight)
        w = self.weight
        if w.ndim == 2:
            xin = x.detach().float().contiguous()
            out = torch.empty((xin.shape[0], w.shape[0]), device=x.device, dtype=torch.float32)
            gain_t = gain.detach().float() if isinstance(gain, Tensor) else torch.tensor(float(gain), device=x.device)
            descs, max_o = ops.make_affine_descs(
                [dict(w=w.detach(), gain=gain_t, out=out, groups=1, bias=0.0,
                      normalize=self.training and not self.disable_weight_norm)], x.device)
            ops.emb_affine(descs, 1, max_o, xin)
            return out.to(x.dtype)
        k = w.shape[-1]
        if w.shape[-2] != k or k not in (1, 3):
            raise NotImplementedError(f"MPConv: kernel {tuple(w.shape[2:])} is not on the hot path (1x1, 3x3)")
        y = ops.mpconv(_to_nhwc(x), self._prep(gain), k, self.groups)
        return _from_nhwc(y, x)

    @torch.no_grad()
    def normalize_weights(self) -> None:
        """mp_tools.py:375-378."""
        if not self.disable_weight_norm:
            self.weight.copy_(normalize(self.weight))
