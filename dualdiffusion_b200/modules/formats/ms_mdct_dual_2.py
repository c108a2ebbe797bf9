"""B200-native drop-in for the second lineage of the reference's dual format
(`modules.formats.ms_mdct_dual_2.MS_MDCT_DualFormat`, src/modules/formats/ms_mdct_dual_2.py:34-291),
the format the dae_p1 / dae_q1 / ddec_p* / ddec_q4 trainers build their inputs with (training/module_trainers/*.py).

Same config dataclass (field for field) and methods: `raw_to_mel_spec`, `mel_spec_to_linear`, `raw_to_mdct`, `mdct_to_raw`,
`raw_to_mdct_phase_psd`, the normalize / unnormalize helpers and the shape helpers.

* mel side: one 4096-sample STFT per Hann-power window (`dd_stft_mel`: magnitude, 1/mel-density and the RMS-normalised
  triangular filterbank in one launch per window), then `dd_mel_blend` applies the per-filter window weights, the 0.25 power
  and the output affine in one pass; `mel_spec_to_linear` is `dd_mel_linearize` + the minimum-norm inverse filterbank as one
  fp32 GEMM (`dd_gemm_f32`, density^0.5 and the scales folded into the matrix).
* MDCT side (utils/mdct/functional.py:9-150, sin window, hop = win/2, reflect padding): the frame transform is linear, so
  each direction is one fp32 GEMM against a host-built (fp64) matrix -- forward frames gathered from the raw signal by the
  GEMM's operand loads, inverse followed by the 50 % overlap-add (`dd_mdct_ola`); `dd_mdct_phase_psd` splits the complex
  coefficients into the normalised phase / PSD pair.
No CPU fallback.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Literal, Optional, Sequence

import numpy as np
import torch

from ... import _lib as L
from ... import ops
from .format import DualDiffusionFormat, DualDiffusionFormatConfig
from .ms_mdct_dual import FrequencyScale


@dataclass
class MS_MDCT_DualFormatConfig(DualDiffusionFormatConfig):
    """ms_mdct_dual_2.py:34-91 (same fields and defaults)."""
    sample_rate: int = 32000
    num_raw_channels: int = 2
    default_raw_length: int = 1408768
    raw_to_mdct_scale: float = 0.00395184212251821011433253029603
    mdct_psd_scale: float = 0.07179056842448940381561506832112
    mdct_psd_offset: float = -0.1806843343919556
    mdct_psd_exponent: float = 0.25
    mdct_phase_scale: float = 1
    mdct_window_len: int = 512
    mdct_window_func: Literal["sin", "kaiser_bessel_derived", "vorbis"] = "sin"
    raw_to_mel_spec_scale: float = 0.48693139085749312574067728443989
    raw_to_mel_spec_offset: float = -1.530891040808645
    mel_spec_to_linear_scale: float = 15.11100987193986714324861053997
    mel_spec_to_linear_offset: float = 0
    ms_abs_exponent: float = 0.25
    ms_freq_min: float = 0
    ms_num_filters: int = 256
    ms_ideal_num_filter_bins: float = 3
    ms_window_length: int = 4096
    ms_blend_sharpness: float = 30
    ms_window_exponents: Sequence[float] = (9, 32, 112)

    @property
    def mdct_num_frequencies(self) -> int:
        return self.mdct_window_len // 2

    @property
    def mdct_frame_hop_length(self) -> int:
        return self.mdct_window_len // 2

    @property
    def ms_num_stft_bins(self) -> int:
        return self.ms_window_length // 2 + 1

    @property
    def ms_hop_length(self) -> int:
        return self.mdct_frame_hop_length

    @property
    def ms_width_alignment(self) -> int:
        return self.mdct_frame_hop_length // 2

    @property
    def ms_freq_max(self) -> float:
        return self.sample_rate / 2


def _mdct_window(name: str, n: int) -> np.ndarray:
    """utils/mdct/windows.py: sin (:95-116), vorbis (:59-92), kaiser_bessel_derived (:10-56, beta 12), in fp64."""
    arg = (np.arange(n, dtype=np.float64) + 0.5) / n * np.pi
    if name == "sin":
        return np.sin(arg)
    if name == "vorbis":
        return np.sin(np.pi / 2.0 * np.sin(arg) ** 2)
    if name == "kaiser_bessel_derived":
        kais = torch.kaiser_window(n // 2 + 1, True, 12.0, dtype=torch.float64)
        cs = torch.cumsum(kais, dim=-1)
        half = torch.sqrt(cs[:-1] / cs[-1]).numpy()
        return np.concatenate([half, half[::-1]])
    raise ValueError(f"Unsupported mdct window function: {name}. Supported functions are 'sin', 'kaiser_bessel_derived', and 'vorbis'.")


class MS_MDCT_DualFormat(DualDiffusionFormat):

    config_class = MS_MDCT_DualFormatConfig

    def __init__(self, config: MS_MDCT_DualFormatConfig) -> None:
        super().__init__()
        self.config = config
        self.ms_freq_scale = FrequencyScale(config.ms_freq_min, config.ms_freq_max, config.sample_rate, config.ms_num_stft_bins,
                                            config.ms_num_filters, "slaney")
        c = config
        # ---- buffers of the reference constructor (:98-156), host side ----
        hann = torch.hann_window(c.ms_window_length, periodic=True)
        ms_windows = torch.stack([hann ** e for e in c.ms_window_exponents], dim=0)
        self.ms_windows = ms_windows / ms_windows.pow(2).mean(dim=1, keepdim=True).pow(0.5)
        mel_freqs = self.ms_freq_scale.get_unscaled(c.ms_num_filters + 2)
        self.ms_filter_center_hz = mel_freqs[1:-1]
        self.ms_filter_bandwidths = mel_freqs[2:] - mel_freqs[:-2]
        num_filter_bins = self.ms_filter_bandwidths / c.sample_rate * c.ms_num_stft_bins * 2
        self.ms_ideal_filter_widths = (c.ms_ideal_num_filter_bins / num_filter_bins * c.ms_window_length).to(torch.float64)
        fb = self.ms_freq_scale.get_filters()
        self.ms_filters = fb / fb.pow(2).mean(dim=0, keepdim=True).pow(0.5)
        self.ms_window_widths = torch.tensor([2 * np.arccos(2 ** (-1 / e)) / np.pi * 2 * c.ms_window_length
                                              for e in c.ms_window_exponents], dtype=torch.float64)
        ww = torch.zeros((c.ms_num_filters, ms_windows.shape[0]), dtype=torch.float32)
        for i in range(c.ms_num_filters):
            w = (-c.ms_blend_sharpness * (self.ms_ideal_filter_widths[i] / self.ms_window_widths).log() ** 2).exp()
            ww[i] = (w / w.sum()).to(torch.float32)
        self.ms_filter_window_weights = ww
        self.ms_stft_mel_density = (1127.0 / (700.0 + torch.linspace(0, c.sample_rate / 2, c.ms_num_stft_bins))).view(1, 1, -1, 1)
        self.mdct_hz = (torch.arange(c.mdct_num_frequencies) + 0.5) * c.sample_rate / c.mdct_window_len
        self.mdct_mel_density = (1127.0 / (700.0 + self.mdct_hz)).view(1, 1, -1, 1)
        _mdct_window(c.mdct_window_func, c.mdct_window_len)            # raises for an unknown window name, as the reference
        self._dev_cache = {}

    # ---- shapes (:180-196, :233-243) ----
    def _get_ms_shape(self, raw_shape: tuple) -> tuple:
        return tuple(raw_shape[:-1]) + (self.config.ms_num_filters, 1 + raw_shape[-1] // self.config.ms_hop_length)

    def _get_ms_raw_shape(self, mel_spec_shape: tuple) -> tuple:
        return tuple(mel_spec_shape[:-2]) + ((mel_spec_shape[-1] - 1) * self.config.ms_hop_length,)

    def get_raw_crop_width(self, raw_length: Optional[int] = None) -> int:
        raw_length = raw_length or self.config.default_raw_length
        n = self._get_ms_shape((1, raw_length))[-1]
        n = n // self.config.ms_width_alignment * self.config.ms_width_alignment
        return self._get_ms_raw_shape((1, n))[-1]

    def get_mel_spec_shape(self, bsz: int = 1, raw_length: Optional[int] = None) -> tuple:
        return self._get_ms_shape((bsz, self.config.num_raw_channels, self.get_raw_crop_width(raw_length)))

    def _get_mdct_raw_crop_width(self, raw_length: Optional[int] = None) -> int:
        c = self.config
        raw_length = raw_length or c.default_raw_length
        return raw_length // c.mdct_window_len // c.ms_width_alignment * c.ms_width_alignment * c.mdct_window_len + c.mdct_window_len

    def get_mdct_shape(self, bsz: int = 1, raw_length: Optional[int] = None) -> tuple:
        n_bins = self.config.mdct_num_frequencies
        return (bsz, self.config.num_raw_channels, n_bins, self.get_raw_crop_width(raw_length=raw_length) // n_bins + 1)

    # ---- device tables ----
    def _tables(self, device: torch.device) -> dict:
        key = str(device)
        t = self._dev_cache.get(key)
        if t is not None:
            return t
        c = self.config
        n_fft, n = c.ms_window_length, c.ms_window_length // 2
        density = self.ms_stft_mel_density.flatten()
        coef = (1.0 / math.sqrt(n_fft)) / density                          # torch.stft(normalized=True) and 1/mel-density
        fb = self.ms_filters
        starts, counts, offsets, weights = [], [], [], []
        for f in range(fb.shape[1]):
            nz = torch.nonzero(fb[:, f]).flatten()
            lo, hi = (int(nz[0]), int(nz[-1]) + 1) if nz.numel() else (0, 0)
            starts.append(lo); counts.append(hi - lo); offsets.append(len(weights))
            weights.extend(fb[lo:hi, f].tolist())
        k = np.arange(n, dtype=np.float64)
        tw = np.exp(-2j * np.pi * k / n)
        tw_half = np.exp(-2j * np.pi * np.arange(n + 1, dtype=np.float64) / n_fft)
        to = lambda x, dt: torch.as_tensor(x, dtype=dt).to(device).contiguous()
        # mel_spec_to_linear (:218-223): lin = (scale * clip(mel - offset/scale, 0)) ** (1/e); psd = pinv(slaney bank) @ lin *
        # density^0.5, last bin dropped, (psd + lin_offset) / lin_scale
        a64 = self.ms_freq_scale.get_filters().double().numpy().T                       # (filters, bins)
        gram = a64 @ a64.T
        if np.linalg.cond(gram) > 1e12:
            raise ValueError("mel filterbank is rank deficient: min-norm inverse undefined")
        pinv = a64.T @ np.linalg.inv(gram)                                              # (bins, filters)
        pinv = pinv * np.sqrt(density.double().numpy())[:, None] * (c.raw_to_mel_spec_scale ** (1.0 / c.ms_abs_exponent))
        pinv = pinv[:-1] / c.mel_spec_to_linear_scale
        # MDCT matrices (utils/mdct/functional.py): W = 2N samples per frame, N bins
        W = c.mdct_window_len
        N = W // 2
        w = _mdct_window(c.mdct_window_func, W)
        j = np.arange(W, dtype=np.float64)[:, None]
        kk = np.arange(N, dtype=np.float64)[None, :]
        scaling = 1.0 / math.sqrt(W * N)
        fwd = (w[:, None] * np.exp(-1j * np.pi * j / W) * np.exp(-2j * np.pi * j * kk / W)
               * np.exp(-1j * np.pi / W * (N + 1) * (kk + 0.5)) * scaling)                # [W samples][N bins], complex
        dens = self.mdct_mel_density.flatten().double().numpy()
        fwd_scaled = fwd / (dens * c.raw_to_mdct_scale)[None, :]
        # inverse (real input): frame[j] = 2 w[j] Re(post[j] * sum_k pre[k] X[k] e^{-2 pi i jk / W}) / scaling
        pre = np.exp(-1j * np.pi / (2 * N) * (N + 1) * np.arange(N, dtype=np.float64))
        post = np.exp(-1j * np.pi / (2 * N) * (np.arange(W, dtype=np.float64) + 0.5 + N / 2)) / N
        inv = 2.0 * w[None, :] * np.real(pre[:, None] * np.exp(-2j * np.pi * kk.T * j.T / W) * post[None, :]) / scaling
        inv = inv * (dens * c.raw_to_mdct_scale)[:, None]                               # [N bins][W samples]
        f32 = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32).to(device)
        t = dict(windows=[w_.float().to(device).contiguous() for w_ in self.ms_windows],
                 coef=coef.float().to(device).contiguous(),
                 ww=self.ms_filter_window_weights.to(device).contiguous(),
                 tw=to(np.stack([tw.real, tw.imag], -1), torch.float32),
                 tw_half=to(np.stack([tw_half.real, tw_half.imag], -1), torch.float32),
                 fb=dict(start=to(starts, torch.int32), count=to(counts, torch.int32), offset=to(offsets, torch.int32),
                         weight=to(weights, torch.float32)),
                 pinv=f32(pinv), N=N,
                 fwd=f32(np.concatenate([fwd_scaled.real.T, fwd_scaled.imag.T], 0)),     # [2N rows (re | im)][W]
                 fwd_raw=f32(np.concatenate([fwd.real.T, fwd.imag.T], 0)),
                 inv=f32(inv), inv_density=f32(1.0 / dens))
        self._dev_cache[key] = t
        return t

    # ---- mel side ----
    @torch.no_grad()
    def raw_to_mel_spec(self, raw_samples: torch.Tensor) -> torch.Tensor:
        """:198-216 -> (B, C, ms_num_filters, 1 + L // hop)."""
        L.require_cuda(raw_samples)
        c = self.config
        t = self._tables(raw_samples.device)
        B, C, n = raw_samples.shape
        raw = raw_samples.detach().float().contiguous().view(B * C, n)
        n_win = len(t["windows"])
        T = 1 + n // c.ms_hop_length
        mels = torch.empty((n_win, B * C, c.ms_num_filters, T), device=raw.device, dtype=torch.float32)
        for i, w in enumerate(t["windows"]):
            ops.stft_mel(raw, w, t["tw"], t["tw_half"], c.ms_window_length, c.ms_hop_length, t["fb"], 1.0, 0.0, 1.0,
                         coef1=t["coef"], out=mels[i])
        out = ops.mel_blend(mels, t["ww"], c.ms_abs_exponent, c.raw_to_mel_spec_offset, 1.0 / c.raw_to_mel_spec_scale)
        return out.view(B, C, c.ms_num_filters, T)

    @torch.no_grad()
    def mel_spec_to_linear(self, mel_spec: torch.Tensor) -> torch.Tensor:
        """:218-223 -> (B, C, ms_num_stft_bins - 1, T)."""
        L.require_cuda(mel_spec)
        c = self.config
        if c.raw_to_mel_spec_scale <= 0:
            raise ValueError("raw_to_mel_spec_scale must be positive")
        t = self._tables(mel_spec.device)
        lin = ops.mel_linearize(mel_spec.detach().float().contiguous(), c.raw_to_mel_spec_offset / c.raw_to_mel_spec_scale,
                                1.0 / c.ms_abs_exponent)
        Bm, Cm, F, T = lin.shape
        nb = t["pinv"].shape[0]
        psd = torch.empty((Bm, Cm, nb, T), device=lin.device, dtype=torch.float32)
        ops.gemm_f32(t["pinv"], (F, 1, 0), lin, (T, 1, F * T), psd, (T, 1, nb * T), nb, T, F, Bm * Cm)
        if c.mel_spec_to_linear_offset != 0:
            psd = psd + c.mel_spec_to_linear_offset / c.mel_spec_to_linear_scale
        return psd

    # ---- MDCT side ----
    def _mclt_rows(self, raw_samples: torch.Tensor, mat: str) -> torch.Tensor:
        """(B, C, L) -> [B*C][2N][T]: real rows then imaginary rows of the MCLT; T = ceil(L / N) + 1 frames at hop N over
        the signal reflect-padded by N on the left (functional.py:36-43: the extra right padding only feeds the dropped frame)."""
        L.require_cuda(raw_samples)
        t = self._tables(raw_samples.device)
        N = t["N"]
        B, C, n = raw_samples.shape
        raw = raw_samples.detach().float().contiguous().view(B * C, n)
        T = -(-n // N) + 1
        y = torch.empty((B * C, 2 * N, T), device=raw.device, dtype=torch.float32)
        return ops.gemm_f32(t[mat], (2 * N, 1, 0), raw, (0, 0, n), y, (T, 1, 2 * N * T), 2 * N, T, 2 * N, B * C, gather=(N, N, n))

    @torch.no_grad()
    def raw_to_mdct(self, raw_samples: torch.Tensor, random_phase_augmentation: bool = False) -> torch.Tensor:
        """:245-254 -> (B, C, N, T)."""
        if random_phase_augmentation:
            raise NotImplementedError("random_phase_augmentation (training-time augmentation) is not implemented")
        B, C, _ = raw_samples.shape
        y = self._mclt_rows(raw_samples, "fwd")
        N, T = y.shape[1] // 2, y.shape[2]
        return y[:, :N].reshape(B, C, N, T)

    @torch.no_grad()
    def mdct_to_raw(self, mdct: torch.Tensor) -> torch.Tensor:
        """:256-261 -> (B, C, (T-1)*N)."""
        L.require_cuda(mdct)
        t = self._tables(mdct.device)
        N = t["N"]
        B, C, Nb, T = mdct.shape
        if Nb != N:
            raise ValueError(f"expected {N} MDCT bins, got {Nb}")
        x = mdct.detach().float().reshape(B * C, N, T).contiguous()
        y = torch.empty((B * C, T, 2 * N), device=x.device, dtype=torch.float32)
        ops.gemm_f32(x, (1, T, N * T), t["inv"], (2 * N, 1, 0), y, (2 * N, 1, T * 2 * N), T, 2 * N, N, B * C)   # A = x^T as strides
        return ops.mdct_ola(y).view(B, C, -1)

    def normalize_psd(self, mdct_psd: torch.Tensor) -> torch.Tensor:
        return (mdct_psd + self.config.mdct_psd_offset) / self.config.mdct_psd_scale

    def unnormalize_psd(self, norm_mdct_psd: torch.Tensor) -> torch.Tensor:
        return norm_mdct_psd * self.config.mdct_psd_scale - self.config.mdct_psd_offset

    def normalize_phase(self, mdct_phase: torch.Tensor) -> torch.Tensor:
        return mdct_phase / self.config.mdct_phase_scale

    def unnormalize_phase(self, norm_mdct_phase: torch.Tensor) -> torch.Tensor:
        return norm_mdct_phase * self.config.mdct_phase_scale

    @torch.no_grad()
    def raw_to_mdct_phase_psd(self, raw_samples: torch.Tensor, random_phase_augmentation: bool = False):
        """:275-289 -> (normalised phase, normalised PSD), each (B, C, N, T)."""
        if random_phase_augmentation:
            raise NotImplementedError("random_phase_augmentation (training-time augmentation) is not implemented")
        c = self.config
        B, C, _ = raw_samples.shape
        y = self._mclt_rows(raw_samples, "fwd_raw")
        t = self._tables(raw_samples.device)
        phase, psd = ops.mdct_phase_psd(y, t["inv_density"], c.mdct_psd_exponent, c.mdct_psd_offset, 1.0 / c.mdct_psd_scale,
                                        math.sqrt(2.0) / c.mdct_phase_scale)
        N, T = phase.shape[1], phase.shape[2]
        return phase.view(B, C, N, T), psd.view(B, C, N, T)

    def mdct_phase_psd_to_raw(self, mdct_phase: torch.Tensor, mdct_psd: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError()                                    # as the reference (:297)
