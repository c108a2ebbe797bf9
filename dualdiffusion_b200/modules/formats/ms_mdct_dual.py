"""B200-native drop-in for the mel-spectrogram side of the reference's live format
(`modules.formats.ms_mdct_dual.MS_MDCT_DualFormat`, src/modules/formats/ms_mdct_dual.py:36-257):
`raw_to_mel_spec`, the shape helpers, and `ms_freq_scale.get_unscaled` (the UNet's positional channel reads it,
unet_edm2_b4.py:246).  The two-window magnitude STFT, per-bin blend, 1/mel-density, slaney mel filterbank and the
output affine run as ONE kernel launch (`dd_stft_mel` with a second window).  The MDCT side (`raw_to_mdct`,
`raw_to_mdct_psd`, `mdct_to_raw`, `mel_spec_to_mdct_psd`) runs each transform as one fp32 library GEMM
against a host-built (fp64) matrix with framing / |.| / overlap-add kernels around it (csrc/mdct.cu).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Literal, Optional

import numpy as np
import torch

from ... import _lib as L
from ... import ops
from .format import DualDiffusionFormat, DualDiffusionFormatConfig


@dataclass
class MS_MDCT_DualFormatConfig(DualDiffusionFormatConfig):
    """ms_mdct_dual.py:36-66 (same fields and defaults)."""
    sample_rate: int = 32000
    num_raw_channels: int = 2
    default_raw_length: int = 1408768
    raw_to_mel_spec_scale: float = 50
    raw_to_mel_spec_offset: float = 0
    mel_spec_to_mdct_psd_scale: float = 0.18
    mel_spec_to_mdct_psd_offset: float = 0
    mdct_to_raw_scale: float = 2
    raw_to_mdct_scale: float = 12.1
    mdct_window_len: int = 512
    mdct_window_func: Literal["sin", "kaiser_bessel_derived"] = "kaiser_bessel_derived"
    mdct_psd_num_bins: int = 2048
    mdct_dual_channel: bool = False
    ms_abs_exponent: float = 1
    ms_filter_shape: Literal["triangular", "cos"] = "triangular"
    ms_freq_min: float = 0
    ms_width_alignment: int = 128
    ms_num_frequencies: int = 256
    ms_step_size_ms: int = 8
    ms_window_duration_ms: int = 128
    ms_padded_duration_ms: int = 128
    ms_window_exponent_low: float = 17
    ms_window_exponent_high: Optional[float] = 58
    ms_window_periodic: bool = True
    ms_window_func: Literal["hann", "blackman_harris"] = "blackman_harris"

    @property
    def ms_num_stft_bins(self) -> int:
        return self.ms_frame_padded_length // 2 + 1

    @property
    def ms_frame_padded_length(self) -> int:
        return int(self.ms_padded_duration_ms / 1000.0 * self.sample_rate)

    @property
    def ms_win_length(self) -> int:
        return int(self.ms_window_duration_ms / 1000.0 * self.sample_rate)

    @property
    def ms_frame_hop_length(self) -> int:
        return int(self.ms_step_size_ms / 1000.0 * self.sample_rate)


class FrequencyScale:
    """The part of frequency_scale.py:85-149 the hot path touches: mel points (`get_unscaled`) and the filterbank."""

    def __init__(self, freq_min: float, freq_max: float, sample_rate: int, num_stft_bins: int, num_filters: int,
                 filter_norm: Optional[str] = None) -> None:
        self.freq_min, self.freq_max, self.sample_rate = freq_min, freq_max, sample_rate
        self.num_stft_bins, self.num_filters, self.filter_norm = num_stft_bins, num_filters, filter_norm

    def get_unscaled(self, num_points: int, device=None) -> torch.Tensor:
        lo = 2595.0 * np.log10(1.0 + self.freq_min / 700.0)
        hi = 2595.0 * np.log10(1.0 + self.freq_max / 700.0)
        return 700.0 * (10.0 ** (torch.linspace(lo, hi, num_points, device=device) / 2595.0) - 1.0)

    def get_filters(self) -> torch.Tensor:
        stft_freqs = torch.linspace(0, self.sample_rate / 2, self.num_stft_bins)
        f_pts = self.get_unscaled(self.num_filters + 2)
        f_diff = f_pts[1:] - f_pts[:-1]
        slopes = f_pts.unsqueeze(0) - stft_freqs.unsqueeze(1)
        fb = torch.max(torch.zeros(1), torch.min((-1.0 * slopes[:, :-2]) / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))
        if self.filter_norm == "slaney":
            fb = fb * (2.0 / (f_pts[2:self.num_filters + 2] - f_pts[:self.num_filters])).unsqueeze(0)
        return fb


def _window(cfg: MS_MDCT_DualFormatConfig, exponent: float) -> torch.Tensor:
    """ms_mdct_dual.py:90-101."""
    n = cfg.ms_win_length
    if cfg.ms_window_func == "blackman_harris":
        x = torch.arange(n) / n * 2 * torch.pi                                  # utils/mclt.py:69-71
        w = 0.35875 - 0.48829 * torch.cos(x) + 0.14128 * torch.cos(2 * x) - 0.01168 * torch.cos(3 * x)
    elif cfg.ms_window_func == "hann":
        w = torch.hann_window(n, periodic=cfg.ms_window_periodic)
    else:
        raise ValueError(f"Unsupported window function: {cfg.ms_window_func}")
    return w ** exponent


class MS_MDCT_DualFormat(DualDiffusionFormat):

    # resolved by from_pretrained (module.py:72); explicit because this file's annotations are strings
    config_class = MS_MDCT_DualFormatConfig

    def __init__(self, config: MS_MDCT_DualFormatConfig) -> None:
        super().__init__()
        self.config = config
        if config.ms_win_length != config.ms_frame_padded_length:
            raise NotImplementedError("window shorter than n_fft is not implemented (reference default: equal)")
        if config.ms_filter_shape != "triangular":
            raise NotImplementedError("only the triangular filter shape is implemented")
        self.ms_freq_scale = FrequencyScale(config.ms_freq_min, config.sample_rate / 2, config.sample_rate,
                                            config.ms_num_stft_bins, config.ms_num_frequencies, "slaney")
        self.ms_lowest_filter_freq = float(self.ms_freq_scale.get_unscaled(config.ms_num_frequencies + 2)[1])
        self._dev_cache = {}

    # ---- shapes (ms_mdct_dual.py:211-228) ----
    def _get_ms_shape(self, raw_shape: tuple) -> tuple:
        c = self.config
        num_frames = 1 + (raw_shape[-1] + c.ms_frame_padded_length - c.ms_win_length) // c.ms_frame_hop_length
        return tuple(raw_shape[:-1]) + (c.ms_num_frequencies, num_frames)

    def _get_ms_raw_shape(self, mel_spec_shape: tuple) -> tuple:
        c = self.config
        audio_len = (mel_spec_shape[-1] - 1) * c.ms_frame_hop_length + c.ms_win_length - c.ms_frame_padded_length
        return tuple(mel_spec_shape[:-2]) + (audio_len,)

    def get_raw_crop_width(self, raw_length: Optional[int] = None) -> int:
        raw_length = raw_length or self.config.default_raw_length
        mel_spec_len = self._get_ms_shape((1, raw_length))[-1]
        mel_spec_len = mel_spec_len // self.config.ms_width_alignment * self.config.ms_width_alignment
        return self._get_ms_raw_shape((1, mel_spec_len))[-1]

    def get_mel_spec_shape(self, bsz: int = 1, raw_length: Optional[int] = None) -> tuple:
        return self._get_ms_shape((bsz, self.config.num_raw_channels, self.get_raw_crop_width(raw_length)))

    # ---- device tables ----
    def _tables(self, device: torch.device) -> dict:
        key = str(device)
        t = self._dev_cache.get(key)
        if t is not None:
            return t
        c = self.config
        n_fft, n = c.ms_frame_padded_length, c.ms_frame_padded_length // 2
        w_low = _window(c, c.ms_window_exponent_low)
        hz = torch.linspace(0, c.sample_rate / 2, c.ms_num_stft_bins)
        density = 1127.0 / (700.0 + hz)                                         # get_mel_density
        norm_low = 1.0 / w_low.pow(2.0).sum().sqrt()                            # torchaudio normalized="window"
        if c.ms_window_exponent_high is not None:
            w_high = _window(c, c.ms_window_exponent_high)
            norm_high = 1.0 / w_high.pow(2.0).sum().sqrt()
            bw = (density / density.amax()) ** 2                                # :181-184
            coef1 = norm_low * bw / density
            coef2 = norm_high * (1 - bw) / density
        else:
            w_high, coef2 = None, None
            coef1 = norm_low / density
        fb = self.ms_freq_scale.get_filters()
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
        t = dict(window=w_low.float().to(device).contiguous(),
                 window2=None if w_high is None else w_high.float().to(device).contiguous(),
                 coef1=coef1.float().to(device).contiguous(),
                 coef2=None if coef2 is None else coef2.float().to(device).contiguous(),
                 tw=to(np.stack([tw.real, tw.imag], -1), torch.float32),
                 tw_half=to(np.stack([tw_half.real, tw_half.imag], -1), torch.float32),
                 fb=dict(start=to(starts, torch.int32), count=to(counts, torch.int32), offset=to(offsets, torch.int32),
                         weight=to(weights, torch.float32)))
        self._dev_cache[key] = t
        return t

    # ---- ms_mdct_dual.py:230-257 ----
    @torch.no_grad()
    def raw_to_mel_spec(self, raw_samples: torch.Tensor, use_slicing: bool = False) -> torch.Tensor:
        c = self.config
        if c.ms_freq_min > 0 and (self.ms_lowest_filter_freq - c.ms_freq_min) > 0:
            raise NotImplementedError("ms_freq_min > 0 (full-length FFT high-pass, :190-207) is not implemented")
        L.require_cuda(raw_samples)
        t = self._tables(raw_samples.device)
        B, C, n = raw_samples.shape
        raw = raw_samples.detach().float().contiguous().view(B * C, n)
        if c.raw_to_mel_spec_scale == 0:
            raise ValueError("raw_to_mel_spec_scale must be non-zero")
        # kernel computes (v ** e - mean) * scale == v ** e * scale + offset
        out = ops.stft_mel(raw, t["window"], t["tw"], t["tw_half"], c.ms_frame_padded_length, c.ms_frame_hop_length,
                           t["fb"], c.ms_abs_exponent, -c.raw_to_mel_spec_offset / c.raw_to_mel_spec_scale,
                           c.raw_to_mel_spec_scale, window2=t["window2"], coef1=t["coef1"], coef2=t["coef2"])
        return out.view(B, C, c.ms_num_frequencies, out.shape[-1])

    # ---- MDCT side (ms_mdct_dual.py:259-318; utils/mclt.py:87-130) ----
    def _mdct_tables(self, device: torch.device) -> dict:
        """Host-built (fp64) transform matrices.  forward  Mt [2N][2N]: rows 0..N-1 = Re, N..2N-1 = Im of
        window * pre-shift * DFT * post-shift * 2 sqrt(N) / (2N) * raw_to_mdct_scale / mel_density[k];
        inverse  St [2N][2N]: rows 0..N-1 multiply Re(x), N..2N-1 multiply Im(x), columns = the 2N samples of a frame,
        with mel_density / raw_to_mdct_scale, 2 sqrt(N) and mdct_to_raw_scale folded in;
        P [bins][filters]: min-norm inverse mel filterbank (lstsq(gels), frequency_scale.py:136) times mel_spec_to_mdct_psd_scale."""
        key = "mdct:" + str(device)
        t = self._dev_cache.get(key)
        if t is not None:
            return t
        c = self.config
        if c.mdct_window_func != "kaiser_bessel_derived":
            raise NotImplementedError(f"mdct_window_func {c.mdct_window_func!r} is not implemented (default: kaiser_bessel_derived)")
        N = c.mdct_window_len // 2
        kais = torch.kaiser_window(N + 1, beta=4.0, periodic=False).double()          # utils/mclt.py:44-62
        cs = torch.cumsum(kais[:-1] ** 2, dim=0)
        half = torch.sqrt(cs / cs[-1]).numpy()
        w = np.concatenate([half, half[::-1]])
        n = np.arange(2 * N, dtype=np.float64)[:, None]
        k = np.arange(N, dtype=np.float64)[None, :]
        hz = (np.arange(N) + 0.5) * c.sample_rate / c.mdct_window_len
        dens = 1127.0 / (700.0 + hz)
        fwd = (w[:, None] * np.exp(-1j * np.pi * n / (2 * N)) * np.exp(-2j * np.pi * k * n / (2 * N)) / (2 * N)
               * np.exp(-1j * np.pi * (N + 1) * (k + 0.5) / (2 * N)) * 2 * np.sqrt(N) * (c.raw_to_mdct_scale / dens)[None, :])
        phi = np.pi * (N + 1) * (k + 0.5) / (2 * N) + 2 * np.pi * k * n / (2 * N) + np.pi * n / (2 * N)
        gsc = (dens / c.raw_to_mdct_scale)[None, :] * 2 * np.sqrt(N) * c.mdct_to_raw_scale
        s_re = w[:, None] / (2 * N) * np.cos(phi) * gsc
        s_im = -w[:, None] / (2 * N) * np.sin(phi) * gsc
        if c.mdct_psd_num_bins == c.ms_num_stft_bins - 1:
            fb = self.ms_freq_scale.get_filters()
        else:
            fb = FrequencyScale(c.ms_freq_min, c.sample_rate / 2, c.sample_rate, c.mdct_psd_num_bins, c.ms_num_frequencies,
                                "slaney").get_filters()
        a64 = fb.double().numpy().T                                                     # (filters, bins)
        gram = a64 @ a64.T
        if np.linalg.cond(gram) > 1e12:
            raise ValueError("mel filterbank is rank deficient: min-norm inverse undefined")
        pinv = a64.T @ np.linalg.inv(gram) * c.mel_spec_to_mdct_psd_scale
        to = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32).to(device)
        t = dict(N=N, fwd=to(np.concatenate([fwd.real.T, fwd.imag.T], 0)), inv=to(np.concatenate([s_re.T, s_im.T], 0)),
                 pinv=to(pinv))
        self._dev_cache[key] = t
        return t

    def _get_mdct_raw_crop_width(self, raw_length: Optional[int] = None) -> int:
        c = self.config
        raw_length = raw_length or c.default_raw_length
        return raw_length // c.mdct_window_len // c.ms_width_alignment * c.ms_width_alignment * c.mdct_window_len + c.mdct_window_len

    def get_mdct_shape(self, bsz: int = 1, raw_length: Optional[int] = None) -> tuple:
        c = self.config
        n_bins = c.mdct_window_len // 2
        return (bsz, c.num_raw_channels * (2 if c.mdct_dual_channel else 1), n_bins,
                self.get_raw_crop_width(raw_length=raw_length) // n_bins + 1)

    def _check_no_high_pass(self) -> None:
        c = self.config
        if c.ms_freq_min > 0 and (self.ms_lowest_filter_freq - c.ms_freq_min) > 0:
            raise NotImplementedError("ms_freq_min > 0 (full-length FFT high-pass, :190-207) is not implemented")

    def _mclt_rows(self, raw_samples: torch.Tensor) -> torch.Tensor:
        """(B, C, L) -> [B*C][2N][T]: real rows then imaginary rows of the scaled, density-weighted MCLT."""
        self._check_no_high_pass()
        L.require_cuda(raw_samples)
        t = self._mdct_tables(raw_samples.device)
        N = t["N"]
        B, C, n = raw_samples.shape
        raw = raw_samples.detach().float().contiguous().view(B * C, n)
        rem = n % N
        padded = n + 2 * N + (N - rem if rem else 0)
        T = (padded - 2 * N) // N + 1
        # y[s][r][t] = sum_k fwd[r][k] * frame[s][t][k]; the frames (hop N, reflect padding N) are gathered from `raw` by the
        # GEMM's B-operand loads, never materialised
        y = torch.empty((B * C, 2 * N, T), device=raw.device, dtype=torch.float32)
        return ops.gemm_f32(t["fwd"], (2 * N, 1, 0), raw, (0, 0, n), y, (T, 1, 2 * N * T), 2 * N, T, 2 * N, B * C,
                            gather=(N, N, n))

    @torch.no_grad()
    def raw_to_mdct(self, raw_samples: torch.Tensor, random_phase_augmentation: bool = False) -> torch.Tensor:
        """ms_mdct_dual.py:283-298 -> (B, C, N, T), or (B, 2C, N, T) = [real | imag] when mdct_dual_channel."""
        if random_phase_augmentation:
            raise NotImplementedError("random_phase_augmentation (training-time augmentation) is not implemented")
        B, C, _ = raw_samples.shape
        y = self._mclt_rows(raw_samples)
        N, T = y.shape[1] // 2, y.shape[2]
        if self.config.mdct_dual_channel:
            return y.view(B, C, 2, N, T).permute(0, 2, 1, 3, 4).reshape(B, 2 * C, N, T)
        return y[:, :N].reshape(B, C, N, T)

    @torch.no_grad()
    def raw_to_mdct_psd(self, raw_samples: torch.Tensor) -> torch.Tensor:
        """ms_mdct_dual.py:300-306."""
        B, C, _ = raw_samples.shape
        y = self._mclt_rows(raw_samples)
        out = ops.complex_abs(y.contiguous(), 1.0 / math.sqrt(2.0))
        return out.view(B, C, out.shape[1], out.shape[2])

    @torch.no_grad()
    def mdct_to_raw(self, mdct: torch.Tensor) -> torch.Tensor:
        """ms_mdct_dual.py:308-318 -> (B, C, (T-1)*N)."""
        L.require_cuda(mdct)
        t = self._mdct_tables(mdct.device)
        N = t["N"]
        x = mdct.detach().float()
        B, Cx, Nb, T = x.shape
        if Nb != N:
            raise ValueError(f"expected {N} MDCT bins, got {Nb}")
        if self.config.mdct_dual_channel:
            C = Cx // 2
            x = torch.cat((x[:, :C], x[:, C:]), dim=2)                       # [B][C][2N (re | im)][T]
            mat = t["inv"]
        else:
            C = Cx
            mat = t["inv"][:N]
        x = x.reshape(B * C, -1, T).contiguous()                              # [S][K][T], K = N or 2N (re | im)
        K = x.shape[1]
        y = torch.empty((B * C, T, 2 * N), device=x.device, dtype=torch.float32)
        ops.gemm_f32(x, (1, T, K * T), mat, (2 * N, 1, 0), y, (2 * N, 1, T * 2 * N), T, 2 * N, K, B * C)   # A = x^T as strides
        return ops.mdct_ola(y).view(B, C, -1)

    @torch.no_grad()
    def mel_spec_to_mdct_psd(self, mel_spec: torch.Tensor) -> torch.Tensor:
        """ms_mdct_dual.py:259-270 -> (B, C, mdct_psd_num_bins, T)."""
        L.require_cuda(mel_spec)
        c = self.config
        t = self._mdct_tables(mel_spec.device)
        lin = ops.mel_linearize(mel_spec.detach().float().contiguous(), c.raw_to_mel_spec_offset, 1.0 / c.ms_abs_exponent)
        Bm, Cm, F, T = lin.shape
        nb = t["pinv"].shape[0]
        psd = torch.empty((Bm, Cm, nb, T), device=lin.device, dtype=torch.float32)
        ops.gemm_f32(t["pinv"], (F, 1, 0), lin, (T, 1, F * T), psd, (T, 1, nb * T), nb, T, F, Bm * Cm)
        if c.mdct_psd_num_bins == c.ms_num_stft_bins - 1:
            psd = psd[:, :, :-1, :]
        if c.mel_spec_to_mdct_psd_offset != 0:
            psd = psd + c.mel_spec_to_mdct_psd_offset
        return psd
