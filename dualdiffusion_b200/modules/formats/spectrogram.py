"""B200-native drop-in for the reference's mel-spectrogram format with FGLA phase reconstruction
(`modules.formats.old.spectrogram.SpectrogramFormat`, src/modules/formats/old/spectrogram.py:33-238; the module
index of the shipped default model names it as `modules.formats.spectrogram`).

Same config fields/defaults and public methods: `raw_to_sample(raw, unscaled_spectrogram=False)`,
`sample_to_raw(samples, n_fgla_iters=None, quiet=False, unscaled_spectrogram=False)`, `get_sample_shape(bsz, length)`,
`sample_raw_crop_width(length)`.  The transforms run as fused CUDA kernels (dd_stft_mel, dd_fgla_istft,
dd_fgla_stft_update, dd_ola_finalize): shared-memory mixed-radix FFTs with the window, |.|, sparse mel
filterbank, overlap-add and Griffin-Lim momentum update fused around them.  The one dense contraction, the
min-norm inverse mel (`torch.linalg.lstsq(gels)`, frequency_scale.py:136), is a plain library GEMM with the
fp64-precomputed pseudo-inverse P = A^T (A A^T)^-1 (identical solution for the full-row-rank filterbank).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Literal, Optional

import numpy as np
import torch

from ... import _lib as L
from ... import ops
from .format import DualDiffusionFormat, DualDiffusionFormatConfig


@dataclass
class SpectrogramFormatConfig(DualDiffusionFormatConfig):
    """old/spectrogram.py:33-74 (fields used by the hot path, same defaults)."""
    sample_raw_channels: int = 2
    sample_raw_length: int = 1408768
    raw_to_sample_scale: float = 2.247
    sample_to_raw_scale: float = 0.445
    sample_mean: float = 1.295
    abs_exponent: float = 0.25
    sample_rate: int = 32000
    step_size_ms: int = 8
    window_duration_ms: int = 200
    padded_duration_ms: int = 200
    window_exponent: float = 32
    window_periodic: bool = True
    freq_scale_type: Literal["mel", "log"] = "mel"
    num_frequencies: int = 256
    min_frequency: int = 20
    max_frequency: int = 16000
    freq_scale_norm: Optional[str] = None
    num_fgla_iters: int = 200
    fgla_momentum: float = 0.99
    stereo_coherence: float = 0.67

    @property
    def stereo(self) -> bool:
        return self.sample_raw_channels == 2

    @property
    def num_stft_bins(self) -> int:
        return self.padded_length // 2 + 1

    @property
    def padded_length(self) -> int:
        return int(self.padded_duration_ms / 1000.0 * self.sample_rate)

    @property
    def win_length(self) -> int:
        return int(self.window_duration_ms / 1000.0 * self.sample_rate)

    @property
    def hop_length(self) -> int:
        return int(self.step_size_ms / 1000.0 * self.sample_rate)


def mel_filterbank(cfg: SpectrogramFormatConfig) -> torch.Tensor:
    """frequency_scale.py:30-34,45-58,144-169 -> (n_stft_bins, n_filters) fp32 (host-side setup)."""
    if cfg.freq_scale_type != "mel":
        raise NotImplementedError("only the mel frequency scale is implemented")
    lo = 2595.0 * np.log10(1.0 + cfg.min_frequency / 700.0)
    hi = 2595.0 * np.log10(1.0 + cfg.max_frequency / 700.0)
    f_pts = 700.0 * (10.0 ** (torch.linspace(lo, hi, cfg.num_frequencies + 2) / 2595.0) - 1.0)
    all_freqs = torch.linspace(0, cfg.sample_rate / 2, cfg.num_stft_bins)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    fb = torch.max(torch.zeros(1), torch.min((-1.0 * slopes[:, :-2]) / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))
    if cfg.freq_scale_norm == "slaney":
        fb = fb * (2.0 / (f_pts[2:cfg.num_frequencies + 2] - f_pts[:cfg.num_frequencies])).unsqueeze(0)
    return fb


class SpectrogramFormat(DualDiffusionFormat):

    # resolved by from_pretrained (module.py:72); explicit because this file's annotations are strings
    config_class = SpectrogramFormatConfig

    def __init__(self, config: SpectrogramFormatConfig) -> None:
        super().__init__()
        self.config = config
        if config.win_length != config.padded_length:
            raise NotImplementedError("window shorter than n_fft is not implemented (reference default: equal)")
        self._dev_cache = {}

    # ---- shapes (old/spectrogram.py:160-216) ----
    def _spectrogram_len(self, audio_len: int) -> int:
        c = self.config
        return 1 + (audio_len + c.padded_length - c.win_length) // c.hop_length

    def _audio_len(self, spectrogram_len: int) -> int:
        c = self.config
        return (spectrogram_len - 1) * c.hop_length + c.win_length - c.padded_length

    def sample_raw_crop_width(self, length: Optional[int] = None) -> int:
        length = length or self.config.sample_raw_length
        return self._audio_len(self._spectrogram_len(length) // 128 * 128)

    def get_sample_shape(self, bsz: int = 1, length: Optional[int] = None) -> tuple:
        crop = self.sample_raw_crop_width(length)
        return (bsz, self.config.sample_raw_channels, self.config.num_frequencies, self._spectrogram_len(crop))

    # ---- device-side constant tables ----
    def _tables(self, device: torch.device) -> dict:
        key = str(device)
        t = self._dev_cache.get(key)
        if t is not None:
            return t
        c = self.config
        n_fft, n = c.padded_length, c.padded_length // 2
        window = torch.hann_window(c.win_length, periodic=c.window_periodic) ** c.window_exponent    # :99-105
        k = np.arange(n, dtype=np.float64)
        tw = np.exp(-2j * np.pi * k / n)
        kh = np.arange(n + 1, dtype=np.float64)
        tw_half = np.exp(-2j * np.pi * kh / n_fft)
        fb = mel_filterbank(c)                                               # (bins, filters)
        starts, counts, offsets, weights = [], [], [], []
        for f in range(fb.shape[1]):
            nz = torch.nonzero(fb[:, f]).flatten()
            if nz.numel() == 0:
                starts.append(0); counts.append(0); offsets.append(len(weights)); continue
            lo, hi = int(nz[0]), int(nz[-1]) + 1
            starts.append(lo); counts.append(hi - lo); offsets.append(len(weights))
            weights.extend(fb[lo:hi, f].tolist())
        a64 = fb.double().numpy().T                                          # (filters, bins)
        gram = a64 @ a64.T
        if np.linalg.cond(gram) > 1e12:
            raise ValueError("mel filterbank is rank deficient (an all-zero filter?): min-norm inverse undefined")
        pinv = (a64.T @ np.linalg.inv(gram)).astype(np.float32)             # (bins, filters), = lstsq(gels) min-norm
        to = lambda x, dt: torch.as_tensor(x, dtype=dt).to(device).contiguous()
        t = dict(window=window.to(device).contiguous(),
                 tw=to(np.stack([tw.real, tw.imag], -1), torch.float32),
                 tw_half=to(np.stack([tw_half.real, tw_half.imag], -1), torch.float32),
                 fb=dict(start=to(starts, torch.int32), count=to(counts, torch.int32), offset=to(offsets, torch.int32),
                         weight=to(weights, torch.float32)),
                 pinv_t=to(pinv.T.copy(), torch.float32),                    # (filters, bins)
                 env={})
        self._dev_cache[key] = t
        return t

    def _envelope(self, t: dict, n_frames: int, device: torch.device) -> torch.Tensor:
        """Sum of squared windows at every padded-domain sample (the divisor torch.istft applies)."""
        env = t["env"].get(n_frames)
        if env is None:
            c = self.config
            w2 = (t["window"].double().cpu().numpy()) ** 2
            e = np.zeros(c.padded_length + c.hop_length * (n_frames - 1), dtype=np.float64)
            for i in range(n_frames):
                e[i * c.hop_length: i * c.hop_length + c.padded_length] += w2
            env = torch.as_tensor(e, dtype=torch.float32).to(device)
            t["env"][n_frames] = env
        return env

    # ---- encode (old/spectrogram.py:176-179, 218-226) ----
    @torch.no_grad()
    def raw_to_sample(self, raw_samples: torch.Tensor, unscaled_spectrogram: bool = False) -> torch.Tensor:
        if unscaled_spectrogram:
            raise NotImplementedError("unscaled_spectrogram=True is not on the hot path")
        L.require_cuda(raw_samples)
        c = self.config
        t = self._tables(raw_samples.device)
        B, C, n = raw_samples.shape
        raw = raw_samples.detach().float().contiguous().view(B * C, n)
        out = ops.stft_mel(raw, t["window"], t["tw"], t["tw_half"], c.padded_length, c.hop_length, t["fb"],
                           c.abs_exponent, c.sample_mean, c.raw_to_sample_scale)
        return out.view(B, C, c.num_frequencies, out.shape[-1])

    # ---- decode (old/spectrogram.py:181-185, 229-238; old/phase_recovery.py:40-129) ----
    @torch.no_grad()
    def sample_to_raw(self, samples: torch.Tensor, n_fgla_iters: Optional[int] = None, quiet: bool = False,
                      unscaled_spectrogram: bool = False) -> torch.Tensor:
        if unscaled_spectrogram:
            raise NotImplementedError("unscaled_spectrogram=True is not on the hot path")
        L.require_cuda(samples)
        c = self.config
        n_iter = n_fgla_iters or c.num_fgla_iters
        if n_iter <= 0:
            raise ValueError("n_fgla_iters must be positive")
        t = self._tables(samples.device)
        B, C, F, T = samples.shape
        S = B * C
        n_fft, hop = c.padded_length, c.hop_length
        # mag[s][t][bin] = relu(sum_f mel_lin[s][f][t] * pinv_t[f][bin]) with mel_lin = (x/scale + mean).clip(0) ** (1/abs_exponent)
        # (old/spectrogram.py:229-233, frequency_scale.py:130-142): one launch, the linearisation on the A-operand loads and
        # the relu on the stores
        x32 = samples.detach().float().contiguous().view(S, F, T)
        nbins = t["pinv_t"].shape[1]
        mag = torch.empty((S, T, nbins), device=samples.device, dtype=torch.float32)
        ops.gemm_f32(x32, (1, T, F * T), t["pinv_t"], (nbins, 1, 0), mag, (nbins, 1, T * nbins), T, nbins, F, S,
                     a_transform=(1.0 / c.raw_to_sample_scale, c.sample_mean, 1.0 / c.abs_exponent), relu=True)
        env = self._envelope(t, T, samples.device)
        ola = torch.empty((S, n_fft + hop * (T - 1)), device=samples.device, dtype=torch.float32)
        state = torch.empty((S, T, c.num_stft_bins, 2), device=samples.device, dtype=torch.float32)
        momentum = c.fgla_momentum / (1 + c.fgla_momentum)                               # phase_recovery.py:58
        stereo = c.stereo
        args = (t["window"], t["tw"], t["tw_half"], n_fft, hop)
        for i in range(n_iter):
            interp_t = i / n_iter - c.stereo_coherence                                   # :84
            ops.fgla_istft(None if i == 0 else state, mag, stereo, interp_t, *args, ola)
            ops.fgla_stft_update(ola, env, state, momentum, i == 0, *args)
        ops.fgla_istft(state, mag, False, 0.0, *args, ola)                               # :121-124
        wave = ops.ola_finalize(ola, env, n_fft, hop * (T - 1))
        return wave.view(B, C, -1)
