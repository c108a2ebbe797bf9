"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's mel-STFT encode and FGLA (fast Griffin-Lim)
phase-reconstruction decode: `SpectrogramFormat.raw_to_sample / sample_to_raw`
(src/modules/formats/old/spectrogram.py:176-238), `FrequencyScale` (src/modules/formats/frequency_scale.py:30-58,
127-169) and `griffinlim` (src/modules/formats/old/phase_recovery.py:40-129).

The reference delegates the transforms to torch / torchaudio (not vendored): torchaudio.transforms.Spectrogram ==
torch.stft(center=True, pad_mode="reflect", window, onesided, normalized=False); torch.istft; torch.linalg.lstsq.
This restatement calls torch.stft / torch.istft for those and writes everything else out explicitly, including the
in-place aliasing of the momentum update (`tprev` holds the momentum-ADJUSTED spectrum).
Pinned by tests/golden/format_small.pt (reference output, see tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

Tensor = torch.Tensor


@dataclass
class SpectrogramSpec:
    """old/spectrogram.py:33-74 defaults."""
    raw_to_sample_scale: float = 2.247
    sample_mean: float = 1.295
    abs_exponent: float = 0.25
    sample_rate: int = 32000
    step_size_ms: int = 8
    window_duration_ms: int = 200
    padded_duration_ms: int = 200
    window_exponent: float = 32
    window_periodic: bool = True
    num_frequencies: int = 256
    min_frequency: int = 20
    max_frequency: int = 16000
    num_fgla_iters: int = 200
    fgla_momentum: float = 0.99
    stereo_coherence: float = 0.67

    @property
    def n_fft(self) -> int:
        return int(self.padded_duration_ms / 1000.0 * self.sample_rate)

    @property
    def win_length(self) -> int:
        return int(self.window_duration_ms / 1000.0 * self.sample_rate)

    @property
    def hop_length(self) -> int:
        return int(self.step_size_ms / 1000.0 * self.sample_rate)

    @property
    def num_stft_bins(self) -> int:
        return self.n_fft // 2 + 1


def window(spec: SpectrogramSpec) -> Tensor:
    """old/spectrogram.py:99-105 — hann(periodic) ** exponent, fp32."""
    return torch.hann_window(spec.win_length, periodic=spec.window_periodic) ** spec.window_exponent


def mel_filterbank(spec: SpectrogramSpec) -> Tensor:
    """frequency_scale.py:30-34,45-58,144-169 — HTK mel points, triangular filters, no area norm -> (bins, filters)."""
    lo = 2595.0 * math.log10(1.0 + spec.min_frequency / 700.0)
    hi = 2595.0 * math.log10(1.0 + spec.max_frequency / 700.0)
    f_pts = 700.0 * (10.0 ** (torch.linspace(lo, hi, spec.num_frequencies + 2) / 2595.0) - 1.0)
    all_freqs = torch.linspace(0, spec.sample_rate / 2, spec.num_stft_bins)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def stft(x: Tensor, spec: SpectrogramSpec) -> Tensor:
    shape = x.shape
    y = torch.stft(x.reshape(-1, shape[-1]), n_fft=spec.n_fft, hop_length=spec.hop_length, win_length=spec.win_length,
                   window=window(spec), center=True, pad_mode="reflect", normalized=False, onesided=True,
                   return_complex=True)
    return y.reshape(shape[:-1] + y.shape[-2:])


def raw_to_sample(raw: Tensor, spec: SpectrogramSpec) -> Tensor:
    """old/spectrogram.py:176-179, 218-226: |STFT| -> mel -> ** 0.25 -> (x - mean) * scale."""
    mag = stft(raw.float(), spec).abs()
    mel = torch.matmul(mag.transpose(-1, -2), mel_filterbank(spec)).transpose(-1, -2)      # frequency_scale.py:127-128
    return (mel ** spec.abs_exponent - spec.sample_mean) * spec.raw_to_sample_scale


def unscale(mel_lin: Tensor, spec: SpectrogramSpec) -> Tensor:
    """frequency_scale.py:130-142 — min-norm least-squares inverse of the filterbank (lstsq gels) + relu."""
    shape = mel_lin.shape
    s = mel_lin.reshape(-1, shape[-2], shape[-1])
    sol = torch.linalg.lstsq(mel_filterbank(spec).transpose(-1, -2)[None], s, driver="gels").solution
    return torch.relu(sol).view(shape[:-2] + (spec.num_stft_bins, shape[-1]))


def griffinlim_step(tprev: Tensor, specgram: Tensor, merged: Optional[Tensor], spec: SpectrogramSpec, i: int,
                    n_iter: int) -> Tensor:
    """One iteration of the loop at phase_recovery.py:78-119, as a map on the momentum-adjusted spectrum
    T_i -> T_{i+1} (`tprev`; None before the first iteration, where angles == 1):
        A = T/(|T|+1e-16);  rebuilt = STFT(ISTFT(A * M_i));  T' = rebuilt - m/(1+m) * T.
    `angles.sub_(tprev, alpha=momentum)` (:113) acts in place on `rebuilt`, so the `tprev = rebuilt` of :117 stores
    the momentum-ADJUSTED spectrum."""
    momentum = spec.fgla_momentum / (1 + spec.fgla_momentum)                                # :58
    kw = dict(n_fft=spec.n_fft, hop_length=spec.hop_length, win_length=spec.win_length, window=window(spec))
    if merged is not None:
        t = i / n_iter - spec.stereo_coherence                                              # :84-88
        interp = torch.lerp(merged, specgram, t) if t > 0 else merged
    else:
        interp = specgram
    if tprev is None:
        angles = torch.full((1,) + tuple(specgram.shape[1:]), 1, dtype=torch.cfloat)        # :73
    else:
        angles = tprev / (tprev.abs() + 1e-16)                                              # :115
    inverse = torch.istft(angles * interp, length=None, **kw)                               # :92-95
    rebuilt = torch.stft(inverse, center=True, pad_mode="reflect", normalized=False, onesided=True,
                         return_complex=True, **kw)                                         # :97-108
    return rebuilt if tprev is None else rebuilt - momentum * tprev                         # :110-113 (tprev_0 == 0)


def griffinlim(specgram: Tensor, spec: SpectrogramSpec, n_iter: int, stereo: bool = True) -> Tensor:
    """phase_recovery.py:40-129 with rand_init=False, length=None."""
    shape = specgram.shape
    specgram = specgram.reshape([-1] + list(shape[-2:]))
    merged = ((specgram[0::2] + specgram[1::2]) / 2).repeat_interleave(2, dim=0) if stereo else None   # :63-64
    tprev = None
    for i in range(n_iter):
        tprev = griffinlim_step(tprev, specgram, merged, spec, i, n_iter)
    angles = tprev / (tprev.abs() + 1e-16)
    kw = dict(n_fft=spec.n_fft, hop_length=spec.hop_length, win_length=spec.win_length, window=window(spec))
    wave = torch.istft(angles * specgram, length=None, **kw)                                # :121-124
    return wave.reshape(shape[:-2] + wave.shape[-1:])


def sample_to_raw(samples: Tensor, spec: SpectrogramSpec, n_fgla_iters: Optional[int] = None) -> Tensor:
    """old/spectrogram.py:229-238, :181-185."""
    s = (samples / spec.raw_to_sample_scale + spec.sample_mean).clip(min=0)
    amplitudes = unscale(s ** (1 / spec.abs_exponent), spec)
    return griffinlim(amplitudes, spec, n_fgla_iters or spec.num_fgla_iters, stereo=True)


# --------------------------------------------------------------------------------------
# live format: MS_MDCT_DualFormat.raw_to_mel_spec (src/modules/formats/ms_mdct_dual.py:230-257)
# --------------------------------------------------------------------------------------
@dataclass
class MSDualSpec:
    """ms_mdct_dual.py:36-66 defaults (mel-spectrogram side)."""
    sample_rate: int = 32000
    raw_to_mel_spec_scale: float = 50
    raw_to_mel_spec_offset: float = 0
    ms_abs_exponent: float = 1
    ms_freq_min: float = 0
    ms_num_frequencies: int = 256
    ms_step_size_ms: int = 8
    ms_window_duration_ms: int = 128
    ms_padded_duration_ms: int = 128
    ms_window_exponent_low: float = 17
    ms_window_exponent_high: Optional[float] = 58

    @property
    def n_fft(self) -> int:
        return int(self.ms_padded_duration_ms / 1000.0 * self.sample_rate)

    @property
    def hop_length(self) -> int:
        return int(self.ms_step_size_ms / 1000.0 * self.sample_rate)

    @property
    def num_stft_bins(self) -> int:
        return self.n_fft // 2 + 1


def blackman_harris(n: int) -> Tensor:
    """utils/mclt.py:69-71."""
    x = torch.arange(n) / n * 2 * torch.pi
    return 0.35875 - 0.48829 * torch.cos(x) + 0.14128 * torch.cos(2 * x) - 0.01168 * torch.cos(3 * x)


def ms_mel_filterbank(spec: MSDualSpec) -> Tensor:
    """FrequencyScale(mel, freq_min, sr/2, slaney norm, triangular) — frequency_scale.py:151-169."""
    lo = 2595.0 * math.log10(1.0 + spec.ms_freq_min / 700.0)
    hi = 2595.0 * math.log10(1.0 + (spec.sample_rate / 2) / 700.0)
    f_pts = 700.0 * (10.0 ** (torch.linspace(lo, hi, spec.ms_num_frequencies + 2) / 2595.0) - 1.0)
    all_freqs = torch.linspace(0, spec.sample_rate / 2, spec.num_stft_bins)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    fb = torch.max(torch.zeros(1), torch.min((-1.0 * slopes[:, :-2]) / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))
    enorm = 2.0 / (f_pts[2:spec.ms_num_frequencies + 2] - f_pts[:spec.ms_num_frequencies])
    return fb * enorm.unsqueeze(0)


def raw_to_mel_spec(raw: Tensor, spec: MSDualSpec) -> Tensor:
    """ms_mdct_dual.py:230-257 with ms_freq_min == 0 (no high-pass): two magnitude STFTs (blackman-harris ** 17 / ** 58,
    torchaudio `normalized="window"`), per-bin blend by (mel_density/max)^2, / mel_density, mel filterbank, affine."""
    def mag_stft(exponent: float) -> Tensor:
        w = blackman_harris(spec.n_fft) ** exponent
        shape = raw.shape
        y = torch.stft(raw.float().reshape(-1, shape[-1]), n_fft=spec.n_fft, hop_length=spec.hop_length,
                       win_length=spec.n_fft, window=w, center=True, pad_mode="reflect", normalized=False,
                       onesided=True, return_complex=True)
        y = y / w.pow(2.0).sum().sqrt()                                    # normalized="window"
        return y.abs().reshape(shape[:-1] + y.shape[-2:])
    hz = torch.linspace(0, spec.sample_rate / 2, spec.num_stft_bins)
    density = (1127.0 / (700.0 + hz)).view(1, 1, -1, 1)                    # get_mel_density, frequency_scale.py:36-37
    low = mag_stft(spec.ms_window_exponent_low)
    if spec.ms_window_exponent_high is not None:
        bw = ((density / density.amax()) ** 2)
        blended = low * bw + mag_stft(spec.ms_window_exponent_high) * (1 - bw)   # :252
    else:
        blended = low
    mel = torch.matmul((blended / density).transpose(-1, -2), ms_mel_filterbank(spec)).transpose(-1, -2)
    return mel ** spec.ms_abs_exponent * spec.raw_to_mel_spec_scale + spec.raw_to_mel_spec_offset   # :256-257


# --------------------------------------------------------------------------------------
# live format, MDCT side: utils/mclt.py:87-130 and
# MS_MDCT_DualFormat.{mel_spec_to_mdct_psd, raw_to_mdct, raw_to_mdct_psd, mdct_to_raw} (ms_mdct_dual.py:259-325)
# --------------------------------------------------------------------------------------
@dataclass
class MDCTSpec:
    """ms_mdct_dual.py:36-66 defaults (MDCT side)."""
    sample_rate: int = 32000
    mdct_window_len: int = 512
    mdct_psd_num_bins: int = 2048
    mdct_dual_channel: bool = False
    raw_to_mdct_scale: float = 12.1
    mdct_to_raw_scale: float = 2
    mel_spec_to_mdct_psd_scale: float = 0.18
    mel_spec_to_mdct_psd_offset: float = 0

    @property
    def num_bins(self) -> int:
        return self.mdct_window_len // 2


def kbd_window(n: int, beta: float = 4.0) -> Tensor:
    """WindowFunction.kaiser_bessel_derived (utils/mclt.py:44-62)."""
    k = torch.kaiser_window(n // 2 + 1, beta=beta, periodic=False)
    c = torch.cumsum(k[:-1] ** 2, dim=0)
    half = torch.sqrt(c / c[-1])
    return torch.cat((half, half.flip(0)), dim=0)


def mclt(x: Tensor, block_width: int) -> Tensor:
    """utils/mclt.py:87-109 with the kaiser-bessel-derived window, exponent 1: (..., L) -> (..., frames, N) complex."""
    pad_l = pad_r = block_width // 2
    rem = x.shape[-1] % (block_width // 2)
    if rem > 0:
        pad_r += block_width // 2 - rem
    x = torch.nn.functional.pad(x, (pad_l, pad_r), mode="reflect").unfold(-1, block_width, block_width // 2)
    n_bins = block_width // 2
    n = torch.arange(2 * n_bins)
    k = torch.arange(0.5, n_bins + 0.5)
    pre = torch.exp(-1j * torch.pi / 2 / n_bins * n)
    post = torch.exp(-1j * torch.pi / 2 / n_bins * (n_bins + 1) * k)
    return torch.fft.fft(x * pre * kbd_window(2 * n_bins), norm="forward")[..., :n_bins] * post * (2 * n_bins ** 0.5)


def imclt(x: Tensor) -> Tensor:
    """utils/mclt.py:111-130: (..., frames, N) complex -> (..., (frames-1)*N) complex."""
    n_bins = x.shape[-1]
    n = torch.arange(2 * n_bins)
    k = torch.arange(0.5, n_bins + 0.5)
    pre = torch.exp(-1j * torch.pi / 2 / n_bins * n)
    post = torch.exp(-1j * torch.pi / 2 / n_bins * (n_bins + 1) * k)
    y = (torch.fft.ifft(x / post, norm="backward", n=2 * n_bins) / pre) * kbd_window(2 * n_bins)
    total = (y.shape[-2] + 1) * y.shape[-1] // 2
    out = torch.zeros(y.shape[:-2] + (total,), dtype=y.dtype)
    even = y[..., ::2, :].reshape(*y[..., ::2, :].shape[:-2], -1)
    odd = y[..., 1::2, :].reshape(*y[..., 1::2, :].shape[:-2], -1)
    out[..., :even.shape[-1]] = even
    out[..., n_bins:odd.shape[-1] + n_bins] += odd
    return out[..., n_bins:-n_bins] * (2 * n_bins ** 0.5)


def mdct_mel_density(spec: MDCTSpec) -> Tensor:
    hz = (torch.arange(spec.num_bins) + 0.5) * spec.sample_rate / spec.mdct_window_len      # ms_mdct_dual.py:177-179
    return (1127.0 / (700.0 + hz)).view(1, 1, -1, 1)


def raw_to_mdct(raw: Tensor, spec: MDCTSpec) -> Tensor:
    """ms_mdct_dual.py:283-298 (ms_freq_min == 0: no high-pass; no phase augmentation)."""
    m = mclt(raw.float(), spec.mdct_window_len).permute(0, 1, 3, 2)
    if spec.mdct_dual_channel:
        m = torch.cat((m.real, m.imag), dim=1)
    else:
        m = m.real
    return m / mdct_mel_density(spec) * spec.raw_to_mdct_scale


def raw_to_mdct_psd(raw: Tensor, spec: MDCTSpec) -> Tensor:
    """ms_mdct_dual.py:300-306."""
    m = mclt(raw.float(), spec.mdct_window_len).permute(0, 1, 3, 2)
    return m.abs() / mdct_mel_density(spec) * spec.raw_to_mdct_scale / 2 ** 0.5


def mdct_to_raw(mdct: Tensor, spec: MDCTSpec) -> Tensor:
    """ms_mdct_dual.py:308-318."""
    m = mdct * mdct_mel_density(spec) / spec.raw_to_mdct_scale
    if spec.mdct_dual_channel:
        m = torch.complex(*m.chunk(2, dim=1))
    return imclt(m.permute(0, 1, 3, 2).to(torch.complex64)).real * spec.mdct_to_raw_scale


def mel_spec_to_mdct_psd(mel_spec: Tensor, ms: MSDualSpec, spec: MDCTSpec) -> Tensor:
    """ms_mdct_dual.py:259-270: min-norm inverse of the mel filterbank (torch.linalg.lstsq, frequency_scale.py:136) of
    the linearised mel spectrogram; the PSD has ms.num_stft_bins - 1 bins (last bin cropped) when that equals
    mdct_psd_num_bins, otherwise its own filterbank with mdct_psd_num_bins STFT bins."""
    lin = (mel_spec - ms.raw_to_mel_spec_offset).float().clip(min=0) ** (1 / ms.ms_abs_exponent)
    if spec.mdct_psd_num_bins == ms.num_stft_bins - 1:
        fb, crop = ms_mel_filterbank(ms), True
    else:
        fb, crop = _mel_filterbank_bins(ms, spec.mdct_psd_num_bins), False
    shape = lin.shape
    sol = torch.linalg.lstsq(fb.t()[None], lin.reshape(-1, shape[-2], shape[-1]), driver="gels").solution   # :136, driver gels
    sol = sol.view(shape[:-2] + (fb.shape[0], shape[-1]))
    if crop:
        sol = sol[:, :, :-1, :]
    return sol * spec.mel_spec_to_mdct_psd_scale + spec.mel_spec_to_mdct_psd_offset


def _mel_filterbank_bins(spec: MSDualSpec, num_stft_bins: int) -> Tensor:
    """ms_mel_filterbank with an explicit number of STFT bins (the ms_freq_scale_mdct_psd scale, ms_mdct_dual.py:158-168)."""
    lo = 2595.0 * math.log10(1.0 + spec.ms_freq_min / 700.0)
    hi = 2595.0 * math.log10(1.0 + (spec.sample_rate / 2) / 700.0)
    f_pts = 700.0 * (10.0 ** (torch.linspace(lo, hi, spec.ms_num_frequencies + 2) / 2595.0) - 1.0)
    all_freqs = torch.linspace(0, spec.sample_rate / 2, num_stft_bins)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    fb = torch.max(torch.zeros(1), torch.min((-1.0 * slopes[:, :-2]) / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))
    enorm = 2.0 / (f_pts[2:spec.ms_num_frequencies + 2] - f_pts[:spec.ms_num_frequencies])
    return fb * enorm.unsqueeze(0)
