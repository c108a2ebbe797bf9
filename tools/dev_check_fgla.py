import sys, torch
sys.path.insert(0, ".")
from oracle import format_oracle as fo
from dualdiffusion_b200 import ops
from dualdiffusion_b200.modules.formats.spectrogram import SpectrogramFormat, SpectrogramFormatConfig
dev = torch.device("cuda")
g = torch.load("tests/golden/format_small.pt", weights_only=False)
spec = fo.SpectrogramSpec()
fmt = SpectrogramFormat(SpectrogramFormatConfig())
c = fmt.config
t = fmt._tables(dev)
mel = g["mel"]
B, C, F, T = mel.shape; S = B * C
s = (mel / c.raw_to_sample_scale + c.sample_mean).clip(min=0) ** 4
mag_ref = fo.unscale(s, spec)                              # [B,C,K,T]
mag = torch.matmul(s.view(S, F, T).transpose(1, 2).to(dev), t["pinv_t"]).clamp_(min=0)   # [S,T,K]
def rel(a, b): return ((a.double().cpu() - b.double().cpu()).norm() / b.double().norm()).item()
print("mag vs lstsq", rel(mag.transpose(1, 2), mag_ref.view(S, -1, T)))
n_fft, hop = c.padded_length, c.hop_length
w = fo.window(spec)
kw = dict(n_fft=n_fft, hop_length=hop, win_length=n_fft, window=w)
# (b) istft of zero-phase magnitudes
X = mag_ref.view(S, -1, T).to(torch.cfloat)
y_ref = torch.istft(X, length=None, **kw)
env = fmt._envelope(t, T, dev)
ola = torch.empty((S, n_fft + hop * (T - 1)), device=dev)
args = (t["window"], t["tw"], t["tw_half"], n_fft, hop)
magx = mag_ref.view(S, -1, T).transpose(1, 2).contiguous().to(dev)
ops.fgla_istft(None, magx, False, 0.0, *args, ola)
y = ops.ola_finalize(ola, env, n_fft, hop * (T - 1))
print("istft vs torch", rel(y, y_ref), "first", y[0, :4].tolist(), y_ref[0, :4].tolist(), "last", y[0, -3:].tolist(), y_ref[0, -3:].tolist())
e_ref = torch.istft(torch.ones_like(X), length=None, **kw)  # not meaningful
# (c) stft of y_ref
R_ref = torch.stft(y_ref, center=True, pad_mode="reflect", normalized=False, onesided=True, return_complex=True, **kw)   # [S,K,T]
ola2 = torch.zeros_like(ola); ola2[:, n_fft // 2: n_fft // 2 + y_ref.shape[-1]] = y_ref.to(dev) * env[n_fft // 2: n_fft // 2 + y_ref.shape[-1]]
state = torch.empty((S, T, n_fft // 2 + 1, 2), device=dev)
ops.fgla_stft_update(ola2, env, state, 0.33, True, *args)
R = torch.view_as_complex(state).transpose(1, 2)
print("stft vs torch", rel(torch.view_as_real(R), torch.view_as_real(R_ref)))
# (d) istft with complex state
st = torch.view_as_real(R_ref.transpose(1, 2).contiguous()).contiguous().to(dev)
ops.fgla_istft(st, magx, False, 0.0, *args, ola)
y2 = ops.ola_finalize(ola, env, n_fft, hop * (T - 1))
ang = R_ref / (R_ref.abs() + 1e-16)
y2_ref = torch.istft(ang * mag_ref.view(S, -1, T), length=None, **kw)
print("istft(state) vs torch", rel(y2, y2_ref))
# stereo merge
ops.fgla_istft(None, magx, True, -0.5, *args, ola)
y3 = ops.ola_finalize(ola, env, n_fft, hop * (T - 1))
m = mag_ref.view(S, -1, T); merged = ((m[0::2] + m[1::2]) / 2).repeat_interleave(2, dim=0)
print("istft merged vs torch", rel(y3, torch.istft(merged.to(torch.cfloat), length=None, **kw)))
# envelope vs torch's
print("env check: y*env ratio", (y[0, :3] / y_ref[0, :3].to(dev)).tolist())
