"""Golden vectors for the second MS_MDCT_DualFormat lineage from the UNMODIFIED reference
(src/modules/formats/ms_mdct_dual_2.py, default configuration) on CPU: raw_to_mel_spec, mel_spec_to_linear, raw_to_mdct,
mdct_to_raw, raw_to_mdct_phase_psd and the shape helpers on a short seeded stereo signal.
    python tests/golden/make_golden_ms_dual2.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
from modules.formats.ms_mdct_dual_2 import MS_MDCT_DualFormat, MS_MDCT_DualFormatConfig  # noqa: E402

torch.set_grad_enabled(False)
fmt = MS_MDCT_DualFormat(MS_MDCT_DualFormatConfig())
g = torch.Generator().manual_seed(61)
n = 256 * 40
raw = 0.05 * torch.randn(2, 2, n, generator=g)
t = torch.arange(n) / 32000.0
raw[0, 0] += 0.3 * torch.sin(2 * torch.pi * 440.0 * t)
raw[0, 1] += 0.2 * torch.sin(2 * torch.pi * 3520.0 * t + 0.5)
raw[1, 0] += 0.25 * torch.sin(2 * torch.pi * 95.0 * t)
raw_odd = raw[:1, :, :256 * 9 + 77].contiguous()                     # a length that is not a multiple of the hop
mel = fmt.raw_to_mel_spec(raw)
mdct = fmt.raw_to_mdct(raw)
phase, psd = fmt.raw_to_mdct_phase_psd(raw)
out = dict(raw=raw, raw_odd=raw_odd, mel=mel, mel_linear=fmt.mel_spec_to_linear(mel), mdct=mdct, mdct_odd=fmt.raw_to_mdct(raw_odd),
           raw_back=fmt.mdct_to_raw(mdct), phase=phase, psd=psd,
           mel_spec_shape=tuple(fmt.get_mel_spec_shape(3, 1408768)), mdct_shape=tuple(fmt.get_mdct_shape(3, 1408768)),
           raw_crop_width=int(fmt.get_raw_crop_width(1408768)), window_weights=fmt.ms_filter_window_weights.clone())
torch.save(out, os.path.join(ROOT, "tests", "golden", "ms_dual2_small.pt"))
print({k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in out.items()})
print("mel std", float(mel.std()), "mdct std", float(mdct.std()), "round trip", float((out["raw_back"] - raw).abs().max()))
