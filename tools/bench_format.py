"""Developer GPU benchmark: mel-STFT encode + FGLA decode at BASELINE config 3 (batch 64 stereo, 45 s)."""
import sys, time, torch
sys.path.insert(0, ".")
from dualdiffusion_b200.modules.formats.spectrogram import SpectrogramFormat, SpectrogramFormatConfig
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
fmt = SpectrogramFormat(SpectrogramFormatConfig())
L = fmt.sample_raw_crop_width(1408768)
print("crop", L, fmt.get_sample_shape(B, 1408768))
g = torch.Generator(device=dev).manual_seed(0)
raw = 0.1 * torch.randn(B, 2, L, device=dev, generator=g)
def timed(fn, n=1):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out
fmt.raw_to_sample(raw[:2])
ms, mel = timed(lambda: fmt.raw_to_sample(raw), 3)
frames = B * 2 * mel.shape[-1]
print(f"encode: {ms:.2f} ms for B={B} -> {B/ms*1e3:.1f} stereo items/s, {frames/ms*1e3/1e6:.2f} Mframes/s, algorithmic {(raw.numel()+mel.numel())*4/ms/1e6:.1f} GB/s")
fmt.sample_to_raw(mel[:1], n_fgla_iters=2)
ms, wave = timed(lambda: fmt.sample_to_raw(mel, n_fgla_iters=iters), 1)
S = 3201 * mel.shape[-1]
alg = (28 * S + 2 * 4 * L) * (B * 2) * iters
print(f"FGLA: {ms:.1f} ms for B={B}, {iters} iters -> {B/ms*1e3:.2f} stereo items/s; {ms/iters:.2f} ms/iter; algorithmic {alg/ms/1e6:.0f} GB/s")
print("peak mem GB", torch.cuda.max_memory_allocated() / 1e9, "wave std", wave.std().item())
