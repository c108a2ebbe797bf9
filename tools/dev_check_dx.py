"""Developer GPU check of the tap-stacked 3x3 kernel (csrc/conv3x3_dx.cu): every epilogue mode and shape class against
torch on the GPU, then per-layer timing (L2 flushed between iterations).  Usage:
    python tools/dev_check_dx.py [check|time]          (DD_DISABLE_DX=1 times the round-1 halo kernel instead)"""
import math, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from dualdiffusion_b200 import ops, _lib as L

dev = "cuda"
torch.manual_seed(0)

def rel(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()

def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)

def ref_conv(x_nhwc, w, groups):
    xf = x_nhwc.float().permute(0, 3, 1, 2)
    wf = (w.float() / math.sqrt(w[0].numel())).to(torch.bfloat16).float()
    return F.conv2d(xf, wf, padding=1, groups=groups).permute(0, 2, 3, 1)

def check(B, H, W, Cin, Cout, g):
    x = nhwc(torch.randn(B, Cin, H, W, device=dev))
    w = torch.randn(Cout, Cin // g, 3, 3, device=dev)
    wp = ops.weight_prep(w)
    yr = ref_conv(x, w, g)
    errs = {}
    errs["none"] = rel(ops.mpconv(x, wp, 3, g), yr)
    sc = torch.randn(B, Cout, device=dev) * 0.3 + 1
    y, y2 = ops.mpconv(x, wp, 3, g, epi=L.EPI_SCALE_SILU, scale=sc, epi2=L.EPI2_RAW)
    errs["scale_silu"] = rel(y, F.silu(yr * sc[:, None, None, :]) / 0.596)
    errs["raw2"] = rel(y2, yr)
    res = nhwc(torch.randn(B, Cout, H, W, device=dev))
    sc2 = torch.randn(B, Cout, device=dev)
    y, y2 = ops.mpconv(x, wp, 3, g, epi=L.EPI_RESIDUAL, alpha=0.4, beta=0.9, clip=1.5, residual=res, epi2=L.EPI2_SCALE, scale2=sc2)
    ref = (0.4 * yr + 0.9 * res.float()).clamp(-1.5, 1.5)
    errs["residual_clip"] = rel(y, ref)
    errs["scale2"] = rel(y2, ref * sc2[:, None, None, :])
    y, y2 = ops.mpconv(x, wp, 3, g, epi=L.EPI_RESIDUAL, alpha=0.4, beta=0.9, residual=res, epi2=L.EPI2_SILU)
    ref = 0.4 * yr + 0.9 * res.float()
    errs["residual"] = rel(y, ref)
    errs["silu2"] = rel(y2, F.silu(ref) / 0.596)
    y = ops.mpconv(x, wp, 3, g, epi=L.EPI_RESIDUAL, alpha=0.7, beta=0.3, clip=256.0, residual=res)
    errs["residual_1out"] = rel(y, (0.7 * yr + 0.3 * res.float()).clamp(-256, 256))
    torch.cuda.synchronize()
    worst = max(errs.values())
    print(f"B{B} {H}x{W} {Cin}->{Cout} g{g}: worst {worst:.2e} " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()),
          "OK" if worst < 6e-3 else "FAIL", flush=True)
    return worst < 6e-3

def timeit(B, H, W, Cin, Cout, g, epi):
    x = nhwc(torch.randn(B, Cin, H, W, device=dev))
    wp = ops.weight_prep(torch.randn(Cout, Cin // g, 3, 3, device=dev))
    kw = {}
    if epi == 1: kw = dict(epi=1, scale=torch.ones(B, Cout, device=dev))
    if epi == 2: kw = dict(epi=2, alpha=0.7, beta=0.3, clip=256.0, residual=nhwc(torch.randn(B, Cout, H, W, device=dev)))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(3): ops.mpconv(x, wp, 3, g, **kw)
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.mpconv(x, wp, 3, g, **kw); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    # back-to-back (inputs / weights L2-warm), as inside the captured UNet graph
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.mpconv(x, wp, 3, g, **kw)
    e1.record(); torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) * 1e3 / 20
    us = sorted(ts)[len(ts) // 2]
    fl = 2 * B * H * W * Cout * (Cin // g) * 9
    print(f"B{B} {H}x{W} {Cin}->{Cout} g{g} epi{epi}: cold {us:7.1f} us {fl/us/1e6:6.1f} TF/s | back-to-back {warm:7.1f} us {fl/warm/1e6:6.1f} TF/s", flush=True)

mode = sys.argv[1] if len(sys.argv) > 1 else "check"
if mode == "check":
    ok = True
    for shp in [(1, 8, 30, 64, 64, 2), (1, 8, 16, 128, 64, 2), (2, 9, 37, 128, 128, 2), (3, 8, 64, 64, 128, 2),
                (1, 16, 24, 256, 512, 8), (2, 12, 61, 512, 256, 8), (1, 8, 172, 768, 1536, 8), (1, 8, 50, 1536, 768, 8),
                (1, 16, 70, 1280, 1024, 8), (1, 8, 40, 2048, 1024, 8), (2, 32, 688, 256, 512, 8), (2, 32, 688, 512, 256, 8),
                (1, 16, 344, 1024, 512, 8), (1, 32, 100, 768, 512, 8), (1, 8, 33, 96, 96, 1), (1, 10, 31, 160, 64, 1)]:
        ok &= check(*shp)
    print("ALL OK" if ok else "SOME FAILED")
else:
    for shp in [(2, 32, 688, 256, 512, 8, 1), (2, 32, 688, 512, 256, 8, 2), (2, 32, 688, 512, 256, 8, 0),
                (2, 32, 688, 512, 1024, 8, 1), (2, 32, 688, 1024, 512, 8, 0), (2, 32, 688, 768, 512, 8, 1),
                (2, 32, 688, 512, 512, 8, 1),
                (2, 16, 344, 512, 1024, 8, 1), (2, 16, 344, 1024, 512, 8, 2), (2, 16, 344, 1280, 1024, 8, 1),
                (2, 16, 344, 768, 1536, 8, 1), (2, 16, 344, 1536, 768, 8, 0),
                (2, 8, 172, 768, 1536, 8, 1), (2, 8, 172, 1536, 768, 8, 2), (2, 8, 172, 2048, 1024, 8, 0),
                (2, 8, 172, 1792, 1536, 8, 1)]:
        timeit(*shp)
