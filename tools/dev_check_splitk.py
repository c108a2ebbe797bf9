"""Developer GPU check of the split-K path of conv_igemm (levels 3-4 shapes): parity vs torch and timing.
    python tools/dev_check_splitk.py          (DD_DISABLE_SPLITK=1 for the A/B)"""
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
def nhwc(x): return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
def ref_conv(x, w, g):
    wf = (w.float() / math.sqrt(w[0].numel())).to(torch.bfloat16).float()
    return F.conv2d(x.float().permute(0, 3, 1, 2), wf, padding=w.shape[-1] // 2, groups=g).permute(0, 2, 3, 1)
shapes = [(2, 2, 43, 1280, 1280, 1, 1), (2, 2, 43, 2560, 1280, 3, 8), (2, 2, 43, 1280, 2560, 1, 1), (2, 4, 86, 1024, 1024, 1, 1),
          (2, 4, 86, 2048, 1024, 3, 8), (2, 4, 86, 1024, 2048, 1, 1), (2, 2, 43, 1280, 2560, 3, 8), (2, 4, 86, 768, 768, 1, 1),
          (2, 2, 43, 2304, 2560, 3, 8), (1, 3, 7, 256, 96, 1, 1), (2, 8, 172, 768, 768, 1, 1), (2, 4, 86, 2560, 1280, 3, 8)]
ok = True
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for (B, H, W, Cin, Cout, k, g) in shapes:
    x = nhwc(torch.randn(B, Cin, H, W, device=dev)); w = torch.randn(Cout, Cin // g, k, k, device=dev)
    wp = ops.weight_prep(w); yr = ref_conv(x, w, g)
    res = nhwc(torch.randn(B, Cout, H, W, device=dev)); sc = torch.randn(B, Cout, device=dev) * 0.3 + 1
    errs = []
    for _ in range(3):      # repeated launches reuse workspace slots: they must come back clean
        errs.append(rel(ops.mpconv(x, wp, k, g), yr))
        y, y2 = ops.mpconv(x, wp, k, g, epi=L.EPI_SCALE_SILU, scale=sc, epi2=L.EPI2_RAW)
        errs.append(rel(y, F.silu(yr * sc[:, None, None, :]) / 0.596)); errs.append(rel(y2, yr))
        y, y2 = ops.mpconv(x, wp, k, g, epi=L.EPI_RESIDUAL, alpha=0.4, beta=0.9, clip=1.5, residual=res, epi2=L.EPI2_SCALE, scale2=sc)
        r_ = (0.4 * yr + 0.9 * res.float()).clamp(-1.5, 1.5)
        errs.append(rel(y, r_)); errs.append(rel(y2, r_ * sc[:, None, None, :]))
    torch.cuda.synchronize()
    for _ in range(3): ops.mpconv(x, wp, k, g)
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.mpconv(x, wp, k, g); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    # in-graph: 20 launches captured once and replayed (no host launch cost: the regime of the captured UNet)
    out = torch.empty(B, H, W, Cout, device=dev, dtype=torch.bfloat16)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(20): ops.mpconv(x, wp, k, g, out=out)
    torch.cuda.current_stream().wait_stream(side)
    graph.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): graph.replay()
    e1.record(); torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) * 1e3 / 100
    good = max(errs) < 6e-3
    ok &= good
    print(f"{(B,H,W,Cin,Cout,k,g)}: worst rel {max(errs):.2e} {'OK' if good else 'FAIL'} | cold {sorted(ts)[3]:6.1f} us, in-graph {warm:6.1f} us", flush=True)
print("ALL OK" if ok else "SOME FAILED")
