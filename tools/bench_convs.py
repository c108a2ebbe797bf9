"""Developer GPU microbenchmark: every distinct MPConv shape of the default UNet at the 45 s latent
(B=2), CUDA-event timed with an L2 flush between iterations.  Usage: python tools/bench_convs.py [tag]"""
import json, sys, collections
import torch
sys.path.insert(0, ".")
from dualdiffusion_b200 import ops, _lib as L
dev = "cuda"
tr = json.load(open("profiles/r01_unet_fwd_trace_v1.json"))
shapes = collections.OrderedDict()
for op, det in tr:
    if op == "mpconv":
        shapes[tuple(det)] = shapes.get(tuple(det), 0) + 1
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
tot = 0.0
rows = []
for (B, H, W, Cin, Cout, k, g, e1, e2), cnt in shapes.items():
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    wp = ops.weight_prep(torch.randn(Cout, Cin // g, k, k, device=dev))
    kw = {}
    if e1 == 1: kw = dict(epi=1, scale=torch.ones(B, Cout, device=dev))
    if e1 == 2: kw = dict(epi=2, alpha=0.5, beta=0.5, residual=torch.randn(B, H, W, Cout, device=dev).to(torch.bfloat16))
    if e2 == 2: kw.update(epi2=2, scale2=torch.ones(B, Cout, device=dev))
    for _ in range(2): ops.mpconv(x, wp, k, g, **kw)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.mpconv(x, wp, k, g, **kw); e1_.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1_) * 1e3)
    us = sorted(ts)[len(ts) // 2]
    fl = 2 * B * H * W * Cout * (Cin // g) * k * k
    tot += us * cnt
    rows.append(((B, H, W, Cin, Cout, k, g, e1, e2), cnt, us, fl / us / 1e6))
    print(f"{(B,H,W,Cin,Cout,k,g,e1,e2)} x{cnt}: {us:7.1f} us {fl/us/1e6:7.1f} TF/s", flush=True)
print("TOTAL conv time per UNet call (us):", tot)
tag = sys.argv[1] if len(sys.argv) > 1 else "run"
json.dump(rows, open(f"gpurun_out/bench_convs_{tag}.json", "w"))
