"""In-graph time of one MPConv layer with each fused epilogue:  python tools/layer_time.py B H W Cin Cout k groups"""
import sys
import torch
sys.path.insert(0, ".")
from dualdiffusion_b200 import ops, _lib as L

B, H, W, Ci, Co, k, g = [int(a) for a in sys.argv[1:8]]
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(0)
x = torch.randn(B, H, W, Ci, generator=gen).to(dev, torch.bfloat16)
wp = ops.weight_prep(torch.randn(Co, Ci // g, k, k, generator=gen).to(dev))
res = torch.randn(B, H, W, Co, generator=gen).to(dev, torch.bfloat16)
sc = torch.ones(B, Co, device=dev)
out = torch.empty(B, H, W, Co, device=dev, dtype=torch.bfloat16)
out2 = torch.empty_like(out)
flop = 2.0 * B * H * W * Co * (Ci // g) * k * k
cases = {"none": dict(), "scale_silu": dict(epi=L.EPI_SCALE_SILU, scale=sc),
         "residual": dict(epi=L.EPI_RESIDUAL, alpha=0.7, beta=0.3, clip=256.0, residual=res),
         "residual+silu2": dict(epi=L.EPI_RESIDUAL, alpha=0.7, beta=0.3, clip=256.0, residual=res, epi2=L.EPI2_SILU, out2=out2)}
for name, kw in cases.items():
    f = lambda: ops.mpconv(x, wp, k, g, out=out, **kw)
    f(); torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(8):
            f()
    graph.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 40
    print(f"{name:16s} {us:9.1f} us  {flop / us / 1e6:7.1f} TFLOP/s")
