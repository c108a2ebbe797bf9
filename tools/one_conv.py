"""ncu target: a few launches of one MPConv shape.  Usage: python tools/one_conv.py B H W Cin Cout k g [epi]"""
import sys, torch
sys.path.insert(0, ".")
from dualdiffusion_b200 import ops
B, H, W, Cin, Cout, k, g = map(int, sys.argv[1:8])
epi = int(sys.argv[8]) if len(sys.argv) > 8 else 0
dev = "cuda"
x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
wp = ops.weight_prep(torch.randn(Cout, Cin // g, k, k, device=dev))
kw = dict(epi=1, scale=torch.ones(B, Cout, device=dev)) if epi == 1 else {}
for _ in range(4): ops.mpconv(x, wp, k, g, **kw)
torch.cuda.synchronize()
