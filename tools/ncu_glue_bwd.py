"""ncu target: one silu_scale_bwd and one pixnorm_silu_bwd launch at the train step's level-0 shapes (batch 4)."""
import sys
import torch
sys.path.insert(0, ".")
from dualdiffusion_b200 import ops
dev = torch.device("cuda:0")
B, H, W = 4, 32, 688
dy = torch.randn(B, H, W, 512, device=dev).to(torch.bfloat16); pre = torch.randn(B, H, W, 512, device=dev).to(torch.bfloat16)
scale = 1 + 0.1 * torch.randn(B, 512, device=dev); dscale = torch.zeros(B, 512, device=dev)
g, ds, t0 = (torch.randn(B, H, W, 256, device=dev).to(torch.bfloat16) for _ in range(3))
ops.silu_scale_bwd(dy, 0.7, pre, scale, dscale); ops.pixnorm_silu_bwd(g, 0.8, ds, t0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.silu_scale_bwd(dy, 0.7, pre, scale, dscale)
ops.pixnorm_silu_bwd(g, 0.8, ds, t0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
