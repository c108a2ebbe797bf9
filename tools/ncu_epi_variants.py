"""Four launches of the 512->1024 grouped 3x3 layer (2 x 32 x 688) for an ncu capture of the two epilogue variants:
residual + silu copy staged / direct, residual only staged / direct (in that order).
  ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -c 4 -o gpurun_out/epi_variants \
      python tools/ncu_epi_variants.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dualdiffusion_b200 import ops, _lib as L   # noqa: E402

dev = torch.device("cuda:0")
B, H, W, Cin, Cout, k, g = 2, 32, 688, 512, 1024, 3, 8
x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
wp = ops.weight_prep(torch.randn(Cout, Cin // g, k, k, device=dev))
res = torch.randn(B, H, W, Cout, device=dev).to(torch.bfloat16)
out, out2 = torch.empty_like(res), torch.empty_like(res)
for kw in (dict(epi2=L.EPI2_SILU, out2=out2), dict()):
    for staged in ("1", "0"):
        os.environ["DD_EPI_STAGED"] = staged
        ops.mpconv(x, wp, k, g, epi=L.EPI_RESIDUAL, alpha=0.5, beta=0.5, residual=res, out=out, **kw)
        torch.cuda.synchronize()
