"""Experiment: does UMMA honour a row-shifted start address inside a 128B-swizzled tile, and what does the
descriptor's base_offset field need to be?  1x1 conv, one 128-pixel tile (W=128,H=1): output row m should
equal the un-shifted result of row m+shift."""
import os, sys
import torch
sys.path.insert(0, ".")
from dualdiffusion_b200 import ops
dev = "cuda"
torch.manual_seed(0)
x = torch.randn(1, 1, 128, 64, device=dev).to(torch.bfloat16)
wp = ops.weight_prep(torch.randn(64, 64, 1, 1, device=dev))
os.environ.pop("DD_DBG_SHIFT", None); os.environ.pop("DD_DBG_BO", None)
y0 = ops.mpconv(x, wp, 1).float()
for shift in (1, 2, 3, 8):
    for bo in (0, shift % 8):
        os.environ["DD_DBG_SHIFT"] = str(shift); os.environ["DD_DBG_BO"] = str(bo)
        y = ops.mpconv(x, wp, 1).float()
        torch.cuda.synchronize()
        n = 128 - shift - 8
        err = (y[0, 0, :n] - y0[0, 0, shift:shift + n]).abs().max().item()
        print(f"shift {shift} base_offset {bo}: max err {err:.4f} (ref scale {y0.abs().max().item():.2f})")
