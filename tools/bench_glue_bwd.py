"""Device-time of the backward glue kernels at the train step's level-0 / level-1 shapes (batch 4), cycling through
buffer sets larger than L2 (cold inputs): achieved HBM bandwidth per launch."""
import sys
import torch
sys.path.insert(0, ".")
from dualdiffusion_b200 import ops

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)


def timed(fn, sets, reps=3):
    for s in sets:
        fn(*s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for s in sets:
            fn(*s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(sets))


for (B, H, W, C) in ((4, 32, 688, 512), (4, 16, 344, 1024), (4, 8, 172, 1536)):
    n_sets = max(2, int(600e6 // (3 * B * H * W * C * 2)) + 1)
    sets = []
    for _ in range(n_sets):
        dy = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
        pre = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
        scale = 1 + 0.1 * torch.randn(B, C, device=dev)
        dscale = torch.zeros(B, C, device=dev)
        sets.append((dy, pre, scale, dscale))
    us = timed(lambda dy, pre, scale, dscale: ops.silu_scale_bwd(dy, 0.7, pre, scale, dscale), sets)
    mb = 3 * B * H * W * C * 2 / 1e6
    print(f"silu_scale_bwd {B}x{H}x{W}x{C}: {us:7.1f} us  {mb / us * 1e-3 * 1e3:6.2f} GB/ms = {mb / us:5.2f} TB/s x1e-3".replace(" x1e-3", ""), flush=True)
    del sets
    torch.cuda.empty_cache()

for (B, H, W, C) in ((4, 32, 688, 256), (4, 16, 344, 512)):
    n_sets = max(2, int(600e6 // (4 * B * H * W * C * 2)) + 1)
    sets = []
    for _ in range(n_sets):
        sets.append(tuple(torch.randn(B, H, W, C, device=dev).to(torch.bfloat16) for _ in range(3)))
    us = timed(lambda g, ds, t0: ops.pixnorm_silu_bwd(g, 0.8, ds, t0), sets)
    mb = 4 * B * H * W * C * 2 / 1e6
    print(f"pixnorm_silu_bwd {B}x{H}x{W}x{C}: {us:7.1f} us  {mb / us:5.2f} TB/s", flush=True)
    del sets
    torch.cuda.empty_cache()
