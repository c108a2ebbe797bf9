"""Profiling driver for the optimizer-side sweep (run under ncu by tools/r02_first_gpu_session.sh): a few
clip_grad_norm_ + step() iterations of FusedAdamW over the default UNet's 293 M parameters with two fp32 EMA copies."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as uo  # noqa: E402  (spec only)
from dualdiffusion_b200.modules.unets.unet_edm2_b4 import UNet, UNetConfig  # noqa: E402
from dualdiffusion_b200.training.optim import FusedAdamW  # noqa: E402


def main() -> None:
    dev = torch.device("cuda:0")
    spec = uo.default_spec()
    cfg = UNetConfig(**{k: getattr(spec, k) for k in UNetConfig.__dataclass_fields__ if hasattr(spec, k)})
    net = UNet(cfg).to(dev).train()
    params = list(net.parameters())
    g = torch.Generator(device=dev).manual_seed(7)
    for p in params:
        p.grad = torch.randn(p.shape, device=dev, generator=g)
    emas = [[p.detach().clone() for p in params] for _ in range(2)]
    opt = FusedAdamW(params, lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.0)
    opt.attach_module(net)
    opt.attach_emas(emas, [0.9999, 0.99999], [0.9999, None])
    for _ in range(4):
        norm = opt.clip_grad_norm_(10.0)
        opt.step()
    torch.cuda.synchronize()
    n = sum(p.numel() for p in params)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        opt.clip_grad_norm_(10.0)
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"grad norm {float(norm):.4f} params {n}  sweep {ms:.3f} ms = {n * 52 / ms / 1e6:.0f} GB/s algorithmic "
          f"(52 B/param; DD_OPTIM_SMEM_SEARCH={os.environ.get('DD_OPTIM_SMEM_SEARCH', '0')} "
          f"DD_OPTIM_PERSISTENT={os.environ.get('DD_OPTIM_PERSISTENT', '0')})")


if __name__ == "__main__":
    main()
