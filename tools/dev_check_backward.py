"""Prints the measured error of every parameter gradient of the assembled train step (small config) against
autograd through the CPU oracle -- the numbers the tolerances in tests/test_gpu_backward.py are set from."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, rel_err  # noqa: E402
from oracle import unet_oracle as uo  # noqa: E402
from test_gpu_backward import make_train_unet, product_loss  # noqa: E402

dev = torch.device("cuda:0")
spec = uo.small_spec()
sd = uo.synth_state_dict(spec, seed=0)
g = load_golden("unet_small_train.pt")
net = make_train_unet(spec, sd, dev)
loss, den = product_loss(net, spec, g["samples"], g["noise"], g["sigma"], g["clap"], g["mask"], dev)
print("loss", float(loss), "ref", float(g["loss"]), "denoised rel", rel_err(den, g["denoised"]))
loss.backward()
sdg = {k: (v.clone().requires_grad_(True) if "fourier" not in k else v) for k, v in sd.items()}
uo.train_loss(sdg, spec, g["samples"], g["noise"], g["sigma"], g["clap"], g["mask"]).backward()
rows = []
for name, p in net.named_parameters():
    ref = sdg[name].grad
    got = p.grad.float().cpu()
    cos = float((got * ref).sum() / (got.norm() * ref.norm() + 1e-30))
    rows.append((rel_err(got, ref), name, float(ref.norm()), float(got.norm()), cos, g["grad_stats"][name][0]))
for r in sorted(rows, reverse=True):
    print(f"{r[0]:9.4f}  cos {r[4]:8.5f}  |ref| {r[2]:10.3e} |got| {r[3]:10.3e} |golden| {r[5]:10.3e}  {r[1]}")
