"""The DAE-decode and ddec-forward legs of bench.py alone (A/B runs of kernel dispatch knobs)."""
import sys
import torch
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
d = bench.bench_dae(dev)
print("dae_decode", round(d["value"], 2), "samples/s", round(d["roofline"]["frac"], 4))
d = bench.bench_ddec(dev)
print("ddec_forward", round(d["value"], 2), "samples/s", round(d["roofline"]["frac"], 4))
