"""Developer GPU check: default UNet at the 45 s latent shape — parity vs the CPU oracle and timing
(eager launches vs CUDA-graph replay).  Usage: python tools/dev_check_unet.py [--no-oracle] [--iters N]"""
import sys, time
import torch
sys.path.insert(0, ".")
from oracle import unet_oracle as uo
from dualdiffusion_b200 import ops
from dualdiffusion_b200.modules.unets.unet_edm2_b4 import UNet, UNetConfig

dev = torch.device("cuda:0")
spec = uo.default_spec()
t0 = time.time(); sd = uo.synth_state_dict(spec, seed=0); print("weights", time.time() - t0, flush=True)
cfg = UNetConfig(**{k: getattr(spec, k) for k in UNetConfig.__dataclass_fields__ if hasattr(spec, k)})
net = UNet(cfg); net.load_state_dict(sd, strict=True)
net = net.requires_grad_(False).train(False).to(device=dev)
g = torch.Generator().manual_seed(1)
B = 2
x = torch.randn(B, 4, 32, 688, generator=g); sigma = torch.tensor([3.0, 3.0]); clap = torch.randn(1, 512, generator=g)
mask = torch.tensor([True, False])
emb = net.get_embeddings(clap, mask)
xd, sd_ = x.to(dev), sigma.to(dev)

def timeit(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.time(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.time() - t) / n * 1e3

n = 20
for a in sys.argv:
    if a.startswith("--iters="): n = int(a.split("=")[1])
net.use_cuda_graphs = False
before = ops.launch_count
d_eager = net(xd, sd_, None, emb)
print("launches per forward:", ops.launch_count - before)
ms, wall = timeit(lambda: net(xd, sd_, None, emb), n)
print(f"eager: {ms:.3f} ms device, {wall:.3f} ms wall per UNet call (B={B})")
net.use_cuda_graphs = True
d_graph = net(xd, sd_, None, emb)
ms, wall = timeit(lambda: net(xd, sd_, None, emb), n)
print(f"graph: {ms:.3f} ms device, {wall:.3f} ms wall per UNet call (B={B});  {0.489e12*B/ms/1e9:.1f} TFLOP/s")
print("graph vs eager max diff", (d_graph - d_eager).abs().max().item())
if "--no-oracle" not in sys.argv:
    t0 = time.time()
    ref = uo.unet_forward(sd, spec, x, sigma, uo.get_embeddings(sd, clap, mask))
    print("oracle CPU time", time.time() - t0)
    err = (d_graph.cpu() - ref).norm() / ref.norm()
    c_skip = 1 / (1 + sigma.view(-1, 1, 1, 1) ** 2)
    body = ((d_graph.cpu() - c_skip * x) - (ref - c_skip * x)).norm() / (ref - c_skip * x).norm()
    print(f"full-size parity vs oracle: rel {err:.3e}, body-only rel {body:.3e}")
