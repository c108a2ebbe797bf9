"""ncu target: one UNet train step (default config, device batch 4 x 4x32x688, fwd + bwd) between
cudaProfilerStart/Stop.  Writes the op trace (launch order) to gpurun_out/train_trace.json."""
import json, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from oracle import unet_oracle as uo
from dualdiffusion_b200 import ops
from dualdiffusion_b200.ddp import GradAllReducer
from dualdiffusion_b200.modules.unets.unet_edm2_b4 import UNet, UNetConfig
dev = torch.device("cuda:0")
spec = uo.default_spec(); sd = uo.synth_state_dict(spec, seed=0)
cfg = UNetConfig(**{k: getattr(spec, k) for k in UNetConfig.__dataclass_fields__ if hasattr(spec, k)})
net = UNet(cfg); net.load_state_dict(sd, strict=True)
net = net.to(dev).train()
net.grad_sync = GradAllReducer()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
g = torch.Generator().manual_seed(1)
x = torch.randn(B, 4, 32, 688, generator=g).to(dev); sigma = torch.full((B,), 1.5).to(dev)
clap = torch.randn(B, 512, generator=g).to(dev); mask = torch.ones(B, dtype=torch.bool, device=dev)
def step():
    net.zero_grad(set_to_none=True)
    emb = net.get_embeddings(clap, mask)
    d = net(x, sigma, None, emb)
    loss = F.mse_loss(d, x)
    loss.backward()
for _ in range(2): step()
torch.cuda.synchronize()
ops.trace = []
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
json.dump(ops.trace, open("gpurun_out/train_trace.json", "w"))
print("traced", len(ops.trace))
