"""ncu target: one eager UNet forward (default config, 2x4x32x688) between cudaProfilerStart/Stop.
Writes the op trace (launch order) to gpurun_out/unet_trace.json so the ncu launch list can be attributed."""
import json, sys
import torch
sys.path.insert(0, ".")
from oracle import unet_oracle as uo
from dualdiffusion_b200 import ops
from dualdiffusion_b200.modules.unets.unet_edm2_b4 import UNet, UNetConfig
dev = torch.device("cuda:0")
spec = uo.default_spec(); sd = uo.synth_state_dict(spec, seed=0)
cfg = UNetConfig(**{k: getattr(spec, k) for k in UNetConfig.__dataclass_fields__ if hasattr(spec, k)})
net = UNet(cfg); net.load_state_dict(sd, strict=True)
net = net.requires_grad_(False).train(False).to(device=dev)
net.use_cuda_graphs = False
g = torch.Generator().manual_seed(1)
x = torch.randn(2, 4, 32, 688, generator=g).to(dev); sigma = torch.tensor([3.0, 3.0]).to(dev)
emb = net.get_embeddings(torch.randn(1, 512, generator=g), torch.tensor([True, False]))
for _ in range(2): net(x, sigma, None, emb)
torch.cuda.synchronize()
ops.trace = []
torch.cuda.profiler.start()
net(x, sigma, None, emb)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
json.dump(ops.trace, open("gpurun_out/unet_trace.json", "w"))
print("traced", len(ops.trace))
