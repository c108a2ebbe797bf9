"""ncu target: one DAE_D3.decode (default config, batch 2 latents 8x32x688) between cudaProfilerStart/Stop."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import dae_oracle as do
from dualdiffusion_b200.modules.daes.dae_edm2_d3 import DAE_D3, DAE_D3_Config
dev = torch.device("cuda:0")
spec = do.DAESpec()
net = DAE_D3(DAE_D3_Config(channel_mult_enc=spec.channel_mult_enc))
net.load_state_dict(do.synth_dae_state_dict(spec, seed=0), strict=True)
net = net.requires_grad_(False).train(False).to(dev)
net.use_cuda_graphs = False
g = torch.Generator().manual_seed(1)
lat = torch.randn(2, 8, 32, 688, generator=g).to(dev)
emb = net.get_embeddings(torch.randn(2, spec.in_channels_emb, generator=g))
for _ in range(2): net.decode(lat, emb)
torch.cuda.synchronize()
torch.cuda.profiler.start()
net.decode(lat, emb)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
