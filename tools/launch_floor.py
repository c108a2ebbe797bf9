"""Dependent-launch floor inside a CUDA graph: chains of tiny kernels at the coarsest UNet level (2x2x43 pixels), each
consuming the previous one's output.  Run twice (with / without DD_DISABLE_PDL=1) to see what programmatic dependent
launch buys.  Prints us per launch."""
import sys
import torch
sys.path.insert(0, ".")
from dualdiffusion_b200 import ops, _lib as L

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
N = 64


def timed(fn, label):
    fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record()
        for _ in range(10):
            graph.replay()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 100.0 / N)
    print(f"{label:60s} {best:7.2f} us per launch")


x = torch.randn(2, 2, 43, 1280, generator=g).to(dev, torch.bfloat16)
w1 = ops.weight_prep(torch.randn(1280, 1280, 1, 1, generator=g).to(dev))
w3 = ops.weight_prep(torch.randn(1280, 320, 3, 3, generator=g).to(dev))     # 2560 -> 1280, g8
x2 = torch.randn(2, 2, 43, 2560, generator=g).to(dev, torch.bfloat16)
xs = torch.randn(2, 4, 86, 1024, generator=g).to(dev, torch.bfloat16)
ws = ops.weight_prep(torch.randn(1024, 1024, 1, 1, generator=g).to(dev))


def chain_pixnorm():
    t = x
    for _ in range(N):
        t, _s = ops.pixnorm_silu(t)


def chain_avgpool_like():
    t = x
    for _ in range(N):
        t = ops.axpby(t, t, 0.5, 0.5)


def chain_conv1():
    t = x
    for _ in range(N):
        t = ops.mpconv(t, w1, 1)


def chain_conv1_l3():
    t = xs
    for _ in range(N):
        t = ops.mpconv(t, ws, 1)


def chain_conv3():
    t = x2
    for _ in range(N // 2):
        t = ops.mpconv(t, w3, 3, 8)                  # 2560 -> 1280
        t = torch.cat([t, t], dim=-1) if False else t
        t = ops.mpconv(x2, w3, 3, 8, epi=L.EPI_RESIDUAL, alpha=1.0, beta=1.0, residual=t)


def indep_conv1():
    for _ in range(N):
        ops.mpconv(x, w1, 1)


qk86 = torch.randn(2, 2, 43, 2560, generator=g).to(dev, torch.bfloat16)
v86 = torch.randn(2, 2, 43, 1280, generator=g).to(dev, torch.bfloat16)
qk344 = torch.randn(2, 4, 86, 2048, generator=g).to(dev, torch.bfloat16)
v344 = torch.randn(2, 4, 86, 1024, generator=g).to(dev, torch.bfloat16)
sc20 = torch.ones(2, 1280, device=dev)
sc16 = torch.ones(2, 1024, device=dev)


def chain_attn86():
    for _ in range(N):
        ops.attention(qk86, v86, sc20, 20, 64)


def chain_attn344():
    for _ in range(N):
        ops.attention(qk344, v344, sc16, 16, 64)


timed(chain_attn86, "attention 86 tokens x 20 heads x batch 2")
timed(chain_attn344, "attention 344 tokens x 16 heads x batch 2")
timed(chain_pixnorm, "pixnorm_silu 172 x 1280, dependent chain")
timed(chain_avgpool_like, "axpby 172 x 1280, dependent chain")
timed(chain_conv1, "1x1 conv 1280->1280 at 2x2x43, dependent chain")
timed(indep_conv1, "1x1 conv 1280->1280 at 2x2x43, same input (no data dependence)")
timed(chain_conv1_l3, "1x1 conv 1024->1024 at 2x4x86, dependent chain")
timed(chain_conv3, "3x3 g8 conv 2560->1280 at 2x2x43, dependent chain")
