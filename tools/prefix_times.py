"""In-graph cost of every launch of one UNet evaluation (default config, 2x4x32x688).

For n = 0..N the launch schedule is captured into a CUDA graph with every C-ABI call after the n-th skipped (buffers
are still allocated, so the schedule is unchanged), and the graph is replayed; T(n) - T(n-1) is what launch n adds to
the captured call *in situ* (warm L2 / overlapped branches / dependent-launch overlap included), which the serialised
cold-cache ncu list cannot show.  Writes gpurun_out/prefix_times.csv.

    python tools/prefix_times.py [stride]
"""
import sys
import torch
sys.path.insert(0, ".")
from oracle import unet_oracle as uo
from dualdiffusion_b200 import ops, _lib as L
from dualdiffusion_b200.modules.unets.unet_edm2_b4 import UNet, UNetConfig


class Gate:
    """ctypes library proxy: calls past `limit` return 0 without launching."""

    def __init__(self, lib):
        self.lib, self.n, self.limit, self.names = lib, 0, 1 << 30, []

    def __getattr__(self, name):
        fn = getattr(self.lib, name)
        if name in ("dd_last_error", "dd_abi_version"):
            return fn

        def call(*a):
            self.n += 1
            if self.names is not None:
                self.names.append(name[3:] + ":" + "x".join(str(v) for v in a[3:10] if isinstance(v, int) and v < 100000))
            return fn(*a) if self.n <= self.limit else 0
        return call


def main():
    stride = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    dev = torch.device("cuda:0")
    spec = uo.default_spec(); sd = uo.synth_state_dict(spec, seed=0)
    cfg = UNetConfig(**{k: getattr(spec, k) for k in UNetConfig.__dataclass_fields__ if hasattr(spec, k)})
    net = UNet(cfg); net.load_state_dict(sd, strict=True)
    net = net.requires_grad_(False).train(False).to(device=dev)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 4, 32, 688, generator=g).to(dev); sigma = torch.tensor([3.0, 3.0]).to(dev)
    emb = net.get_embeddings(torch.randn(1, 512, generator=g), torch.tensor([True, False])).float()
    net.use_cuda_graphs = False
    for _ in range(2):
        net(x, sigma, None, emb)
    plan = net._get_plan(); plan.refresh_weights()
    lf = net._ln_freqs(plan, None, 32)
    gate = Gate(L.load()); L._lib = gate
    with torch.no_grad():
        net._run(plan, x, x, sigma, emb, lf, None)
    names, gate.names = gate.names, None
    N = gate.n
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = {}
    cuts = sorted(set(list(range(1, N + 1, stride)) + [N]))
    for n in cuts:
        gate.n, gate.limit = 0, n
        graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(graph):
            net._run(plan, x, x, sigma, emb, lf, None)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            ev0.record()
            for _ in range(10):
                graph.replay()
            ev1.record(); torch.cuda.synchronize()
            best = min(best, ev0.elapsed_time(ev1) * 100.0)      # us per replay
        times[n] = best
        del graph
    with open("gpurun_out/prefix_times.csv", "w") as fh:
        fh.write("n,name,prefix_us,delta_us\n")
        prev = None
        for n in cuts:
            d = "" if prev is None else f"{times[n] - times[prev]:.2f}"
            fh.write(f"{n},{names[n - 1] if n > 0 else ''},{times[n]:.2f},{d}\n")
            prev = n
    print("launches", N, "full", times[N])


if __name__ == "__main__":
    main()
