"""Developer tool: host-side profile of the e2e sampler step (pinned H2D -> step -> D2H -> sync)."""
import cProfile, pstats, sys, time, torch
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda:0")
with torch.inference_mode():
    net, pipe, state, gen = bench.build_sampler(dev, 100)
    shape = tuple(state.sample.shape)
    noise = torch.randn(shape, device=dev)
    def step(i): state.step(3 + (i % 90), noise)
    for i in range(4): step(i)
    torch.cuda.synchronize()
    host_in = torch.empty(shape, dtype=torch.float32).pin_memory(); host_out = torch.empty(shape, dtype=torch.float32).pin_memory()
    host_in.copy_(state.sample.cpu())
    def e2e(K):
        global host_in, host_out
        t0 = time.perf_counter()
        for i in range(K):
            state.set_sample(host_in); step(i); host_out.copy_(state.sample, non_blocking=True); torch.cuda.synchronize()
            host_in, host_out = host_out, host_in
        return (time.perf_counter() - t0) / K * 1e3
    e2e(3)
    print("e2e ms/step", e2e(20))
    # host time to enqueue one step without sync
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(20): step(i)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("enqueue ms/step", (t1 - t0) / 20 * 1e3, "drain ms/step", (t2 - t0) / 20 * 1e3)
    pr = cProfile.Profile(); pr.enable(); e2e(20); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
