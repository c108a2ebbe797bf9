"""The DDP train-step leg of bench.py alone under torchrun (A/B runs of the gradient-exchange knobs):
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_train_only.py [label]"""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
import bench
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
d = bench.bench_train(dev, dist if world > 1 else None, world, steps=10, warmup=3)
if rank == 0:
    print(sys.argv[1] if len(sys.argv) > 1 else "", "train", round(d["value"], 1), "samples/s", round(d["ms_per_step"], 3), "ms", flush=True)
if os.environ.get("DD_DDP_TRACE") == "1" and rank == 0:
    import gc
    for o in gc.get_objects():
        if type(o).__name__ == "GradAllReducer" and getattr(o, "_events", None):
            print("   buckets (MB, all-reduce start, end; ms relative to the end of the backward schedule):", o.trace_report(), flush=True)
if world > 1:
    dist.destroy_process_group()
