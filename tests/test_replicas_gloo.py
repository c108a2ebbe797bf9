"""CPU test of the N>1 path (world_size 2, gloo): the sampler / codec shard as replicas over a partition of the
items, with torch.distributed used only for the max-over-ranks timing and the item counts."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dualdiffusion_b200 import replicas


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_items: int, q) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = replicas.shard_range(n_items, rank, world)
    seeds = replicas.shard_seeds(1000, n_items, rank, world)
    t = replicas.max_over_ranks(1.0 + rank, dist)              # rank 1 is the slow one
    counts = replicas.gather_counts(hi - lo, dist)
    q.put((rank, lo, hi, seeds, t, list(counts)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_replica_sharding():
    world, n_items = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = []
    for rank, lo, hi, seeds, t, counts in res:
        covered += list(range(lo, hi))
        assert seeds == [1000 + i for i in range(lo, hi)]      # item -> seed is independent of the sharding
        assert t == 2.0                                        # max over ranks
        assert counts == [4, 3] and sum(counts) == n_items
    assert covered == list(range(n_items))                     # disjoint and complete


def test_shard_range_properties():
    for n in (0, 1, 5, 64, 129):
        for world in (1, 2, 3, 8):
            spans = [replicas.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
