"""CPU test (world_size 2, gloo) of the train step's gradient exchange host logic
(dualdiffusion_b200/ddp.py): bucket-wise all-reduce(mean) of the flat gradient buffer as buckets complete,
`.grad` installed as views of that buffer, accumulation across micro-steps under no_sync().  The backward
schedule itself (CUDA kernels) is replaced by a stand-in that fills the buckets with rank-dependent values;
the kernels are covered by tests/test_gpu_backward.py."""
import os
import socket
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dualdiffusion_b200.ddp import GradAllReducer


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_state():
    params = [torch.nn.Parameter(torch.zeros(6, 4)), torch.nn.Parameter(torch.zeros(10)), torch.nn.Parameter(torch.zeros(3, 3))]
    gains = [torch.nn.Parameter(torch.zeros([])), torch.nn.Parameter(torch.zeros([]))]
    total = sum(p.numel() for p in params) + len(gains)
    flat = torch.zeros(total)
    slots, off = {}, 0
    for i, p in enumerate(params):
        slots[f"p{i}"] = SimpleNamespace(param=p, grad=flat[off:off + p.numel()].view_as(p))
        off += p.numel()
    ts = SimpleNamespace(slots=slots, grad_flat=flat, dgains=flat[off:off + 2],
                         bucket_ranges=[(0, 24), (24, total)])        # bucket 0 = p0, bucket 1 = p1, p2, gains
    plan = SimpleNamespace(gain_params=gains, device=torch.device("cpu"))
    return params, gains, ts, plan


def _worker(rank: int, world: int, port: int, q) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params, gains, ts, plan = _fake_state()
    order = []

    def fake_backward(net, plan_, saved, dD, accumulate=False, bucket_done=None):
        for i, (lo, hi) in enumerate(ts.bucket_ranges):
            vals = torch.arange(lo, hi, dtype=torch.float32) * (rank + 1)      # rank-dependent "gradients"
            if accumulate:
                ts.grad_flat[lo:hi] += vals
            else:
                ts.grad_flat[lo:hi] = vals
            order.append((i, accumulate))
            bucket_done(i)
        return torch.zeros(1)

    sync = GradAllReducer()
    # step 1: fresh gradients, exchanged -> mean over ranks of arange * (rank+1) = arange * 1.5
    sync.run_backward(None, plan, ts, None, None, backward_fn=fake_backward)
    ref = torch.arange(ts.grad_flat.numel(), dtype=torch.float32) * 1.5
    ok1 = torch.allclose(ts.grad_flat, ref)
    views = all(s.param.grad.data_ptr() == s.grad.data_ptr() for s in ts.slots.values()) and \
        all(g.grad is not None and g.grad.data_ptr() == ts.dgains[i].data_ptr() for i, g in enumerate(gains))
    # step 2: two micro-steps; the first under no_sync accumulates locally, the second exchanges the sum
    for p in params + gains:
        p.grad.zero_()                                                         # optimizer.zero_grad(set_to_none=False)
    with sync.no_sync():
        sync.run_backward(None, plan, ts, None, None, backward_fn=fake_backward)
    local_only = torch.allclose(ts.grad_flat, torch.arange(ts.grad_flat.numel(), dtype=torch.float32) * (rank + 1))
    sync.run_backward(None, plan, ts, None, None, backward_fn=fake_backward)
    ok2 = torch.allclose(ts.grad_flat, 2 * ref)
    # step 3: grads dropped (set_to_none=True) -> overwrite, not accumulate
    for p in params + gains:
        p.grad = None
    sync.run_backward(None, plan, ts, None, None, backward_fn=fake_backward)
    ok3 = torch.allclose(ts.grad_flat, ref) and order[-1] == (1, False) and order[2] == (0, True)
    q.put((rank, ok1, views, local_only, ok2, ok3, sync.bytes_reduced))
    dist.barrier()
    dist.destroy_process_group()


def _worker_extras(rank: int, world: int, port: int, q) -> None:
    """A parameter whose gradient reaches .grad through ordinary autograd nodes (the UNet's emb_label /
    emb_label_unconditional / logvar_linear) must be exchanged too, after the whole backward pass, including the sum
    accumulated under no_sync()."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params, gains, ts, plan = _fake_state()

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.slots = torch.nn.ParameterList(params + gains)
            self.extra = torch.nn.Parameter(torch.ones(5))           # not in the flat gradient buffer

    net = Net()
    sync = GradAllReducer()

    def fake_backward(net_, plan_, saved, dD, accumulate=False, bucket_done=None):
        for i, (lo, hi) in enumerate(ts.bucket_ranges):
            vals = torch.full((hi - lo,), float(rank + 1))
            if accumulate:
                ts.grad_flat[lo:hi] += vals
            else:
                ts.grad_flat[lo:hi] = vals
            bucket_done(i)
        return torch.zeros(1)

    class Node(torch.autograd.Function):                             # stands in for UNetFunction
        @staticmethod
        def forward(ctx, x):
            return x.clone()

        @staticmethod
        def backward(ctx, g):
            sync.run_backward(net, plan, ts, None, None, backward_fn=fake_backward)
            return g

    def loss():
        return (Node.apply(net.extra * float(rank + 1))).sum()       # d/d extra = rank + 1

    loss().backward()
    ok1 = torch.allclose(net.extra.grad, torch.full((5,), 1.5)) and torch.allclose(ts.grad_flat, torch.full_like(ts.grad_flat, 1.5))
    net.extra.grad.zero_()
    for p in params + gains:
        p.grad.zero_()
    with sync.no_sync():
        loss().backward()
    local = torch.allclose(net.extra.grad, torch.full((5,), float(rank + 1)))
    loss().backward()
    ok2 = torch.allclose(net.extra.grad, torch.full((5,), 3.0)) and torch.allclose(ts.grad_flat, torch.full_like(ts.grad_flat, 3.0))
    every = {id(p) for p in net.parameters() if p.requires_grad}
    covered = {id(s.param) for s in ts.slots.values()} | {id(p) for p in plan.gain_params} | \
        {id(p) for p in GradAllReducer.extra_params(net, plan, ts)}
    q.put((rank, ok1, local, ok2, every == covered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_parameters_outside_the_flat_buffer_are_exchanged():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_extras, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok1, local, ok2, cover in res:
        assert ok1 and local and ok2 and cover, (rank, ok1, local, ok2, cover)


def test_two_rank_bucketed_gradient_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok1, views, local_only, ok2, ok3, nbytes in res:
        assert ok1 and views and local_only and ok2 and ok3, (rank, ok1, views, local_only, ok2, ok3)
        assert nbytes == 3 * 45 * 4                                           # three exchanged steps x 45 floats


def test_single_process_installs_views_without_exchange():
    params, gains, ts, plan = _fake_state()

    def fake_backward(net, plan_, saved, dD, accumulate=False, bucket_done=None):
        ts.grad_flat.fill_(2.0) if not accumulate else ts.grad_flat.add_(2.0)
        for i in range(len(ts.bucket_ranges)):
            bucket_done(i)
        return torch.zeros(1)

    sync = GradAllReducer()
    foreign = torch.ones(10)
    params[1].grad = foreign                                                   # a gradient that is not our view
    sync.run_backward(None, plan, ts, None, None, backward_fn=fake_backward)
    assert params[0].grad.data_ptr() == ts.slots["p0"].grad.data_ptr() and float(params[0].grad[0, 0]) == 2.0
    assert params[1].grad is foreign and torch.equal(foreign, torch.full((10,), 3.0))   # added into
    assert sync.bytes_reduced == 0
