"""Data-parallel gradient exchange for the UNet train step (reference: accelerate's DDP wrapper around the module,
training/trainer.py:1022-1044 -- one gradient all-reduce(mean) per optimizer step).

The backward schedule (modules/unets/unet_train.py) writes parameter gradients into one flat fp32 buffer laid out in
backward-completion order; `GradAllReducer` all-reduces each bucket over NCCL on its own stream as soon as the
bucket's weight-norm backward was enqueued, so the exchange of the decoder's gradients overlaps the encoder's
backward.  Parameters' `.grad` are views of that flat buffer (no per-parameter copies, no autograd accumulation).

    net.grad_sync = GradAllReducer()          # instead of wrapping the module in torch DDP
    ...
    with net.grad_sync.no_sync(): loss.backward()   # micro-steps of gradient accumulation
    loss.backward()                                  # last micro-step: overlapped all-reduce(mean)
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch
import torch.distributed as dist


class GradAllReducer:
    def __init__(self, process_group: Optional["dist.ProcessGroup"] = None) -> None:
        self.pg = process_group
        self.sync_now = True
        self.stream: Optional[torch.cuda.Stream] = None
        self.bytes_reduced = 0

    @contextlib.contextmanager
    def no_sync(self):
        old, self.sync_now = self.sync_now, False
        try:
            yield
        finally:
            self.sync_now = old

    def world_size(self) -> int:
        return dist.get_world_size(self.pg) if dist.is_available() and dist.is_initialized() else 1

    def run_backward(self, net, plan, ts, saved, dD) -> torch.Tensor:
        from .modules.unets.unet_train import train_backward
        params = [s for s in ts.slots.values()]
        installed = all(s.param.grad is not None and s.param.grad.data_ptr() == s.grad.data_ptr() for s in params)
        main = torch.cuda.current_stream(plan.device)
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=plan.device)
        exchange = self.sync_now and self.world_size() > 1

        def bucket_done(i: int) -> None:
            if not exchange:
                return
            lo, hi = ts.bucket_ranges[i]
            self.stream.wait_stream(main)
            with torch.cuda.stream(self.stream):
                dist.all_reduce(ts.grad_flat[lo:hi], op=dist.ReduceOp.AVG, group=self.pg)
            self.bytes_reduced += (hi - lo) * 4

        dlabel = train_backward(net, plan, saved, dD, accumulate=installed, bucket_done=bucket_done)
        if exchange:
            main.wait_stream(self.stream)
        if not installed:
            for s in params:
                if s.param.grad is None:
                    s.param.grad = s.grad
                elif s.param.grad.data_ptr() != s.grad.data_ptr():
                    s.param.grad.add_(s.grad)
            for i, p in enumerate(plan.gain_params):
                gview = ts.dgains[i]
                if p.grad is None:
                    p.grad = gview
                elif p.grad.data_ptr() != gview.data_ptr():
                    p.grad.add_(gview)
        return dlabel
